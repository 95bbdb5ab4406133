// Host-only check of the facade's IMU front end: replays a trace given on stdin and prints theta per frame.
// stdin: n_samples, then rows "ts kind x y z"; n_frames, then frame timestamps.  (tests/test_imu.py drives it.)
#include "../../realsense-pointcloud_b200/host/rspcl.hpp"

int main() {
  int n = 0;
  if (std::scanf("%d", &n) != 1) return 2;
  std::vector<ImuSample> tr(n);
  for (auto& s : tr)
    if (std::scanf("%lf %d %f %f %f", &s.ts_ms, &s.kind, &s.v.x, &s.v.y, &s.v.z) != 5) return 2;
  int f = 0;
  if (std::scanf("%d", &f) != 1) return 2;
  std::vector<double> ts(f);
  for (auto& t : ts)
    if (std::scanf("%lf", &t) != 1) return 2;
  const std::vector<rs_float3> th = thetas_from_imu_trace(tr, ts);
  for (const auto& t : th) std::printf("%.9g %.9g %.9g\n", t.x, t.y, t.z);
  // the guess matrices the two schemes derive from the last angle triple (icp:86-92, ndt:79)
  rs_float3 rel = th.back();
  rel.add(-th[0].x, -th[0].y, -th[0].z);
  const rspcl::Matrix4f gi = rspcl::Matrix4f::AngleAxis(rel.x, 2) * rspcl::Matrix4f::AngleAxis(-rel.y, 1) * rspcl::Matrix4f::AngleAxis(rel.z, 0);
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) std::printf("%.9g%c", gi(r, c), (r == 3 && c == 3) ? '\n' : ' ');
  return 0;
}
