// rs_pcl_b200 -- headless counterpart of the reference CLI's registration modes (main.cpp:185-237), on the B200 path.
//
//   rs_pcl_b200 [--dataset DIR] [--scheme ndt|icp|incremental] --registration <prefix> [deg] <N>
//       loads DIR/<prefix>-<i>.pcd (i < N), runs the scheme (default: NDT edge-based, like main.cpp:208,218), writes
//       DIR/<prefix>-registration (main.cpp:87, no suffix) and prints the per-frame transforms as one JSON line.
//       --imu FILE: IMU trace ("ts_ms kind x y z" rows, kind 0 = gyro, 1 = accel) followed by a line "frames t0 t1 ..";
//       the angle triples of the complementary filter (rotation_estimator.hpp) replace the fixed-degree guess, as in
//       the reference's capture modes (main.cpp:129).
//       --per-call: run the sweep through the PCL-shaped objects call by call (upload / download per call) instead of the
//       device-resident rspcl_register_sequence (default).
//       --dump-pairs: also write DIR/<prefix>-{src,tgt,coarse}-<k>.pcd, the exact inputs of frame k's coarse and fine
//       stages (parity aid, in the spirit of the reference's own dataset/edge-<i>.pcd dumps, icp:66-69).
//   rs_pcl_b200 [--dataset DIR] --edges <file>
//       extract_edge_features of DIR/<file> (main.cpp:58-62); prints the edge count, writes DIR/<file>.edges.pcd.
// The GL viewer loops of the reference (main.cpp:65-73,90-98) and the capture modes are out of scope.
#include <cstdlib>
#include <cstring>
#include "rspcl.hpp"

static void print_mats(const char* key, const std::vector<rspcl::Matrix4f>& T) {
  std::printf(", \"%s\": [", key);
  for (size_t k = 0; k < T.size(); ++k) {
    std::printf("%s[", k ? ", " : "");
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) std::printf("%s%.9g", (r || c) ? ", " : "", T[k](r, c));  // row-major rows
    std::printf("]");
  }
  std::printf("]");
}

static void print_json(const std::vector<rspcl::Matrix4f>& T, const std::vector<int>& acc, size_t npts,
                       const std::vector<rspcl::Matrix4f>& Tc, const std::vector<rspcl::Matrix4f>& Tf) {
  std::printf("{\"points\": %zu, \"accepted\": [", npts);
  for (size_t i = 0; i < acc.size(); ++i) std::printf("%s%d", i ? ", " : "", acc[i]);
  std::printf("]");
  print_mats("transforms", T);
  if (!Tc.empty()) print_mats("coarse", Tc);
  if (!Tf.empty()) print_mats("fine", Tf);
  std::printf("}\n");
}

int main(int argc, char** argv) {
  try {
    std::string dir = "dataset", scheme = "ndt", imu_path;
    bool dump_pairs = false, per_call = false;
    std::vector<std::string> a;
    for (int i = 1; i < argc; ++i) {
      if (!std::strcmp(argv[i], "--dataset") && i + 1 < argc) dir = argv[++i];
      else if (!std::strcmp(argv[i], "--scheme") && i + 1 < argc) scheme = argv[++i];
      else if (!std::strcmp(argv[i], "--imu") && i + 1 < argc) imu_path = argv[++i];
      else if (!std::strcmp(argv[i], "--dump-pairs")) dump_pairs = true;
      else if (!std::strcmp(argv[i], "--per-call")) per_call = true;
      else a.push_back(argv[i]);
    }
    if (a.size() >= 2 && a[0] == "--edges") {
      rgb_point_cloud_pointer c(new rgb_point_cloud);
      rspcl::io::loadPCDFile(dir + "/" + a[1], *c);
      auto e = rspcl::extract_edge_features(c);
      rspcl::io::savePCDFileBinary(dir + "/" + a[1] + ".edges.pcd", *e);
      std::printf("{\"edge_points\": %zu}\n", e->size());
      return 0;
    }
    if (a.size() >= 3 && a[0] == "--registration") {
      const std::string prefix = a[1];
      float rads = -0.523599f;
      int frames;
      if (a.size() == 3) {
        frames = std::atoi(a[2].c_str());
      } else {
        const double deg = std::atof(a[2].c_str());
        rads = static_cast<float>((deg / 180.0) * M_PI);  // main.cpp:214-215
        frames = std::atoi(a[3].c_str());
      }
      std::vector<rgb_point_cloud_pointer> clouds;
      for (int i = 0; i < frames; ++i) {
        rgb_point_cloud_pointer c(new rgb_point_cloud);
        rspcl::io::loadPCDFile(dir + "/" + prefix + "-" + std::to_string(i) + ".pcd", *c);
        clouds.push_back(c);
      }
      rgb_point_cloud_pointer result;
      std::vector<rspcl::Matrix4f> T, Tc, Tf;
      std::vector<int> acc;
      std::vector<rs_float3> thetas;
      if (!imu_path.empty()) {
        std::ifstream f(imu_path);
        if (!f) throw rspcl::Error("cannot open " + imu_path);
        std::vector<ImuSample> trace;
        std::vector<double> frame_ts;
        std::string tok;
        while (f >> tok) {
          if (tok == "frames") {
            double t;
            while (f >> t) frame_ts.push_back(t);
            break;
          }
          ImuSample s;
          s.ts_ms = std::atof(tok.c_str());
          if (!(f >> s.kind >> s.v.x >> s.v.y >> s.v.z)) throw rspcl::Error("bad IMU row in " + imu_path);
          trace.push_back(s);
        }
        if ((int)frame_ts.size() != frames) throw rspcl::Error("IMU trace needs one timestamp per frame");
        thetas = thetas_from_imu_trace(trace, frame_ts);
      }
      if (scheme == "incremental") {
        IncrementalICP s;
        result = s.registration(clouds);
        T = s.transforms;
        acc.assign(T.size(), 1);
      } else if (scheme == "icp") {
        ICPEdgeBasedRegistration s_imu(thetas), s_fix(rads);
        ICPEdgeBasedRegistration& s = thetas.empty() ? s_fix : s_imu;
        if (dump_pairs) s.dump_pairs_prefix = dir + "/" + prefix;
        s.device_resident = !per_call;
        result = s.registration(clouds);
        T = s.transforms, acc = s.accepted, Tc = s.coarse_transforms, Tf = s.fine_transforms;
      } else {
        NDTEdgeBasedRegistration s_imu(thetas), s_fix(rads);
        NDTEdgeBasedRegistration& s = thetas.empty() ? s_fix : s_imu;
        if (dump_pairs) s.dump_pairs_prefix = dir + "/" + prefix;
        s.device_resident = !per_call;
        result = s.registration(clouds);
        T = s.transforms, acc = s.accepted, Tc = s.coarse_transforms, Tf = s.fine_transforms;
      }
      rspcl::io::savePCDFileBinary(dir + "/" + prefix + "-registration", *result);
      print_json(T, acc, result->size(), Tc, Tf);
      return 0;
    }
    std::fprintf(stderr, "usage: rs_pcl_b200 [--dataset DIR] [--scheme ndt|icp|incremental] [--imu FILE] --registration <prefix> [deg] <N>\n"
                         "       rs_pcl_b200 [--dataset DIR] --edges <file>\n");
    return 2;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "rs_pcl_b200: %s\n", e.what());
    return 1;
  }
}
