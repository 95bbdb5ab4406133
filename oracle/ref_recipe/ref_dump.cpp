// ref_dump.cpp -- headless driver around the REFERENCE'S OWN registration headers (hyunminch/realsense-pointcloud),
// linked against real PCL.  Test infrastructure: it produces the fixtures that pin oracle/ (and through it the CUDA path)
// to PCL.  The reference's main() cannot be used: it opens an OpenCV window unconditionally (main.cpp:186) and ends in a
// GL loop, so this file calls the same classes the way main.cpp:76-88 / :117-134 do.
//
//   ref_dump <dataset dir> <prefix> <n frames> <out dir> [rads]
//
// reads <dataset>/<prefix>-<i>.pcd (main.cpp:81 naming; tools/gen_scene.py writes them) and dumps, per frame / pair:
//   edges_<i>        extract_edge_features(cloud)                    (edge_extractor.hpp:7-39)      [n,4] f32 x y z rgb-bits
//   voxel_<i>        ApproximateVoxelGrid 1 cm on edges_<i>          (icp:47,59-60,75-76)
//   crop_<i>         BlurFilter::filter(cloud)                       (blur_filter.hpp:18-36)
//   icp_T_<i>, icp_meta_<i>, icp_aligned_<i>   coarse ICP of voxel_<i> onto voxel_<i-1>, reference settings, guess R_y(rads)
//   icp10_T_<i>      the same pair with 10 forced iterations (eps tightened)
//   fit_<i>          getFitnessScore() of that align
//   ndt_T_<i>, ndt_meta_<i>                    NDT of the same pair (ndt:38-43 settings)
//   scheme_icp, scheme_ndt [, scheme_incr]     the merged cloud registration() returns (types.hpp:30-43)
// Every array is a raw little-endian file <name>.bin listed in manifest.txt as "<name> <dtype> <rows> <cols>".
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <pcl/point_types.h>
#include <pcl/point_cloud.h>
#include <pcl/common/transforms.h>
#include <pcl/common/io.h>
#include <pcl/io/pcd_io.h>
#include <pcl/features/integral_image_normal.h>
#include <pcl/features/organized_edge_detection.h>
#include <pcl/filters/approximate_voxel_grid.h>
#include <pcl/registration/icp.h>
#include <pcl/registration/ndt.h>
#include <pcl/registration/correspondence_rejection_trimmed.h>

// The IMU angle triple of the reference (utils.hpp:30-62).  utils.hpp itself is GLFW / OpenGL / librealsense rendering
// code, so only this 12-byte type is restated here (same members and operations, which is all icp:83-88 / ndt:76-79 use).
struct float3 {
  float x, y, z;
  float3 operator*(float t) { return {x * t, y * t, z * t}; }
  float3 operator-(float t) { return {x - t, y - t, z - t}; }
  void operator*=(float t) { x = x * t; y = y * t; z = z * t; }
  void operator=(float3 other) { x = other.x; y = other.y; z = other.z; }
  void add(float t1, float t2, float t3) { x += t1; y += t2; z += t3; }
};

// ---- the reference's own code, included from where it lies (CMake adds REFERENCE_SRC to the include path)
#include "types.hpp"
#include "edge_extractor.hpp"
#include "blur_filter.hpp"
#ifdef RSPCL_REF_WITH_INCREMENTAL
#include "incremental_icp.hpp"
#endif
#include "icp_edge_based_registration.hpp"
#include "ndt_edge_based_registration.hpp"

static std::string g_out;
static std::ofstream g_manifest;

static void dump(const std::string& name, const char* dtype, const void* data, size_t rows, size_t cols, size_t elem) {
  std::ofstream f(g_out + "/" + name + ".bin", std::ios::binary);
  f.write(reinterpret_cast<const char*>(data), (std::streamsize)(rows * cols * elem));
  g_manifest << name << " " << dtype << " " << rows << " " << cols << "\n";
}

static void dump_cloud(const std::string& name, const rgb_point_cloud& c) {
  std::vector<float> v(c.size() * 4);
  for (size_t i = 0; i < c.size(); ++i) {
    v[4 * i + 0] = c.points[i].x;
    v[4 * i + 1] = c.points[i].y;
    v[4 * i + 2] = c.points[i].z;
    std::memcpy(&v[4 * i + 3], &c.points[i].rgb, 4);  // packed bgra bits, as in a .pcd row
  }
  dump(name, "f32", v.data(), c.size(), 4, 4);
}

static void dump_mat(const std::string& name, const Eigen::Matrix4f& T) {
  float m[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) m[r * 4 + c] = T(r, c);  // row-major
  dump(name, "f32", m, 4, 4, 4);
}

template <typename Reg>
static void set_icp(Reg& icp, int iters, double teps, double feps) {
  icp.setMaximumIterations(iters);        // icp:42
  icp.setMaxCorrespondenceDistance(0.01); // icp:43
  icp.setTransformationEpsilon(teps);     // icp:44
  icp.setEuclideanFitnessEpsilon(feps);   // icp:45
}

int main(int argc, char** argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: ref_dump <dataset dir> <prefix> <n frames> <out dir> [rads]\n");
    return 2;
  }
  const std::string dir = argv[1], prefix = argv[2];
  const int n = std::atoi(argv[3]);
  g_out = argv[4];
  const float rads = argc > 5 ? (float)std::atof(argv[5]) : -0.523599f;  // icp:135
  g_manifest.open(g_out + "/manifest.txt");
  if (!g_manifest) {
    std::fprintf(stderr, "cannot write %s/manifest.txt (create the directory first)\n", g_out.c_str());
    return 2;
  }
  g_manifest << "# produced by oracle/ref_recipe/ref_dump with PCL " << PCL_VERSION_PRETTY << "\n";

  std::vector<rgb_point_cloud_pointer> clouds;
  for (int i = 0; i < n; ++i) {  // main.cpp:79-83
    rgb_point_cloud_pointer c(new rgb_point_cloud);
    const std::string path = dir + "/" + prefix + "-" + std::to_string(i) + ".pcd";
    if (pcl::io::loadPCDFile(path, *c) != 0) {
      std::fprintf(stderr, "cannot read %s\n", path.c_str());
      return 1;
    }
    clouds.push_back(c);
  }

  // ---- stage-by-stage fixtures
  std::vector<rgb_point_cloud_pointer> voxels;
  pcl::ApproximateVoxelGrid<rgb_point> avg;
  avg.setLeafSize(0.01, 0.01, 0.01);  // icp:47
  for (int i = 0; i < n; ++i) {
    rgb_point_cloud_pointer e = extract_edge_features(clouds[i]);  // edge_extractor.hpp:7
    dump_cloud("edges_" + std::to_string(i), *e);
    rgb_point_cloud_pointer v(new rgb_point_cloud);
    avg.setInputCloud(e);
    avg.filter(*v);
    dump_cloud("voxel_" + std::to_string(i), *v);
    voxels.push_back(v);
    rgb_point_cloud_pointer cr(new rgb_point_cloud(*clouds[i]));
    BlurFilter().filter(cr);  // blur_filter.hpp:18
    dump_cloud("crop_" + std::to_string(i), *cr);
  }
  for (int i = 1; i < n; ++i) {
    const std::string s = std::to_string(i);
    Eigen::AngleAxisf ry(rads, Eigen::Vector3f::UnitY());
    const Eigen::Matrix4f guess = (Eigen::Translation3f(0, 0, 0) * ry).matrix();  // icp:98-100
    {
      pcl::IterativeClosestPoint<rgb_point, rgb_point> icp;
      set_icp(icp, 100, 1, 1000);
      icp.setInputSource(voxels[i]);
      icp.setInputTarget(voxels[i - 1]);
      rgb_point_cloud aligned;
      icp.align(aligned, guess);
      dump_mat("icp_T_" + s, icp.getFinalTransformation());
      dump_cloud("icp_aligned_" + s, aligned);
      const double meta[2] = {icp.hasConverged() ? 1.0 : 0.0, icp.getFitnessScore()};
      dump("icp_meta_" + s, "f64", meta, 1, 2, 8);
    }
    {
      pcl::IterativeClosestPoint<rgb_point, rgb_point> icp;  // BASELINE configs[1] style: the iterations are forced
      set_icp(icp, 10, 1e-30, -1e300);
      icp.setInputSource(voxels[i]);
      icp.setInputTarget(voxels[i - 1]);
      rgb_point_cloud aligned;
      icp.align(aligned, guess);
      dump_mat("icp10_T_" + s, icp.getFinalTransformation());
    }
    {
      pcl::NormalDistributionsTransform<rgb_point, rgb_point> ndt;  // ndt:38-43
      ndt.setTransformationEpsilon(0.01);
      ndt.setStepSize(0.1);
      ndt.setResolution(1.0);
      ndt.setMaximumIterations(50);
      ndt.setInputSource(voxels[i]);
      ndt.setInputTarget(voxels[i - 1]);
      rgb_point_cloud aligned;
      ndt.align(aligned, guess);
      dump_mat("ndt_T_" + s, ndt.getFinalTransformation());
      const double meta[3] = {ndt.hasConverged() ? 1.0 : 0.0, (double)ndt.getFinalNumIteration(), ndt.getTransformationProbability()};
      dump("ndt_meta_" + s, "f64", meta, 1, 3, 8);
    }
  }

  // ---- the schemes, exactly as main.cpp drives them (types.hpp:30-43).  They mutate their inputs (icp:54,59-60), so each
  // gets fresh copies; they also write dataset/edge-<i>.pcd / dataset/edge_cloud.pcd (icp:68,126): run from a directory
  // that has a dataset/ folder.
  auto fresh = [&]() {
    std::vector<rgb_point_cloud_pointer> c;
    for (auto& p : clouds) c.push_back(rgb_point_cloud_pointer(new rgb_point_cloud(*p)));
    return c;
  };
  {
    auto c = fresh();
    ICPEdgeBasedRegistration scheme(rads);  // icp:17-19
    dump_cloud("scheme_icp", *scheme.registration(c));
  }
  {
    auto c = fresh();
    NDTEdgeBasedRegistration scheme(rads);  // main.cpp:218
    dump_cloud("scheme_ndt", *scheme.registration(c));
  }
#ifdef RSPCL_REF_WITH_INCREMENTAL
  {
    auto c = fresh();
    IncrementalICP scheme;  // incremental_icp.hpp:33
    dump_cloud("scheme_incr", *scheme.registration(c));
  }
#endif
  std::cout << "fixtures written to " << g_out << std::endl;
  return 0;
}
