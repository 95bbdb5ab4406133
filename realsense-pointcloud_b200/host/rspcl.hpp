// rspcl.hpp -- C++ host facade over the C ABI (include/rspcl.h): the reference's registration surface with the
// arithmetic running on a B200.
//
// It offers, under the reference's own names, what /root/reference/src calls on its hot path:
//   types.hpp:8-10      rgb_point / rgb_point_cloud / rgb_point_cloud_pointer
//   types.hpp:14-44     RegistrationScheme, TwoPhaseRegistrationScheme
//   edge_extractor.hpp  extract_edge_features            blur_filter.hpp  BlurFilter
//   icp_edge_based_registration.hpp  ICPEdgeBasedRegistration     (IMU thetas or fixed angle)
//   ndt_edge_based_registration.hpp  NDTEdgeBasedRegistration
//   incremental_icp.hpp              IncrementalICP
// and the PCL-shaped objects those headers use (IterativeClosestPoint, NormalDistributionsTransform,
// ApproximateVoxelGrid, transformPointCloud, operator+).  Header-only, C++14, no dependency beyond librspcl_b200.so.
// There is no CPU fallback: every arithmetic call goes through the C ABI and throws rspcl::Error on failure.
#ifndef RSPCL_HPP
#define RSPCL_HPP
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rspcl.h"

namespace rspcl {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ------------------------------------------------------------------ point / cloud / matrix types
struct alignas(16) PointXYZRGB {  // memory layout of pcl::PointXYZRGB == RSPCL_LAYOUT_PCL32
  float x = 0.f, y = 0.f, z = 0.f, data3 = 1.f;
  std::uint32_t rgba = 0xff000000u;
  std::uint32_t pad[3] = {0, 0, 0};
};
static_assert(sizeof(PointXYZRGB) == 32, "PointXYZRGB must match pcl::PointXYZRGB");

struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud>;
  std::vector<PointXYZRGB> points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  bool isOrganized() const { return height > 1; }
  PointCloud& operator+=(const PointCloud& rhs) {  // pcl::PointCloud::operator+= : lhs points first, height = 1
    points.insert(points.end(), rhs.points.begin(), rhs.points.end());
    width = static_cast<std::uint32_t>(points.size());
    height = 1;
    is_dense = is_dense && rhs.is_dense;
    return *this;
  }
  PointCloud operator+(const PointCloud& rhs) const {
    PointCloud out = *this;
    out += rhs;
    return out;
  }
};

struct Matrix4f {  // column-major like Eigen::Matrix4f
  float m[16];
  static Matrix4f Identity() {
    Matrix4f r;
    for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.f : 0.f;
    return r;
  }
  float& operator()(int r, int c) { return m[c * 4 + r]; }
  float operator()(int r, int c) const { return m[c * 4 + r]; }
  const float* data() const { return m; }
  float* data() { return m; }
  Matrix4f operator*(const Matrix4f& b) const {
    Matrix4f o;
    for (int c = 0; c < 4; ++c)
      for (int r = 0; r < 4; ++r) {
        float s = 0.f;
        for (int k = 0; k < 4; ++k) s += (*this)(r, k) * b(k, c);
        o(r, c) = s;
      }
    return o;
  }
  // Eigen::AngleAxisf(angle, unit axis).matrix(), axis 0/1/2 = X/Y/Z
  static Matrix4f AngleAxis(float angle, int axis) {
    Matrix4f o = Identity();
    const float c = std::cos(angle), s = std::sin(angle);
    const int a = (axis + 1) % 3, b = (axis + 2) % 3;
    o(a, a) = c;
    o(a, b) = -s;
    o(b, a) = s;
    o(b, b) = c;
    return o;
  }
};

}  // namespace rspcl

// The reference's typedefs (types.hpp:8-10) and IMU triple (utils.hpp:30-62; renamed member-compatible struct, because
// `float3` collides with CUDA's vector type -- SURVEY H8)
typedef rspcl::PointXYZRGB rgb_point;
typedef rspcl::PointCloud rgb_point_cloud;
typedef rgb_point_cloud::Ptr rgb_point_cloud_pointer;
struct rs_float3 {
  float x, y, z;
  rs_float3 operator*(float t) const { return {x * t, y * t, z * t}; }
  rs_float3 operator-(float t) const { return {x - t, y - t, z - t}; }
  void operator*=(float t) { x *= t, y *= t, z *= t; }
  void add(float t1, float t2, float t3) { x += t1, y += t2, z += t3; }
};

// ------------------------------------------------------------------ IMU initial-guess front end
// Host-side complementary filter that turns gyro / accelerometer samples into the per-frame angle triple the scheme
// constructors take (rotation_estimator.hpp:22-79: gyro integration between samples, accelerometer tilt blended with
// weight 1 - alpha, yaw initialised to PI because gravity says nothing about it).  `rs_vector` stands in for
// librealsense's rs2_vector; timestamps are milliseconds like rs2_frame::get_timestamp().  Not thread-safe: the
// reference guards theta with a mutex because its callbacks run on the sensor thread, a replayed trace does not.
struct rs_vector {
  float x, y, z;
};

class RotationEstimator {
 public:
  explicit RotationEstimator(float alpha = 0.98f) : alpha_(alpha) {}

  void process_gyro(rs_vector gyro, double ts_ms) {
    if (first_) {  // until the first accelerometer sample fixes the initial pose only the clock is tracked
      last_ts_gyro_ = ts_ms;
      return;
    }
    const float dt = static_cast<float>((ts_ms - last_ts_gyro_) / 1000.0);
    last_ts_gyro_ = ts_ms;
    // gyro x/y/z = pitch/yaw/roll rates; theta.x <- -roll, theta.y <- -yaw, theta.z <- +pitch (rotation_estimator.hpp:45)
    theta_.add(-(gyro.z * dt), -(gyro.y * dt), gyro.x * dt);
  }

  void process_accel(rs_vector accel) {
    // the reference calls the C double atan2 / sqrt on float arguments and narrows the result
    const float az = static_cast<float>(std::atan2(static_cast<double>(accel.y), static_cast<double>(accel.z)));
    const float ax = static_cast<float>(std::atan2(static_cast<double>(accel.x),
                                                   std::sqrt(static_cast<double>(accel.y * accel.y + accel.z * accel.z))));
    if (first_) {
      first_ = false;
      theta_.x = ax;
      theta_.y = 3.14159265358979323846f;  // PI (utils.hpp)
      theta_.z = az;
    } else {
      theta_.x = theta_.x * alpha_ + ax * (1.0f - alpha_);
      theta_.z = theta_.z * alpha_ + az * (1.0f - alpha_);
    }
  }

  rs_float3 get_theta() const { return theta_; }

 private:
  rs_float3 theta_{0.f, 0.f, 0.f};
  float alpha_;
  bool first_ = true;
  double last_ts_gyro_ = 0;
};

// Replays a recorded IMU trace and samples theta at each frame's capture time -- what capture.hpp does live when it
// pushes estimator.get_theta() next to every captured cloud.  Trace rows: {ts_ms, kind (0 gyro, 1 accel), x, y, z},
// ascending in time; a frame takes the estimate after all samples with ts <= its timestamp.
struct ImuSample {
  double ts_ms;
  int kind;
  rs_vector v;
};

inline std::vector<rs_float3> thetas_from_imu_trace(const std::vector<ImuSample>& trace, const std::vector<double>& frame_ts_ms,
                                                    float alpha = 0.98f) {
  RotationEstimator est(alpha);
  std::vector<rs_float3> out;
  std::size_t k = 0;
  for (double t : frame_ts_ms) {
    for (; k < trace.size() && trace[k].ts_ms <= t; ++k) {
      if (trace[k].kind == 0) est.process_gyro(trace[k].v, trace[k].ts_ms);
      else est.process_accel(trace[k].v);
    }
    out.push_back(est.get_theta());
  }
  return out;
}

namespace rspcl {

// ------------------------------------------------------------------ device context + RAII cloud handles
class Device {
 public:
  static Device& get(int device = 0) {
    static Device d(device);
    return d;
  }
  rspcl_ctx* ctx() const { return ctx_; }
  void check(int rc, const char* what) const {
    if (rc != RSPCL_OK) throw Error(std::string(what) + ": " + rspcl_last_error(ctx_));
  }
  ~Device() { rspcl_ctx_destroy(ctx_); }

 private:
  explicit Device(int device) {
    if (rspcl_ctx_create(device, &ctx_) != RSPCL_OK)
      throw Error("rspcl_ctx_create failed: no CUDA device (there is no CPU fallback)");
  }
  rspcl_ctx* ctx_ = nullptr;
};

class DeviceCloud {
 public:
  DeviceCloud(int n_seg, int stride) {
    Device::get().check(rspcl_cloud_create(Device::get().ctx(), n_seg, stride > 0 ? stride : 1, &h_), "cloud_create");
  }
  explicit DeviceCloud(const PointCloud& c) : DeviceCloud(1, static_cast<int>(c.size())) { upload(c); }
  DeviceCloud(const DeviceCloud&) = delete;
  DeviceCloud& operator=(const DeviceCloud&) = delete;
  ~DeviceCloud() { rspcl_cloud_destroy(Device::get().ctx(), h_); }
  void upload(const PointCloud& c) {
    const std::int32_t n = static_cast<std::int32_t>(c.size());
    const bool org = c.isOrganized() && c.width * c.height == c.size();
    static const PointXYZRGB dummy;
    Device::get().check(rspcl_cloud_upload(Device::get().ctx(), h_, n ? c.points.data() : &dummy, RSPCL_LAYOUT_PCL32, &n, 1,
                                           org ? static_cast<int>(c.width) : 0, org ? static_cast<int>(c.height) : 0),
                        "cloud_upload");
  }
  // n organized clouds of equal dimensions -> the n segments of this batch
  template <typename PtrVec>
  void upload_batch(const PtrVec& clouds) {
    std::vector<PointXYZRGB> all;
    std::vector<std::int32_t> counts;
    for (const auto& c : clouds) {
      all.insert(all.end(), c->points.begin(), c->points.end());
      counts.push_back(static_cast<std::int32_t>(c->size()));
    }
    Device::get().check(rspcl_cloud_upload(Device::get().ctx(), h_, all.data(), RSPCL_LAYOUT_PCL32, counts.data(),
                                           static_cast<int>(counts.size()), static_cast<int>(clouds[0]->width),
                                           static_cast<int>(clouds[0]->height)),
                        "cloud_upload");
    Device::get().check(rspcl_ctx_sync(Device::get().ctx()), "sync");  // `all` is pageable staging
  }
  void download(PointCloud& out) const {
    std::int32_t n = 0;
    Device::get().check(rspcl_cloud_counts(Device::get().ctx(), h_, &n), "cloud_counts");
    out.points.resize(static_cast<std::size_t>(n));
    if (n)
      Device::get().check(rspcl_cloud_download(Device::get().ctx(), h_, out.points.data(), RSPCL_LAYOUT_PCL32, n, nullptr),
                          "cloud_download");
    int w = 0, h = 0;
    rspcl_cloud_dims(h_, &w, &h);
    if (h > 1 && w * h == n) {
      out.width = w;
      out.height = h;
    } else {
      out.width = static_cast<std::uint32_t>(n);
      out.height = 1;
    }
  }
  rspcl_cloud* handle() const { return h_; }

 private:
  rspcl_cloud* h_ = nullptr;
};

// ------------------------------------------------------------------ free functions the reference calls
// pcl::transformPointCloud(in, out, Matrix4f) -- icp:116-117 (in == out allowed)
inline void transformPointCloud(const PointCloud& in, PointCloud& out, const Matrix4f& T) {
  DeviceCloud d(in);
  Device::get().check(rspcl_transform(Device::get().ctx(), d.handle(), T.data(), 1, d.handle()), "transform");
  const std::uint32_t w = in.width, h = in.height;
  const bool dense = in.is_dense;
  d.download(out);
  out.width = w;
  out.height = h;
  out.is_dense = dense;
}

// edge_extractor.hpp:7 -- RGB-Canny edge points of an organized cloud, row-major order
inline rgb_point_cloud_pointer extract_edge_features(rgb_point_cloud_pointer cloud) {
  if (!cloud->isOrganized()) throw Error("extract_edge_features needs an organized cloud");
  DeviceCloud d(*cloud);
  DeviceCloud e(1, static_cast<int>(cloud->size()));
  Device::get().check(rspcl_edge_extract(Device::get().ctx(), d.handle(), 40.f, 100.f, e.handle(), nullptr), "edge_extract");
  rgb_point_cloud_pointer out(new rgb_point_cloud);
  e.download(*out);
  out->is_dense = cloud->is_dense;
  return out;
}

// ------------------------------------------------------------------ pcl::ApproximateVoxelGrid
class ApproximateVoxelGrid {
 public:
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx, leaf_[1] = ly, leaf_[2] = lz; }
  void setInputCloud(const rgb_point_cloud_pointer& c) { input_ = c; }
  void filter(PointCloud& out) {  // out may alias the input (icp:59-60)
    if (!input_) throw Error("ApproximateVoxelGrid: no input cloud");
    DeviceCloud d(*input_);
    Device::get().check(rspcl_voxel_approx(Device::get().ctx(), d.handle(), leaf_, d.handle()), "voxel_approx");
    d.download(out);
    out.is_dense = false;
  }

 private:
  float leaf_[3] = {1.f, 1.f, 1.f};  // PCL default leaf (IncrementalICP never calls setLeafSize, incr:36,54-55)
  rgb_point_cloud_pointer input_;
};

// ------------------------------------------------------------------ pcl::Registration surface
class RegistrationBase {
 public:
  virtual ~RegistrationBase() {}
  void setInputSource(const rgb_point_cloud_pointer& c) { source_ = c; }
  void setInputTarget(const rgb_point_cloud_pointer& c) { target_ = c; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  bool hasConverged() const { return converged_; }
  Matrix4f getFinalTransformation() const { return final_; }
  void align(PointCloud& out) { align(out, Matrix4f::Identity()); }
  virtual void align(PointCloud& out, const Matrix4f& guess) = 0;
  // Registration::getFitnessScore: mean squared NN distance of the transformed source (<= max_range)
  double getFitnessScore(double max_range = DBL_MAX) {
    if (!source_ || !target_) throw Error("getFitnessScore: source/target not set");
    DeviceCloud s(*source_), t(*target_);
    Device::get().check(rspcl_transform(Device::get().ctx(), s.handle(), final_.data(), 1, s.handle()), "transform");
    double f = 0;
    Device::get().check(rspcl_fitness(Device::get().ctx(), s.handle(), t.handle(), max_range, &f), "fitness");
    return f;
  }

 protected:
  rgb_point_cloud_pointer source_, target_;
  int max_iterations_ = 10;
  double transformation_epsilon_ = 0.0;
  bool converged_ = false;
  Matrix4f final_ = Matrix4f::Identity();
};

class IterativeClosestPoint : public RegistrationBase {
 public:
  IterativeClosestPoint() { rspcl_icp_reference_params(&prm_), prm_.max_iterations = 10, prm_.max_corr_dist = std::sqrt(DBL_MAX),
                            prm_.transformation_epsilon = 0.0, prm_.euclidean_fitness_epsilon = -DBL_MAX; }
  void setMaxCorrespondenceDistance(double d) { prm_.max_corr_dist = d; }
  void setEuclideanFitnessEpsilon(double e) { prm_.euclidean_fitness_epsilon = e; }
  using RegistrationBase::align;
  void align(PointCloud& out, const Matrix4f& guess) override {
    if (!source_ || !target_) throw Error("ICP: source/target not set");
    prm_.max_iterations = max_iterations_;
    prm_.transformation_epsilon = transformation_epsilon_;
    DeviceCloud s(*source_), t(*target_), a(1, static_cast<int>(source_->size()));
    rspcl_icp_result r;
    r.prev_mse = prev_mse_;  // DefaultConvergenceCriteria state persists across align() calls on one object
    Device::get().check(rspcl_icp_align(Device::get().ctx(), s.handle(), t.handle(), &prm_, guess.data(), &r, a.handle(), nullptr),
                        "icp_align");
    prev_mse_ = r.prev_mse;
    converged_ = r.converged != 0;
    std::memcpy(final_.m, r.T, sizeof(r.T));
    state_ = r.state;
    iterations_ = r.iterations;
    a.download(out);
    out.width = source_->width, out.height = source_->height, out.is_dense = source_->is_dense;
  }
  int convergenceState() const { return state_; }
  int iterations() const { return iterations_; }

 private:
  rspcl_icp_params prm_;
  double prev_mse_ = DBL_MAX;
  int state_ = 0, iterations_ = 0;
};

class NormalDistributionsTransform : public RegistrationBase {
 public:
  NormalDistributionsTransform() {
    rspcl_ndt_reference_params(&prm_);
    max_iterations_ = 35;  // PCL defaults; the reference overrides them (ndt:39-43)
    transformation_epsilon_ = 0.1;
  }
  void setStepSize(double s) { prm_.step_size = s; }
  void setResolution(float r) { prm_.resolution = r; }
  double getTransformationProbability() const { return trans_probability_; }
  using RegistrationBase::align;
  void align(PointCloud& out, const Matrix4f& guess) override {
    if (!source_ || !target_) throw Error("NDT: source/target not set");
    prm_.max_iterations = max_iterations_;
    prm_.transformation_epsilon = transformation_epsilon_;
    DeviceCloud s(*source_), t(*target_), a(1, static_cast<int>(source_->size()));
    rspcl_ndt_result r;
    Device::get().check(rspcl_ndt_align(Device::get().ctx(), s.handle(), t.handle(), &prm_, guess.data(), &r, a.handle()), "ndt_align");
    converged_ = r.converged != 0;
    std::memcpy(final_.m, r.T, sizeof(r.T));
    trans_probability_ = r.trans_probability;
    a.download(out);
    out.width = source_->width, out.height = source_->height, out.is_dense = source_->is_dense;
  }

 private:
  rspcl_ndt_params prm_;
  double trans_probability_ = 0;
};

}  // namespace rspcl

// ------------------------------------------------------------------ blur_filter.hpp:16-37
class BlurFilter {
 public:
  void filter(rgb_point_cloud_pointer input_cloud) {  // centre 3/5 crop, in place
    if (!input_cloud->isOrganized()) throw rspcl::Error("BlurFilter needs an organized cloud");
    rspcl::DeviceCloud d(*input_cloud);
    const int ow = input_cloud->width * 3 / 5, oh = input_cloud->height * 3 / 5;
    rspcl::DeviceCloud o(1, ow * oh);
    rspcl::Device::get().check(rspcl_crop35(rspcl::Device::get().ctx(), d.handle(), o.handle()), "crop35");
    o.download(*input_cloud);
    input_cloud->width = ow;
    input_cloud->height = oh;
  }
};

namespace rspcl {
namespace io {
inline void savePCDFileBinary(const std::string& path, const PointCloud& cloud);  // defined below
}
}  // namespace rspcl

// ------------------------------------------------------------------ types.hpp:14-44
class RegistrationScheme {
 public:
  virtual ~RegistrationScheme() {}
  virtual rgb_point_cloud_pointer registration(std::vector<rgb_point_cloud_pointer>& clouds) = 0;
};

class TwoPhaseRegistrationScheme : public RegistrationScheme {
 public:
  typedef std::vector<std::pair<rgb_point_cloud_pointer, rgb_point_cloud_pointer>> FeaturePairs;
  virtual rgb_point_cloud_pointer extract_features(rgb_point_cloud_pointer cloud) = 0;
  virtual rgb_point_cloud_pointer global_registration(FeaturePairs& clouds) = 0;
  rgb_point_cloud_pointer registration(std::vector<rgb_point_cloud_pointer>& clouds) override {
    FeaturePairs pairs;
    pairs.reserve(clouds.size());
    for (auto& c : clouds) pairs.emplace_back(extract_features(c), c);  // phase 1: features of every cloud
    return global_registration(pairs);                                  // phase 2
  }
};

// Shared body of the two edge-based schemes (icp:26-130 and ndt:23-117 differ only in the coarse stage and in how
// the IMU angles enter the initial guess).
class EdgeBasedRegistrationBase : public TwoPhaseRegistrationScheme {
 public:
  rgb_point_cloud_pointer extract_features(rgb_point_cloud_pointer cloud) override { return rspcl::extract_edge_features(cloud); }

  // types.hpp:30-43.  By default the whole sweep runs device-resident through rspcl_register_sequence (one upload of the
  // frames, one download of the merged cloud, the edge target accumulating in HBM); device_resident = false takes the
  // PCL-shaped path below call by call (each PCL-style object uploads its inputs and downloads its output), which is
  // what a maintainer gets when only the PCL classes inside the reference's own loop are swapped (INTEGRATION.md level 1).
  rgb_point_cloud_pointer registration(std::vector<rgb_point_cloud_pointer>& clouds) override {
    if (!device_resident || !dump_pairs_prefix.empty() || clouds.empty() || !clouds[0]->isOrganized())
      return TwoPhaseRegistrationScheme::registration(clouds);
    for (auto& c : clouds)
      if (c->width != clouds[0]->width || c->height != clouds[0]->height) return TwoPhaseRegistrationScheme::registration(clouds);
    if (use_imu) assert(clouds.size() == thetas.size());
    const std::size_t n = clouds.size();
    const int w = static_cast<int>(clouds[0]->width), h = static_cast<int>(clouds[0]->height);
    std::vector<float> guess(n * 16, 0.f);
    float acc_rads = 0.f;
    for (std::size_t k = 1; k < n; ++k) {
      rspcl::Matrix4f g;
      if (use_imu) {
        const rs_float3 zero = thetas[0] * -1.0f;
        thetas[k].add(zero.x, zero.y, zero.z);
        g = imu_guess(thetas[k]);
      } else {
        acc_rads += rads;
        g = rspcl::Matrix4f::AngleAxis(acc_rads, 1);
      }
      std::memcpy(&guess[k * 16], g.data(), 64);
    }
    rspcl::Device& dev = rspcl::Device::get();
    rspcl::DeviceCloud frames(static_cast<int>(n), w * h), merged(1, static_cast<int>(n) * w * h);
    frames.upload_batch(clouds);
    std::vector<rspcl_pair_result> res(n);
    rspcl_icp_params ip;
    rspcl_icp_reference_params(&ip);  // icp:42-45,49-52 / ndt:47-50
    rspcl_ndt_params np;
    rspcl_ndt_reference_params(&np);  // ndt:39-43
    const float leaf[3] = {0.01f, 0.01f, 0.01f};
    dev.check(rspcl_register_sequence(dev.ctx(), frames.handle(), coarse_kind(), &ip, &np, leaf, 40.f, 100.f, guess.data(), res.data(),
                                      merged.handle(), nullptr),
              "register_sequence");
    transforms.assign(n, rspcl::Matrix4f::Identity());
    coarse_transforms = fine_transforms = transforms;
    accepted.assign(n, 0);
    for (std::size_t k = 0; k < n; ++k) {
      std::memcpy(coarse_transforms[k].data(), res[k].T_coarse, 64);
      std::memcpy(fine_transforms[k].data(), res[k].T_fine, 64);
      transforms[k] = fine_transforms[k] * coarse_transforms[k];
      accepted[k] = res[k].converged;
    }
    rgb_point_cloud_pointer global(new rgb_point_cloud);
    merged.download(*global);
    return global;
  }
  bool device_resident = true;

  rgb_point_cloud_pointer global_registration(FeaturePairs& clouds) override {
    if (use_imu) assert(clouds.size() == thetas.size());
    rspcl::ApproximateVoxelGrid voxel;
    voxel.setLeafSize(0.01f, 0.01f, 0.01f);
    rspcl::IterativeClosestPoint fine;
    configure_icp(fine);
    prepare_coarse();

    rgb_point_cloud_pointer target = clouds[0].first;  // grows as frames are accepted
    rgb_point_cloud_pointer global(new rgb_point_cloud(*clouds[0].second));
    global->height = 1, global->width = static_cast<std::uint32_t>(global->size());
    voxel.setInputCloud(target);
    voxel.filter(*target);

    transforms.assign(clouds.size(), rspcl::Matrix4f::Identity());
    coarse_transforms = fine_transforms = transforms;
    accepted.assign(clouds.size(), 0);
    accepted[0] = 1;
    float acc_rads = 0.f;
    for (std::size_t k = 1; k < clouds.size(); ++k) {
      rgb_point_cloud_pointer down(new rgb_point_cloud), coarse_out(new rgb_point_cloud), fine_out(new rgb_point_cloud);
      voxel.setInputCloud(clouds[k].first);
      voxel.filter(*down);
      rspcl::Matrix4f guess;
      if (use_imu) {
        const rs_float3 zero = thetas[0] * -1.0f;  // angles relative to the first frame (rewritten in place, icp:83-84)
        thetas[k].add(zero.x, zero.y, zero.z);
        guess = imu_guess(thetas[k]);
      } else {
        acc_rads += rads;
        guess = rspcl::Matrix4f::AngleAxis(acc_rads, 1);
      }
      if (!dump_pairs_prefix.empty()) {  // parity aid: the exact inputs of this frame's stages (like icp:68's edge dumps)
        rspcl::io::savePCDFileBinary(dump_pairs_prefix + "-src-" + std::to_string(k) + ".pcd", *down);
        rspcl::io::savePCDFileBinary(dump_pairs_prefix + "-tgt-" + std::to_string(k) + ".pcd", *target);
      }
      const rspcl::Matrix4f T_coarse = run_coarse(down, target, guess, *coarse_out);
      if (!dump_pairs_prefix.empty())
        rspcl::io::savePCDFileBinary(dump_pairs_prefix + "-coarse-" + std::to_string(k) + ".pcd", *coarse_out);
      fine.setInputSource(coarse_out);
      fine.setInputTarget(target);
      fine.align(*fine_out);
      coarse_transforms[k] = T_coarse;
      fine_transforms[k] = fine.getFinalTransformation();
      transforms[k] = fine.getFinalTransformation() * T_coarse;
      if (!fine.hasConverged()) continue;  // failed frames are skipped silently (icp:113-123)
      accepted[k] = 1;
      rgb_point_cloud moved;
      rspcl::transformPointCloud(*clouds[k].second, moved, T_coarse);
      rspcl::transformPointCloud(moved, moved, fine.getFinalTransformation());
      *target = *fine_out + *target;  // new points first (icp:119)
      *global += moved;               // new points last (icp:120)
    }
    return global;
  }

  std::vector<rspcl::Matrix4f> transforms;  // per frame: T_fine * T_coarse (identity for frame 0)
  std::vector<rspcl::Matrix4f> coarse_transforms, fine_transforms;
  std::vector<int> accepted;
  std::string dump_pairs_prefix;  // non-empty: write <prefix>-{src,tgt,coarse}-<k>.pcd, the inputs of every frame's stages

 protected:
  static void configure_icp(rspcl::IterativeClosestPoint& icp) {  // icp:42-45,49-52 / ndt:47-50
    icp.setMaximumIterations(100);
    icp.setMaxCorrespondenceDistance(0.01);
    icp.setTransformationEpsilon(1);
    icp.setEuclideanFitnessEpsilon(1000);
  }
  virtual int coarse_kind() const = 0;  // RSPCL_COARSE_ICP / RSPCL_COARSE_NDT
  virtual void prepare_coarse() = 0;
  virtual rspcl::Matrix4f imu_guess(const rs_float3& theta) const = 0;
  virtual rspcl::Matrix4f run_coarse(rgb_point_cloud_pointer src, rgb_point_cloud_pointer tgt, const rspcl::Matrix4f& guess,
                                     rgb_point_cloud& out) = 0;
  bool use_imu = false;
  std::vector<rs_float3> thetas;
  float rads = -0.523599f;
};

class ICPEdgeBasedRegistration : public EdgeBasedRegistrationBase {
 public:
  ICPEdgeBasedRegistration() {}
  explicit ICPEdgeBasedRegistration(std::vector<rs_float3>& input_thetas) { thetas = input_thetas, use_imu = true; }
  explicit ICPEdgeBasedRegistration(float usr_def_rads) { rads = usr_def_rads; }

 protected:
  int coarse_kind() const override { return RSPCL_COARSE_ICP; }
  void prepare_coarse() override {
    coarse_.reset(new rspcl::IterativeClosestPoint);
    configure_icp(*coarse_);
  }
  rspcl::Matrix4f imu_guess(const rs_float3& t) const override {  // icp:86-92: R_z(theta.x) R_y(-theta.y) R_x(theta.z)
    return rspcl::Matrix4f::AngleAxis(t.x, 2) * rspcl::Matrix4f::AngleAxis(-t.y, 1) * rspcl::Matrix4f::AngleAxis(t.z, 0);
  }
  rspcl::Matrix4f run_coarse(rgb_point_cloud_pointer src, rgb_point_cloud_pointer tgt, const rspcl::Matrix4f& guess,
                             rgb_point_cloud& out) override {
    coarse_->setInputSource(src);
    coarse_->setInputTarget(tgt);
    coarse_->align(out, guess);
    return coarse_->getFinalTransformation();
  }

 private:
  std::unique_ptr<rspcl::IterativeClosestPoint> coarse_;
};

class NDTEdgeBasedRegistration : public EdgeBasedRegistrationBase {
 public:
  NDTEdgeBasedRegistration() {}
  explicit NDTEdgeBasedRegistration(std::vector<rs_float3>& input_thetas) { thetas = input_thetas, use_imu = true; }
  explicit NDTEdgeBasedRegistration(float usr_def_rads) { rads = usr_def_rads; }

 protected:
  int coarse_kind() const override { return RSPCL_COARSE_NDT; }
  void prepare_coarse() override {
    ndt_.reset(new rspcl::NormalDistributionsTransform);
    ndt_->setTransformationEpsilon(0.01);  // ndt:39-43
    ndt_->setStepSize(0.1);
    ndt_->setResolution(1.0f);
    ndt_->setMaximumIterations(50);
  }
  rspcl::Matrix4f imu_guess(const rs_float3& t) const override { return rspcl::Matrix4f::AngleAxis(-t.y, 1); }  // ndt:79
  rspcl::Matrix4f run_coarse(rgb_point_cloud_pointer src, rgb_point_cloud_pointer tgt, const rspcl::Matrix4f& guess,
                             rgb_point_cloud& out) override {
    ndt_->setInputSource(src);
    ndt_->setInputTarget(tgt);
    ndt_->align(out, guess);
    return ndt_->getFinalTransformation();
  }

 private:
  std::unique_ptr<rspcl::NormalDistributionsTransform> ndt_;
};

// incremental_icp.hpp:33-69: full clouds, source voxel-filtered with the DEFAULT leaf, raw growing target
class IncrementalICP : public RegistrationScheme {
 public:
  rgb_point_cloud_pointer registration(std::vector<rgb_point_cloud_pointer>& clouds) override {
    rspcl::ApproximateVoxelGrid voxel;  // leaf never set by the reference
    rspcl::IterativeClosestPoint icp;
    icp.setMaximumIterations(100);
    icp.setMaxCorrespondenceDistance(0.01);
    icp.setTransformationEpsilon(1);
    icp.setEuclideanFitnessEpsilon(1000);
    rgb_point_cloud_pointer target = clouds[0];
    transforms.assign(clouds.size(), rspcl::Matrix4f::Identity());
    for (std::size_t k = 1; k < clouds.size(); ++k) {
      rgb_point_cloud_pointer down(new rgb_point_cloud);
      rgb_point_cloud aligned;
      voxel.setInputCloud(clouds[k]);
      voxel.filter(*down);
      icp.setInputSource(down);
      icp.setInputTarget(target);
      icp.align(aligned);
      transforms[k] = icp.getFinalTransformation();
      if (!icp.hasConverged()) continue;
      rgb_point_cloud moved;
      rspcl::transformPointCloud(*clouds[k], moved, icp.getFinalTransformation());
      *target += moved;
    }
    return target;
  }
  std::vector<rspcl::Matrix4f> transforms;
};

// ------------------------------------------------------------------ binary .pcd I/O (x y z rgb rows, main.cpp:79-87)
namespace rspcl {
namespace io {
inline void loadPCDFile(const std::string& path, PointCloud& cloud) {
  // pcl::io::loadPCDFile for the field layouts the reference reads and writes (main.cpp:81,103; examples/visualizer/*.pcd):
  // x y z [rgb|rgba], every field 4 bytes, COUNT 1.  TYPE decides how the colour is parsed: F = the float whose BITS are
  // the packed bgra (how PCL writes PointXYZRGB), U / I = the packed integer itself (exampleTemp.pcd: "TYPE F F F U").
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error("cannot open " + path);
  std::string line, data;
  std::size_t n = 0;
  std::uint32_t w = 0, h = 1;
  std::vector<std::string> fields, types;
  std::vector<int> sizes, counts;
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    std::string key, t;
    ss >> key;
    if (key == "FIELDS") {
      while (ss >> t) fields.push_back(t);
    } else if (key == "SIZE") {
      while (ss >> t) sizes.push_back(std::atoi(t.c_str()));
    } else if (key == "TYPE") {
      while (ss >> t) types.push_back(t);
    } else if (key == "COUNT") {
      while (ss >> t) counts.push_back(std::atoi(t.c_str()));
    } else if (key == "WIDTH") ss >> w;
    else if (key == "HEIGHT") ss >> h;
    else if (key == "POINTS") ss >> n;
    else if (key == "DATA") {
      ss >> data;
      break;
    }
  }
  if (fields.size() < 3 || fields[0] != "x" || fields[1] != "y" || fields[2] != "z")
    throw Error("unsupported PCD fields in " + path + " (need x y z [rgb|rgba] first)");
  if (types.empty()) types.assign(fields.size(), "F");
  if (sizes.empty()) sizes.assign(fields.size(), 4);
  if (counts.empty()) counts.assign(fields.size(), 1);
  if (types.size() != fields.size() || sizes.size() != fields.size() || counts.size() != fields.size())
    throw Error("inconsistent PCD header in " + path);
  for (std::size_t k = 0; k < fields.size(); ++k) {
    if (counts[k] != 1) throw Error("unsupported PCD COUNT (must be 1) in " + path);
    if (k < 4 && sizes[k] != 4) throw Error("unsupported PCD SIZE (x y z rgb must be 4 bytes) in " + path);
    if (k < 3 && types[k] != "F") throw Error("unsupported PCD TYPE (x y z must be F) in " + path);
  }
  const bool has_rgb = fields.size() >= 4 && (fields[3] == "rgb" || fields[3] == "rgba");
  const bool rgb_is_int = has_rgb && types[3] != "F";
  std::size_t row_bytes = 0;
  for (int sz : sizes) row_bytes += static_cast<std::size_t>(sz);
  cloud.points.assign(n, PointXYZRGB());
  cloud.width = w;
  cloud.height = h;
  if (data == "binary") {
    std::vector<char> row(row_bytes);
    for (std::size_t i = 0; i < n; ++i) {
      f.read(row.data(), static_cast<std::streamsize>(row_bytes));
      if (!f) throw Error("truncated PCD data in " + path);
      std::memcpy(&cloud.points[i].x, row.data(), 4);
      std::memcpy(&cloud.points[i].y, row.data() + 4, 4);
      std::memcpy(&cloud.points[i].z, row.data() + 8, 4);
      if (has_rgb) std::memcpy(&cloud.points[i].rgba, row.data() + 12, 4);  // float bits or packed integer: the same 4 bytes
    }
  } else if (data == "ascii") {
    for (std::size_t i = 0; i < n; ++i) {
      for (std::size_t k = 0; k < fields.size(); ++k) {
        std::string tok;
        if (!(f >> tok)) throw Error("truncated PCD data in " + path);
        if (k < 3) {
          const float v = std::strtof(tok.c_str(), nullptr);
          (k == 0 ? cloud.points[i].x : k == 1 ? cloud.points[i].y : cloud.points[i].z) = v;
        } else if (k == 3 && has_rgb) {
          if (rgb_is_int) {
            cloud.points[i].rgba = static_cast<std::uint32_t>(std::strtoull(tok.c_str(), nullptr, 10));
          } else {
            const float rgb = std::strtof(tok.c_str(), nullptr);
            std::memcpy(&cloud.points[i].rgba, &rgb, 4);
          }
        }
      }
    }
  } else {
    throw Error("unsupported PCD DATA section in " + path);
  }
}
inline void savePCDFileBinary(const std::string& path, const PointCloud& cloud) {
  std::ofstream f(path, std::ios::binary);
  if (!f) throw Error("cannot write " + path);
  const std::size_t n = cloud.size();
  const std::uint32_t w = cloud.width * cloud.height == n && cloud.width ? cloud.width : static_cast<std::uint32_t>(n);
  const std::uint32_t h = cloud.width * cloud.height == n && cloud.width ? cloud.height : 1;
  f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
    << "WIDTH " << w << "\nHEIGHT " << h << "\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
  for (const auto& p : cloud.points) {
    float row[4] = {p.x, p.y, p.z, 0.f};
    std::memcpy(&row[3], &p.rgba, 4);
    f.write(reinterpret_cast<const char*>(row), 16);
  }
}
}  // namespace io
}  // namespace rspcl

#endif
