// orc_scheme.cpp -- CPU oracle (test infrastructure only): the reference's registration scheme drivers.
//
// Restates src/types.hpp:30-43 (two-phase driver), src/icp_edge_based_registration.hpp:26-130,
// src/ndt_edge_based_registration.hpp:23-117 and src/incremental_icp.hpp:35-69 on top of the PCL
// restatements in orc_edge.cpp / orc_cloud.cpp / orc_icp.cpp / orc_ndt.cpp.  PARITY UNPINNED (see orc.h).
#include "orc.h"
#include <chrono>
#include <cmath>
#include <cfloat>
#include <cstring>
#include <vector>

namespace {
using clk = std::chrono::steady_clock;
inline double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

inline void ident(float* T) {
  for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
}
// Eigen::AngleAxisf(angle, axis).toRotationMatrix() for a unit axis, float
void angle_axis(float angle, int axis, float R[9]) {
  float c = std::cos(angle), s = std::sin(angle);
  if (axis == 0) { float M[9] = {1, 0, 0, 0, c, -s, 0, s, c}; memcpy(R, M, sizeof(M)); }
  if (axis == 1) { float M[9] = {c, 0, s, 0, 1, 0, -s, 0, c}; memcpy(R, M, sizeof(M)); }
  if (axis == 2) { float M[9] = {c, -s, 0, s, c, 0, 0, 0, 1}; memcpy(R, M, sizeof(M)); }
}
void mul3(const float* A, const float* B, float* C) {
  float R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      float s = 0;
      for (int k = 0; k < 3; ++k) s += A[r * 3 + k] * B[k * 3 + c];
      R[r * 3 + c] = s;
    }
  memcpy(C, R, sizeof(R));
}
void rot_to_mat4(const float* R, float* T) {
  ident(T);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T[c * 4 + r] = R[r * 3 + c];
}
void mat4_mul(const float* A, const float* B, float* C) {
  float R[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      float s = A[0 * 4 + r] * B[c * 4 + 0];
      s += A[1 * 4 + r] * B[c * 4 + 1];
      s += A[2 * 4 + r] * B[c * 4 + 2];
      s += A[3 * 4 + r] * B[c * 4 + 3];
      R[c * 4 + r] = s;
    }
  memcpy(C, R, sizeof(R));
}
}  // namespace

extern "C" int orc_register_pair(const OrcPoint* frame_tgt, const OrcPoint* frame_src, int w, int h, int coarse_kind,
                                 const OrcIcpParams* icp, const OrcNdtParams* ndt, const float leaf[3],
                                 const float guess[16], float T_coarse[16], float T_fine[16],
                                 OrcIcpResult* coarse_icp_res, OrcNdtResult* coarse_ndt_res, OrcIcpResult* fine_res,
                                 OrcPoint* transformed_full, OrcSchemeStats* stats) {
  const int npx = w * h;
  auto t0 = clk::now();
  std::vector<OrcPoint> e_t(npx), e_s(npx);
  int n_et = orc_extract_edges(frame_tgt, w, h, 40.f, 100.f, e_t.data(), nullptr);
  int n_es = orc_extract_edges(frame_src, w, h, 40.f, 100.f, e_s.data(), nullptr);
  auto t1 = clk::now();
  std::vector<OrcPoint> v_t(n_et > 0 ? n_et : 1), v_s(n_es > 0 ? n_es : 1);
  int n_vt = orc_approx_voxel(e_t.data(), n_et, leaf, v_t.data());
  int n_vs = orc_approx_voxel(e_s.data(), n_es, leaf, v_s.data());
  auto t2 = clk::now();
  std::vector<OrcPoint> aligned(n_vs > 0 ? n_vs : 1), icp_aligned(n_vs > 0 ? n_vs : 1);
  if (coarse_kind == ORC_COARSE_NDT) {
    OrcNdtResult r;
    orc_ndt_align(v_s.data(), n_vs, v_t.data(), n_vt, ndt, guess, &r, aligned.data());
    memcpy(T_coarse, r.T, sizeof(r.T));
    if (coarse_ndt_res) *coarse_ndt_res = r;
  } else {
    OrcIcpResult r;
    r.prev_mse = DBL_MAX;
    orc_icp_align(v_s.data(), n_vs, v_t.data(), n_vt, icp, guess, &r, aligned.data(), nullptr);
    memcpy(T_coarse, r.T, sizeof(r.T));
    if (coarse_icp_res) *coarse_icp_res = r;
  }
  auto t3 = clk::now();
  OrcIcpResult fr;
  fr.prev_mse = DBL_MAX;
  orc_icp_align(aligned.data(), n_vs, v_t.data(), n_vt, icp, nullptr, &fr, icp_aligned.data(), nullptr);
  memcpy(T_fine, fr.T, sizeof(fr.T));
  if (fine_res) *fine_res = fr;
  auto t4 = clk::now();
  if (fr.converged && transformed_full) {
    orc_transform(frame_src, npx, T_coarse, transformed_full);
    orc_transform(transformed_full, npx, T_fine, transformed_full);
  }
  auto t5 = clk::now();
  if (stats) {
    stats->n_frames = 2;
    stats->n_accepted = fr.converged;
    stats->t_edges = secs(t0, t1);
    stats->t_voxel = secs(t1, t2);
    stats->t_coarse = secs(t2, t3);
    stats->t_fine = secs(t3, t4);
    stats->t_transform = secs(t4, t5);
    stats->t_total = secs(t0, t5);
  }
  return fr.converged;
}

extern "C" int orc_scheme_edge(const OrcPoint* frames, int n, int w, int h, int coarse_kind, int use_imu, float rads,
                               float* thetas, const OrcIcpParams* icp, const OrcNdtParams* ndt, const float leaf[3],
                               OrcPoint* out_global, float* T_out, int32_t* accepted, OrcSchemeStats* stats) {
  const int npx = w * h;
  OrcSchemeStats st;
  memset(&st, 0, sizeof(st));
  st.n_frames = n;
  auto T0 = clk::now();
  // Phase 1 (types.hpp:34-38): per-cloud edge features
  std::vector<std::vector<OrcPoint>> feats(n);
  for (int k = 0; k < n; ++k) {
    feats[k].resize(npx);
    int ne = orc_extract_edges(frames + size_t(k) * npx, w, h, 40.f, 100.f, feats[k].data(), nullptr);
    feats[k].resize(ne);
  }
  auto T1 = clk::now();
  st.t_edges = secs(T0, T1);
  // Phase 2 (icp:26-130 / ndt:23-117)
  float acc_rads = 0.f;
  std::vector<OrcPoint> target = feats[0];                      // target_cloud = clouds[0].first
  size_t n_global = 0;
  memcpy(out_global, frames, sizeof(OrcPoint) * npx);           // global = global + clouds[0].second
  n_global = npx;
  {
    auto a = clk::now();
    std::vector<OrcPoint> tmp(target.size() ? target.size() : 1);
    int m = orc_approx_voxel(target.data(), int(target.size()), leaf, tmp.data());  // icp:59-60 in place
    tmp.resize(m);
    target.swap(tmp);
    st.t_voxel += secs(a, clk::now());
  }
  OrcIcpResult coarse_state, fine_state;  // persistent convergence-criteria state of the two ICP objects
  coarse_state.prev_mse = DBL_MAX;
  fine_state.prev_mse = DBL_MAX;
  if (accepted) accepted[0] = 1;
  if (T_out) ident(T_out);
  for (int k = 1; k < n; ++k) {
    auto a = clk::now();
    std::vector<OrcPoint> down(feats[k].size() ? feats[k].size() : 1);
    int m = orc_approx_voxel(feats[k].data(), int(feats[k].size()), leaf, down.data());  // icp:75-76
    down.resize(m);
    auto b = clk::now();
    st.t_voxel += secs(a, b);
    float guess[16];
    if (use_imu) {
      // icp:83-84: thetas[k] += -thetas[0]  (thetas[0] itself is never zeroed: k starts at 1)
      float ax = thetas[0] * -1.0f, ay = thetas[1] * -1.0f, az = thetas[2] * -1.0f;
      thetas[3 * k + 0] += ax;
      thetas[3 * k + 1] += ay;
      thetas[3 * k + 2] += az;
      float R[9];
      if (coarse_kind == ORC_COARSE_ICP) {
        // icp:86-92: T(0) * AngleAxis(theta.x, Z) * AngleAxis(-theta.y, Y) * AngleAxis(theta.z, X)
        float A[9], B[9], C[9];
        angle_axis(thetas[3 * k + 0], 2, A);
        angle_axis(-thetas[3 * k + 1], 1, B);
        angle_axis(thetas[3 * k + 2], 0, C);
        mul3(A, B, R);
        mul3(R, C, R);
      } else {
        angle_axis(-thetas[3 * k + 1], 1, R);  // ndt:79
      }
      rot_to_mat4(R, guess);
    } else {
      acc_rads += rads;  // icp:98, ndt:86
      float R[9];
      angle_axis(acc_rads, 1, R);
      rot_to_mat4(R, guess);
    }
    std::vector<OrcPoint> aligned(m ? m : 1), icp_aligned(m ? m : 1);
    float T_coarse[16];
    if (coarse_kind == ORC_COARSE_NDT) {
      OrcNdtResult r;
      orc_ndt_align(down.data(), m, target.data(), int(target.size()), ndt, guess, &r, aligned.data());
      memcpy(T_coarse, r.T, sizeof(r.T));
    } else {
      orc_icp_align(down.data(), m, target.data(), int(target.size()), icp, guess, &coarse_state, aligned.data(), nullptr);
      memcpy(T_coarse, coarse_state.T, sizeof(T_coarse));
    }
    auto c = clk::now();
    st.t_coarse += secs(b, c);
    orc_icp_align(aligned.data(), m, target.data(), int(target.size()), icp, nullptr, &fine_state, icp_aligned.data(), nullptr);
    auto d = clk::now();
    st.t_fine += secs(c, d);
    if (accepted) accepted[k] = fine_state.converged;
    if (T_out) mat4_mul(fine_state.T, T_coarse, T_out + 16 * k);
    if (fine_state.converged) {
      OrcPoint* dst = out_global + n_global;
      orc_transform(frames + size_t(k) * npx, npx, T_coarse, dst);  // icp:116
      orc_transform(dst, npx, fine_state.T, dst);                   // icp:117 (in place)
      n_global += npx;                                              // icp:120
      std::vector<OrcPoint> nt(icp_aligned.begin(), icp_aligned.begin() + m);  // icp:119 new points first
      nt.insert(nt.end(), target.begin(), target.end());
      target.swap(nt);
      st.n_accepted++;
    }
    st.t_transform += secs(d, clk::now());
  }
  st.t_total = secs(T0, clk::now());
  if (stats) *stats = st;
  return int(n_global);
}

extern "C" int orc_scheme_incremental(const OrcPoint* frames, int n, int npts, const OrcIcpParams* icp,
                                      const float leaf[3], OrcPoint* out_target, float* T_out, int32_t* accepted) {
  // incremental_icp.hpp:35-69: target = clouds[0] (raw), source voxel-filtered, target += transformed full cloud
  size_t nt = npts;
  memcpy(out_target, frames, sizeof(OrcPoint) * npts);
  OrcIcpResult state;
  state.prev_mse = DBL_MAX;
  if (accepted) accepted[0] = 1;
  if (T_out) ident(T_out);
  for (int k = 1; k < n; ++k) {
    const OrcPoint* cloud = frames + size_t(k) * npts;
    std::vector<OrcPoint> down(npts);
    int m = orc_approx_voxel(cloud, npts, leaf, down.data());
    orc_icp_align(down.data(), m, out_target, int(nt), icp, nullptr, &state, nullptr, nullptr);
    if (accepted) accepted[k] = state.converged;
    if (T_out) memcpy(T_out + 16 * k, state.T, sizeof(state.T));
    if (state.converged) {
      orc_transform(cloud, npts, state.T, out_target + nt);
      nt += npts;
    }
  }
  return int(nt);
}
