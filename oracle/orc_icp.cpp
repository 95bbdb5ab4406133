// orc_icp.cpp -- CPU oracle (test infrastructure only): pcl::IterativeClosestPoint as the reference drives it.
//
// Reference call sites: icp:35,41-52,78-79,95,104,108-113; ndt:32,47-50,96-101; incr:37,46-49,57-61.
// PCL 1.9.1 sources followed (not vendored -- PARITY UNPINNED, see orc.h):
//   registration/impl/registration.hpp           Registration::align, getFitnessScore
//   registration/impl/icp.hpp                    computeTransformation, transformCloud
//   registration/impl/correspondence_estimation.hpp  determineCorrespondences
//   registration/impl/transformation_estimation_svd.hpp + common/impl/eigen.hpp (pcl::umeyama)
//   registration/impl/default_convergence_criteria.hpp  hasConverged
#include "orc.h"
#include "orc_linalg.h"
#include <cmath>
#include <cfloat>
#include <cstring>
#include <vector>

struct OrcKd;
OrcKd* orc_kd_build(const OrcPoint* tgt, int nt);
void orc_kd_free(OrcKd* k);
void orc_kd_query(const OrcKd* k, const OrcPoint* q, int* idx, float* d2);

extern "C" void orc_icp_default_params(OrcIcpParams* p) {
  p->max_iterations = 10;
  p->max_corr_dist = std::sqrt(DBL_MAX);
  p->transformation_epsilon = 0.0;
  p->euclidean_fitness_epsilon = -DBL_MAX;
  p->mse_threshold_absolute = 1e-12;
  p->min_correspondences = 3;
  p->umeyama_float = 0;
}

extern "C" void orc_icp_reference_params(OrcIcpParams* p) {
  orc_icp_default_params(p);
  p->max_iterations = 100;             // icp:42
  p->max_corr_dist = 0.01;             // icp:43
  p->transformation_epsilon = 1;       // icp:44
  p->euclidean_fitness_epsilon = 1000; // icp:45
}

namespace {

inline void mat4_identity(float* T) {
  for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
}
inline bool mat4_is_identity(const float* T) {
  for (int i = 0; i < 16; ++i)
    if (T[i] != ((i % 5 == 0) ? 1.f : 0.f)) return false;
  return true;
}
// C = A * B, column-major float, each entry ((a0*b0 + a1*b1) + a2*b2) + a3*b3
inline void mat4_mul(const float* A, const float* B, float* C) {
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

template <typename S>
void umeyama_impl(const float* src, const float* tgt, int n, float T[16]) {
  // pcl::umeyama(src, dst, with_scaling=false): means, demeaned cross-covariance sigma = dst_dm * src_dm^T / n,
  // JacobiSVD, S = diag(1,1,+-1) by det(U)*det(V), R = U S V^T, t = mu_dst - R mu_src.
  S ms[3] = {0, 0, 0}, mt[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) {
      ms[a] += S(src[3 * i + a]);
      mt[a] += S(tgt[3 * i + a]);
    }
  const S one_over_n = S(1) / S(n);
  for (int a = 0; a < 3; ++a) {
    ms[a] *= one_over_n;
    mt[a] *= one_over_n;
  }
  S sigma[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    S ds[3], dt[3];
    for (int a = 0; a < 3; ++a) {
      ds[a] = S(src[3 * i + a]) - ms[a];
      dt[a] = S(tgt[3 * i + a]) - mt[a];
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) sigma[r * 3 + c] += dt[r] * ds[c];
  }
  for (int k = 0; k < 9; ++k) sigma[k] *= one_over_n;
  S U[9], sv[3], V[9];
  orc::svd3<S>(sigma, U, sv, V);
  S d = orc::det3(U) * orc::det3(V);
  S Sd[3] = {1, 1, d < 0 ? S(-1) : S(1)};
  S R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      S acc = 0;
      for (int k = 0; k < 3; ++k) acc += U[r * 3 + k] * Sd[k] * V[c * 3 + k];
      R[r * 3 + c] = acc;
    }
  mat4_identity(T);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T[c * 4 + r] = float(R[r * 3 + c]);
    S t = mt[r] - (R[r * 3 + 0] * ms[0] + R[r * 3 + 1] * ms[1] + R[r * 3 + 2] * ms[2]);
    T[12 + r] = float(t);
  }
}

}  // namespace

extern "C" void orc_umeyama(const float* src_xyz, const float* tgt_xyz, int n, int use_float, float T[16]) {
  if (use_float)
    umeyama_impl<float>(src_xyz, tgt_xyz, n, T);
  else
    umeyama_impl<double>(src_xyz, tgt_xyz, n, T);
}

static void icp_align_impl(const OrcPoint* src, int ns, const OrcPoint* tgt, int nt, const OrcIcpParams* prm,
                           const float guess[16], OrcIcpResult* res, OrcPoint* aligned, int32_t* first_corr, int n_dump) {
  // Registration::align -> initCompute (kd-tree over target) -> computeTransformation
  // first_corr: n_dump blocks of ns entries, block k = correspondences (target index or -1) of iteration k
  OrcKd* tree = orc_kd_build(tgt, nt);
  std::vector<OrcPoint> work(src, src + ns);  // input_transformed
  float final_T[16], T[16];
  float ident[16];
  mat4_identity(ident);
  const float* g = guess ? guess : ident;
  memcpy(final_T, g, sizeof(final_T));
  if (!mat4_is_identity(g)) orc_transform(src, ns, g, work.data());
  mat4_identity(T);

  const double max_dist_sqr = prm->max_corr_dist * prm->max_corr_dist;
  const double rot_thr = 1.0 - prm->transformation_epsilon;  // transformation_rotation_epsilon_ unset (0)
  const double trans_thr = prm->transformation_epsilon;
  double prev_mse = res->prev_mse;
  int iterations = 0, state = ORC_CONV_NOT_CONVERGED;
  bool converged = false;
  int n_corr = 0;
  double cur_mse = 0.0, last_mse = 0.0;
  std::vector<float> cs, ct, cd;
  if (first_corr)
    for (long long i = 0; i < (long long)ns * n_dump; ++i) first_corr[i] = -1;

  do {
    cs.clear();
    ct.clear();
    cd.clear();
    for (int i = 0; i < ns; ++i) {
      int j;
      float d;
      orc_kd_query(tree, &work[i], &j, &d);
      if (j < 0) continue;
      if (double(d) > max_dist_sqr) continue;
      if (iterations < n_dump && first_corr) first_corr[(size_t)iterations * ns + i] = j;
      cs.push_back(work[i].x);
      cs.push_back(work[i].y);
      cs.push_back(work[i].z);
      ct.push_back(tgt[j].x);
      ct.push_back(tgt[j].y);
      ct.push_back(tgt[j].z);
      cd.push_back(d);
    }
    n_corr = int(cd.size());
    if (n_corr < prm->min_correspondences) {
      state = ORC_CONV_NO_CORRESPONDENCES;
      converged = false;
      break;
    }
    orc_umeyama(cs.data(), ct.data(), n_corr, prm->umeyama_float, T);
    orc_transform(work.data(), ns, T, work.data());
    mat4_mul(T, final_T, final_T);
    ++iterations;
    {  // reported for parity checks: mean squared distance of this iteration's correspondences
      double m = 0;
      for (float d : cd) m += d;
      last_mse = m / double(n_corr);
    }

    // DefaultConvergenceCriteria::hasConverged
    state = ORC_CONV_NOT_CONVERGED;
    if (iterations >= prm->max_iterations) {
      state = ORC_CONV_ITERATIONS;
      converged = true;
      // (cur_mse intentionally not updated: PCL returns before calculateMSE)
      break;
    }
    double cos_angle = 0.5 * (T[0] + T[5] + T[10] - 1);
    double translation_sqr = T[12] * T[12] + T[13] * T[13] + T[14] * T[14];
    if (cos_angle >= rot_thr && translation_sqr <= trans_thr) {
      state = ORC_CONV_TRANSFORM;
      converged = true;
      break;
    }
    double mse = 0;
    for (float d : cd) mse += d;
    mse /= double(n_corr);
    cur_mse = mse;
    if (std::fabs(cur_mse - prev_mse) < prm->mse_threshold_absolute) {
      state = ORC_CONV_ABS_MSE;
      converged = true;
      break;
    }
    if (std::fabs(cur_mse - prev_mse) / prev_mse < prm->euclidean_fitness_epsilon) {
      state = ORC_CONV_REL_MSE;
      converged = true;
      break;
    }
    prev_mse = cur_mse;
  } while (!converged);

  memcpy(res->T, final_T, sizeof(final_T));
  res->converged = converged ? 1 : 0;
  res->state = state;
  res->iterations = iterations;
  res->n_corr = n_corr;
  res->mse = last_mse;
  res->prev_mse = prev_mse;
  if (aligned) orc_transform(src, ns, final_T, aligned);
  orc_kd_free(tree);
}

extern "C" void orc_icp_align(const OrcPoint* src, int ns, const OrcPoint* tgt, int nt, const OrcIcpParams* prm,
                              const float guess[16], OrcIcpResult* res, OrcPoint* aligned, int32_t* first_corr) {
  icp_align_impl(src, ns, tgt, nt, prm, guess, res, aligned, first_corr, 1);
}

extern "C" void orc_icp_align_dump(const OrcPoint* src, int ns, const OrcPoint* tgt, int nt, const OrcIcpParams* prm,
                                   const float guess[16], OrcIcpResult* res, int n_dump, int32_t* corr) {
  icp_align_impl(src, ns, tgt, nt, prm, guess, res, nullptr, corr, n_dump);
}

extern "C" double orc_fitness(const OrcPoint* src_transformed, int ns, const OrcPoint* tgt, int nt, double max_range) {
  OrcKd* tree = orc_kd_build(tgt, nt);
  double score = 0.0;
  int nr = 0;
  for (int i = 0; i < ns; ++i) {
    int j;
    float d;
    orc_kd_query(tree, &src_transformed[i], &j, &d);
    if (j < 0) continue;
    if (double(d) <= max_range) {
      score += d;
      nr++;
    }
  }
  orc_kd_free(tree);
  return nr > 0 ? score / nr : DBL_MAX;
}
