// icp.cu -- pcl::IterativeClosestPoint<PointXYZRGB,PointXYZRGB>::align for a batch of independent (source, target)
// pairs, with PCL's DefaultConvergenceCriteria evaluated on the device.
//
// Reference call sites: icp:35,41-52,78-79,95,104,108-113; ndt:32,47-50,96-101; incr:37,46-49,57-61.
// Kernels:
//   K5  k_icp_step   fused: apply the previous incremental transform to the working source (in place, float,
//                    un-fused -- the same rounding as PCL's transformCloud), look up the nearest target point in the
//                    voxel-hash grid (<= 2x2x2 cells), reject beyond max_corr_dist, and reduce
//                    {n, sum s, sum t, sum s t^T, sum d^2} (17 fp64) per CTA with warp shuffles.  No atomics: CTA
//                    partials are combined in a fixed order so results are reproducible run to run.
//   K5b k_icp_solve  one warp per pair: combine the partials in block order, Umeyama (fp64 polar decomposition by
//                    scaled Newton; Jacobi SVD fallback for degenerate / reflected covariances), final = T * final,
//                    DefaultConvergenceCriteria.
// Algorithmic bytes per source point per iteration: 16 R (source) + 16 R (matched target) = 32 B; the
// in-place update of the working cloud (16 W) and the neighbour-cell probes are overhead on top of that.
// The "unbounded" mode (max_corr_dist larger than the grid can bin, e.g. PCL's default sqrt(DBL_MAX)) uses the
// brute-force exact NN of grid.cu instead of the grid.
#include "grid.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#include <float.h>
#include <math.h>
#include <limits.h>
#include <stdlib.h>
#include <algorithm>

namespace {

constexpr int IT = 256;       // threads per CTA in k_icp_step
// k_icp_rescan is latency-bound (dependent cell scans): four CTAs per SM (64 registers, a few spills) beat two at 128
// registers by 22 % on the 50 M-point pair
#ifndef RESCAN_MINB
#define RESCAN_MINB 4
#endif
constexpr int NRED = 17;      // n, s(3), t(3), s t^T (9), sum d2

struct IcpState {  // device-resident per pair
  float final_T[16];
  float inc_T[16];
  double prev_mse;
  double mse;
  int iterations;
  int state;
  int converged;
  int done;
  int n_corr;
  int apply_inc;  // the working cloud still has to be moved by inc_T
  int pad[2];
};

struct IcpDevParams {
  int max_iterations, min_corr;
  double max_dist_sqr, rot_thr, trans_thr, mse_abs, mse_rel;
  float search_r;
};

__device__ __forceinline__ void mat4_identity(float* T) {
#pragma unroll
  for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
}

__device__ __forceinline__ double det3(const double* M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// ---- 3x3 one-sided Jacobi SVD in fp64: A = U diag(s) V^T, s descending (row-major)
__device__ void svd3(const double* A_in, double* U, double* s, double* V) {
  double B[9];
  for (int i = 0; i < 9; ++i) B[i] = A_in[i];
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 3; ++k) {
          alpha += B[k * 3 + p] * B[k * 3 + p];
          beta += B[k * 3 + q] * B[k * 3 + q];
          gamma += B[k * 3 + p] * B[k * 3 + q];
        }
        if (gamma == 0.0 || fabs(gamma) <= DBL_EPSILON * sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
        for (int k = 0; k < 3; ++k) {
          double bp = B[k * 3 + p], bq = B[k * 3 + q];
          B[k * 3 + p] = c * bp - sn * bq;
          B[k * 3 + q] = sn * bp + c * bq;
          double vp = V[k * 3 + p], vq = V[k * 3 + q];
          V[k * 3 + p] = c * vp - sn * vq;
          V[k * 3 + q] = sn * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double nrm[3];
  for (int j = 0; j < 3; ++j) nrm[j] = sqrt(B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j]);
  int o0 = 0, o1 = 1, o2 = 2, tmp;
  if (nrm[o0] < nrm[o1]) { tmp = o0; o0 = o1; o1 = tmp; }
  if (nrm[o1] < nrm[o2]) { tmp = o1; o1 = o2; o2 = tmp; }
  if (nrm[o0] < nrm[o1]) { tmp = o0; o0 = o1; o1 = tmp; }
  const int ord[3] = {o0, o1, o2};
  double Bs[9], Vs[9];
  for (int j = 0; j < 3; ++j) {
    s[j] = nrm[ord[j]];
    for (int k = 0; k < 3; ++k) {
      Bs[k * 3 + j] = B[k * 3 + ord[j]];
      Vs[k * 3 + j] = V[k * 3 + ord[j]];
    }
  }
  for (int i = 0; i < 9; ++i) V[i] = Vs[i];
  const double tiny = s[0] * DBL_EPSILON * 8.0;
  int rank = 0;
  for (int j = 0; j < 3; ++j) {
    if (s[j] > tiny && s[j] > 0.0) {
      for (int k = 0; k < 3; ++k) U[k * 3 + j] = Bs[k * 3 + j] / s[j];
      rank = j + 1;
    } else {
      break;
    }
  }
  if (rank == 0) {
    for (int i = 0; i < 9; ++i) U[i] = (i % 4 == 0) ? 1.0 : 0.0;
  } else if (rank == 1) {
    double u0[3] = {U[0], U[3], U[6]};
    int m = 0;
    if (fabs(u0[1]) < fabs(u0[m])) m = 1;
    if (fabs(u0[2]) < fabs(u0[m])) m = 2;
    double e[3] = {0, 0, 0};
    e[m] = 1;
    double u1[3] = {u0[1] * e[2] - u0[2] * e[1], u0[2] * e[0] - u0[0] * e[2], u0[0] * e[1] - u0[1] * e[0]};
    double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    for (int k = 0; k < 3; ++k) u1[k] /= n1;
    double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
    for (int k = 0; k < 3; ++k) {
      U[k * 3 + 1] = u1[k];
      U[k * 3 + 2] = u2[k];
    }
  } else if (rank == 2) {
    double u0[3] = {U[0], U[3], U[6]}, u1[3] = {U[1], U[4], U[7]};
    double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
    double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    for (int k = 0; k < 3; ++k) U[k * 3 + 2] = u2[k] / n2;
  }
}

// Orthogonal polar factor of a well-conditioned 3x3 with positive determinant by scaled Newton iteration
// (X <- (g X + X^-T / g) / 2).  For det(sigma) > 0 it equals U V^T of the SVD, i.e. Umeyama's rotation with S = I,
// at a fraction of the latency of a Jacobi SVD on one thread.  Returns false if the matrix is too close to singular
// (or a reflection is needed) and the SVD path must decide.
__device__ bool polar_rotation(const double* A, double* R) {
  double X[9];
  double fro2 = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    X[i] = A[i];
    fro2 += A[i] * A[i];
  }
  const double det0 = det3(A);
  const double scale = sqrt(fro2 / 3.0);
  if (!(det0 > 1e-7 * scale * scale * scale)) return false;
  // Any positive scaling sequence converges to the same polar factor, so the scale g only needs float accuracy
  // (rsqrtf) and is dropped once the iterate is nearly orthogonal; the loop then has no fp64 sqrt and one reciprocal.
  bool scaling = true;
#pragma unroll 1
  for (int it = 0; it < 40; ++it) {
    const double c[9] = {X[4] * X[8] - X[5] * X[7], X[5] * X[6] - X[3] * X[8], X[3] * X[7] - X[4] * X[6],
                         X[2] * X[7] - X[1] * X[8], X[0] * X[8] - X[2] * X[6], X[1] * X[6] - X[0] * X[7],
                         X[1] * X[5] - X[2] * X[4], X[2] * X[3] - X[0] * X[5], X[0] * X[4] - X[1] * X[3]};
    const double det = X[0] * c[0] + X[1] * c[1] + X[2] * c[2];
    if (!(det > 0.0)) return false;
    const double idet = __drcp_rn(det);
    double a = 0.5, bq = 0.5 * idet;  // X <- a X + bq cof(X)   (cof(X) / det = X^-T)
    if (scaling) {
      double nx = 0, ny = 0;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        nx += X[i] * X[i];
        ny += c[i] * c[i];
      }
      const float ratio = (float)(ny * idet * idet / nx);  // ||X^-T||_F^2 / ||X||_F^2
      const float g = sqrtf(sqrtf(ratio));
      if (fabsf(g - 1.0f) < 1e-2f) scaling = false;
      a = 0.5 * (double)g;
      bq = 0.5 * idet / (double)g;
    }
    double diff = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double v = a * X[i] + bq * c[i];
      diff += (v - X[i]) * (v - X[i]);
      X[i] = v;
    }
    if (!scaling && diff <= 1e-22) break;  // quadratic convergence: the NEXT update would be ~1e-22, below fp64 resolution
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = X[i];
  return true;
}

// 1/x to full fp64 accuracy from the fp32 reciprocal and two Newton steps (a handful of DFMAs instead of the ~40
// dependent instructions of an IEEE fp64 division); x must be a normal number inside the float range.
__device__ __forceinline__ double rcp_fast(double x) {
  double r = (double)__frcp_rn((float)x);
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}

// Fast path of the same polar factor.  With an orthogonal seed Y near the answer, M = Y^T A = Q H (Q = exp([k]x) the
// small rotation still missing, H symmetric positive definite) and the skew part of M satisfies
// (tr(H) I - H) k = 2 axial(skew M), a 3x3 SPD solve in fp64; R = Y (I + [k]x + [k]x^2 / 2) squares the error
// (|k| ~ 1e-6 -> ~1e-14 measured).  Seeds, cheapest first:
//   pass 0: Y = I -- an ICP increment is a tiny rotation after the first iterations, so one pass usually finishes;
//   else:   the scaled Newton iteration in fp32 (4-cycle FMAs instead of the fp64 chain), re-orthogonalised by one
//           fp64 Newton step per pass.
// The straight-line fp64 code is kept short on purpose: one thread runs it once per ICP iteration, so its cost is the
// instruction fetch, not the arithmetic.  Returns false (caller falls back to the fp64 iteration / SVD) if anything
// looks degenerate.
__device__ bool polar_rotation_fast(const double* A, double* R) {
  double fro2 = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) fro2 = fma(A[i], A[i], fro2);
  const float scale_f = sqrtf((float)fro2 * (1.0f / 3.0f));
  if (!(scale_f > 1e-30f && scale_f < 1e30f)) return false;
  const double det0 = det3(A), scale = (double)scale_f;
  // The correction passes only need tr(H) I - H positive definite (sigma_2 > |sigma_3|), i.e. they also cover the
  // near-planar / reflected covariances for which Umeyama flips the last singular direction (the stationary point
  // with that G positive definite IS the rotation maximising tr(R^T A)); only the fp32 Newton seed needs det(A) > 0.
  const bool newton_ok = det0 > 1e-7 * scale * scale * scale;
  double Y[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
  bool seeded = false;
#pragma unroll 1
  for (int pass = 0; pass < 7; ++pass) {
    if (pass > 0) {  // one fp64 Newton step: orthogonality error e -> e^2 / 2
      const double c[9] = {Y[4] * Y[8] - Y[5] * Y[7], Y[5] * Y[6] - Y[3] * Y[8], Y[3] * Y[7] - Y[4] * Y[6],
                           Y[2] * Y[7] - Y[1] * Y[8], Y[0] * Y[8] - Y[2] * Y[6], Y[1] * Y[6] - Y[0] * Y[7],
                           Y[1] * Y[5] - Y[2] * Y[4], Y[2] * Y[3] - Y[0] * Y[5], Y[0] * Y[4] - Y[1] * Y[3]};
      const double det = Y[0] * c[0] + Y[1] * c[1] + Y[2] * c[2];
      if (!(det > 0.5 && det < 2.0)) return false;
      const double h = 0.5 * rcp_fast(det);
#pragma unroll
      for (int i = 0; i < 9; ++i) Y[i] = fma(0.5, Y[i], h * c[i]);
    }
    // M = Y^T A, G = tr(H) I - H with H = sym(M), rhs = 2 axial(skew(M))
    double M[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) M[r * 3 + c] = fma(Y[0 + r], A[0 + c], fma(Y[3 + r], A[3 + c], Y[6 + r] * A[6 + c]));
    const double h01 = 0.5 * (M[1] + M[3]), h02 = 0.5 * (M[2] + M[6]), h12 = 0.5 * (M[5] + M[7]);
    const double tr = M[0] + M[4] + M[8];
    const double G[9] = {tr - M[0], -h01, -h02, -h01, tr - M[4], -h12, -h02, -h12, tr - M[8]};
    const double b[3] = {M[7] - M[5], M[2] - M[6], M[3] - M[1]};
    const double gc[6] = {G[4] * G[8] - G[5] * G[5], G[5] * G[2] - G[1] * G[8], G[1] * G[5] - G[4] * G[2],
                          G[0] * G[8] - G[2] * G[2], G[1] * G[2] - G[0] * G[5], G[0] * G[4] - G[1] * G[1]};  // adj(G)
    const double gdet = G[0] * gc[0] + G[1] * gc[1] + G[2] * gc[2];
    const float gdet_f = (float)gdet;
    // G positive definite (Sylvester) <=> H positive definite given det(A) > 0: the seed is within 90 degrees
    bool good = G[0] > 0.0 && gc[5] > 0.0 && gdet_f > 1e-30f && gdet_f < 1e30f;
    double k0 = 0, k1 = 0, k2 = 0, kk = 1.0;
    if (good) {
      const double ig = rcp_fast(gdet);
      k0 = (gc[0] * b[0] + gc[1] * b[1] + gc[2] * b[2]) * ig;
      k1 = (gc[1] * b[0] + gc[3] * b[1] + gc[4] * b[2]) * ig;
      k2 = (gc[2] * b[0] + gc[4] * b[1] + gc[5] * b[2]) * ig;
      kk = k0 * k0 + k1 * k1 + k2 * k2;
      good = kk < 1e-2;  // |k| < 0.1 rad: inside the quadratic-convergence basin
    }
    if (!good) {
      if (seeded || !newton_ok) return false;  // no usable seed: the exact path decides
      // ---- fp32 scaled Newton seed (X <- (g X + X^-T / g) / 2)
      seeded = true;
      const float inv_scale = __frcp_rn(scale_f);
      float X[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) X[i] = (float)A[i] * inv_scale;
      bool scaling = true, done = false;
#pragma unroll 1
      for (int it = 0; it < 20 && !done; ++it) {
        const float c[9] = {X[4] * X[8] - X[5] * X[7], X[5] * X[6] - X[3] * X[8], X[3] * X[7] - X[4] * X[6],
                            X[2] * X[7] - X[1] * X[8], X[0] * X[8] - X[2] * X[6], X[1] * X[6] - X[0] * X[7],
                            X[1] * X[5] - X[2] * X[4], X[2] * X[3] - X[0] * X[5], X[0] * X[4] - X[1] * X[3]};
        const float det = X[0] * c[0] + X[1] * c[1] + X[2] * c[2];
        if (!(det > 0.f)) return false;
        const float idet = __fdividef(1.0f, det);
        float a = 0.5f, bq = 0.5f * idet;
        if (scaling) {
          const float nx = (X[0] * X[0] + X[1] * X[1] + X[2] * X[2]) + (X[3] * X[3] + X[4] * X[4] + X[5] * X[5]) +
                           (X[6] * X[6] + X[7] * X[7] + X[8] * X[8]);
          const float ny = (c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) + (c[3] * c[3] + c[4] * c[4] + c[5] * c[5]) +
                           (c[6] * c[6] + c[7] * c[7] + c[8] * c[8]);
          const float g2 = sqrtf(__fdividef(ny * idet * idet, nx));  // g^2
          const float ig = rsqrtf(g2);
          if (fabsf(g2 - 1.0f) < 2e-2f) scaling = false;
          a = 0.5f * g2 * ig;  // g / 2
          bq = 0.5f * idet * ig;
        }
        float diff = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const float v = a * X[i] + bq * c[i];
          diff += (v - X[i]) * (v - X[i]);
          X[i] = v;
        }
        if (!scaling && diff <= 1e-10f) done = true;
      }
      if (!done) return false;
#pragma unroll
      for (int i = 0; i < 9; ++i) Y[i] = (double)X[i];
      continue;
    }
    // Y <- Y Q, Q = I + K + K^2 / 2, K = [k]x
    const double Q[9] = {1.0 - 0.5 * (k1 * k1 + k2 * k2), -k2 + 0.5 * k0 * k1, k1 + 0.5 * k0 * k2,
                         k2 + 0.5 * k0 * k1, 1.0 - 0.5 * (k0 * k0 + k2 * k2), -k0 + 0.5 * k1 * k2,
                         -k1 + 0.5 * k0 * k2, k0 + 0.5 * k1 * k2, 1.0 - 0.5 * (k0 * k0 + k1 * k1)};
    double Z[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) Z[r * 3 + c] = fma(Y[r * 3 + 0], Q[0 + c], fma(Y[r * 3 + 1], Q[3 + c], Y[r * 3 + 2] * Q[6 + c]));
#pragma unroll
    for (int i = 0; i < 9; ++i) Y[i] = Z[i];
    if (kk < 1e-11) {  // |k| < 3e-6: what is left after this pass is O(k^2) (measured ~1e-14 at |k| = 2e-6)
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = Y[i];
      return true;
    }
  }
  return false;
}

// Cold path of the rotation solve (scaled Newton in fp64, else Jacobi SVD with Umeyama's reflection fix), out of line and
// with its operands passed BY VALUE: the hot path of the caller keeps sigma / R in registers and its instruction stream
// short (one thread runs it once per ICP iteration; what it costs is instruction fetch and dependent latency).
struct Mat3d {
  double m[9];
};
__device__ __noinline__ Mat3d umeyama_rotation_slow(const Mat3d sg) {
  Mat3d out;
  double sigma[9], R[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) sigma[i] = sg.m[i];
  if (!polar_rotation(sigma, R)) {
    double U[9], sv[3], V[9];
    svd3(sigma, U, sv, V);
    const double d = det3(U) * det3(V);
    const double Sd[3] = {1.0, 1.0, d < 0 ? -1.0 : 1.0};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double acc = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc += U[r * 3 + k] * Sd[k] * V[c * 3 + k];
        R[r * 3 + c] = acc;
      }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) out.m[i] = R[i];
  return out;
}

// pcl::umeyama(src, dst, with_scaling=false) from raw fp64 moments: S = {n, sum s, sum t, sum s_r t_c}
__device__ void umeyama_from_moments(const double* S, float* T) {
  const double n = S[0], inv_n = rcp_fast(n);
  const double ms[3] = {S[1] * inv_n, S[2] * inv_n, S[3] * inv_n};
  const double mt[3] = {S[4] * inv_n, S[5] * inv_n, S[6] * inv_n};
  double sigma[9];  // sigma(r,c) = mean((t_r - mt_r)(s_c - ms_c))
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) sigma[r * 3 + c] = S[7 + c * 3 + r] * inv_n - mt[r] * ms[c];
  double R[9];
  if (!polar_rotation_fast(sigma, R)) {
    Mat3d in;
#pragma unroll
    for (int i = 0; i < 9; ++i) in.m[i] = sigma[i];
    const Mat3d out = umeyama_rotation_slow(in);
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = out.m[i];
  }
  mat4_identity(T);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) T[c * 4 + r] = (float)R[r * 3 + c];
    T[12 + r] = (float)(mt[r] - (R[r * 3 + 0] * ms[0] + R[r * 3 + 1] * ms[1] + R[r * 3 + 2] * ms[2]));
  }
}

__global__ void k_icp_init(IcpState* __restrict__ st, const float* __restrict__ guess, const double* __restrict__ prev_mse,
                           const unsigned char* __restrict__ skip, int n_seg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  IcpState S;
  if (guess)
    for (int i = 0; i < 16; ++i) S.final_T[i] = guess[s * 16 + i];
  else
    mat4_identity(S.final_T);
  bool ident = true;
  for (int i = 0; i < 16; ++i)
    if (S.final_T[i] != ((i % 5 == 0) ? 1.f : 0.f)) ident = false;
  // the first k_icp_step moves the source by the guess (icp.hpp: transformCloud(*input_, *input_transformed, guess))
  for (int i = 0; i < 16; ++i) S.inc_T[i] = S.final_T[i];
  S.apply_inc = ident ? 0 : 1;
  S.prev_mse = prev_mse[s];
  S.mse = 0.0;
  S.iterations = 0;
  S.state = RSPCL_CONV_NOT_CONVERGED;
  S.converged = 0;
  S.done = (skip && skip[s]) ? 1 : 0;  // pairs some other path has already finished are left alone by every kernel
  S.n_corr = 0;
  S.pad[0] = S.pad[1] = 0;
  st[s] = S;
}

// Umeyama + DefaultConvergenceCriteria for one pair, given the 17 combined sums.  Called by one whole warp: lane 0 runs
// the serial part, lanes 0-15 each produce one element of final = T * final.
__device__ void icp_solve_pair(IcpState* S, const double* sums, const IcpDevParams& prm, int* n_active, int lane) {
  if (lane == 0) {
    S->apply_inc = 0;
    const int n_corr = (int)(sums[0] + 0.5);
    S->n_corr = n_corr;
    if (n_corr < prm.min_corr) {
      // icp.hpp: "Not enough correspondences found" -> NO_CORRESPONDENCES, converged_ = false, loop exits
      S->state = RSPCL_CONV_NO_CORRESPONDENCES;
      S->converged = 0;
      S->done = 1;
      atomicSub(n_active, 1);
    } else {
      float T[16];
      umeyama_from_moments(sums, T);
#pragma unroll
      for (int i = 0; i < 16; ++i) S->inc_T[i] = T[i];
      S->apply_inc = 1;
    }
  }
  __syncwarp();
  if (!S->apply_inc) return;  // warp-uniform (read after the barrier)
  {  // final = T * final (column-major), entries ((a0 b0 + a1 b1) + a2 b2) + a3 b3 un-fused, as the oracle
    float v = 0.f;
    if (lane < 16) {
      const int r = lane & 3, c = lane >> 2;
      const float* A = S->inc_T;
      const float* B = S->final_T;
      v = fmul(A[0 * 4 + r], B[c * 4 + 0]);
      v = fadd(v, fmul(A[1 * 4 + r], B[c * 4 + 1]));
      v = fadd(v, fmul(A[2 * 4 + r], B[c * 4 + 2]));
      v = fadd(v, fmul(A[3 * 4 + r], B[c * 4 + 3]));
    }
    __syncwarp();
    if (lane < 16) S->final_T[lane] = v;
  }
  if (lane == 0) {
    const float* T = S->inc_T;
    const int n_corr = S->n_corr;
    const int it = ++S->iterations;
    const double mse = sums[16] * rcp_fast((double)n_corr);  // within 1 ulp of the quotient
    S->mse = mse;
    // DefaultConvergenceCriteria::hasConverged
    int state = RSPCL_CONV_NOT_CONVERGED;
    bool conv = false;
    if (it >= prm.max_iterations) {
      state = RSPCL_CONV_ITERATIONS;
      conv = true;
    } else {
      const double cos_angle = 0.5 * (double)(fadd(fadd(fadd(T[0], T[5]), T[10]), -1.0f));
      const double tr2 = (double)fadd(fadd(fmul(T[12], T[12]), fmul(T[13], T[13])), fmul(T[14], T[14]));
      if (cos_angle >= prm.rot_thr && tr2 <= prm.trans_thr) {
        state = RSPCL_CONV_TRANSFORM;
        conv = true;
      } else {
        const double prev = S->prev_mse;
        if (fabs(mse - prev) < prm.mse_abs) {
          state = RSPCL_CONV_ABS_MSE;
          conv = true;
        } else if (fabs(mse - prev) / prev < prm.mse_rel) {
          state = RSPCL_CONV_REL_MSE;
          conv = true;
        } else {
          S->prev_mse = mse;
        }
      }
    }
    S->state = state;
    if (conv) {
      S->converged = 1;
      S->done = 1;
      atomicSub(n_active, 1);
    }
  }
  __syncwarp();
}

template <bool BRUTE>
__global__ void __launch_bounds__(IT) k_icp_step(float4* __restrict__ work, const int* __restrict__ count, int stride,
                                                 const IcpState* __restrict__ st, DevGrid g, IcpDevParams prm,
                                                 const float4* __restrict__ tgt, const int* __restrict__ tcount,
                                                 int tstride, double* __restrict__ partials,
                                                 int* __restrict__ corr_out, int corr_iters) {
  __shared__ float M[16];
  __shared__ double s_red[IT / 32][NRED];
  __shared__ float4 tile[BRUTE ? IT : 1];
  const int seg = blockIdx.y;
  if (st[seg].done) return;
  const int n = count[seg];
  const int apply = st[seg].apply_inc;
  if (threadIdx.x < 16) M[threadIdx.x] = st[seg].inc_T[threadIdx.x];
  __syncthreads();
  // correspondence dump [iteration][pair][stride] (match index or -1); corr_iters = 1: PCL-style first correspondences
  const bool want_corr = corr_out != nullptr && st[seg].iterations < corr_iters;
  int* first_corr = want_corr ? corr_out + (size_t)st[seg].iterations * gridDim.y * stride : nullptr;
  double acc[NRED];
#pragma unroll
  for (int k = 0; k < NRED; ++k) acc[k] = 0.0;

  const int tseg = g.shared_target ? 0 : seg;
  const int nt = BRUTE ? tcount[tseg] : 0;
  for (int base = blockIdx.x * IT; base < n; base += gridDim.x * IT) {  // block-uniform trip count
    const int i = base + threadIdx.x;
    float4 p = make_float4(NAN, NAN, NAN, 0.f);
    if (i < n) {
      p = work[(size_t)seg * stride + i];
      if (apply && finite3(p.x, p.y, p.z)) {
        const float3 q = xform_point(M, p.x, p.y, p.z);
        p.x = q.x;
        p.y = q.y;
        p.z = q.z;
        work[(size_t)seg * stride + i] = p;
      }
    }
    int j = -1;
    float d2 = INFINITY;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (BRUTE) {
      const float4* T = tgt + (size_t)tseg * tstride;
      for (int tb = 0; tb < nt; tb += IT) {
        __syncthreads();
        if (tb + threadIdx.x < nt) tile[threadIdx.x] = T[tb + threadIdx.x];
        __syncthreads();
        const int m = min(IT, nt - tb);
        for (int k = 0; k < m; ++k) {
          const float4 c = tile[k];
          const float d = dist2_l2simple(p.x, p.y, p.z, c.x, c.y, c.z);
          if (d < d2) {
            d2 = d;
            j = tb + k;
            t = c;
          }
        }
      }
    } else if (i < n && finite3(p.x, p.y, p.z)) {
      j = grid_nn_bounded(g, seg, p.x, p.y, p.z, prm.search_r, &d2, &t);
    }
    const bool ok = (i < n) && j >= 0 && !((double)d2 > prm.max_dist_sqr);
    if (want_corr && i < n) first_corr[(size_t)seg * stride + __float_as_int(p.w)] = ok ? j : -1;  // .w = original index
    if (ok) {
      const double sx = p.x, sy = p.y, sz = p.z, tx = t.x, ty = t.y, tz = t.z;
      acc[0] += 1.0;
      acc[1] += sx; acc[2] += sy; acc[3] += sz;
      acc[4] += tx; acc[5] += ty; acc[6] += tz;
      acc[7] += sx * tx; acc[8] += sx * ty; acc[9] += sx * tz;
      acc[10] += sy * tx; acc[11] += sy * ty; acc[12] += sy * tz;
      acc[13] += sz * tx; acc[14] += sz * ty; acc[15] += sz * tz;
      acc[16] += (double)d2;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NRED; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) s_red[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NRED) {
    double v = 0;
#pragma unroll
    for (int w = 0; w < IT / 32; ++w) v += s_red[w][threadIdx.x];
    partials[((size_t)seg * gridDim.x + blockIdx.x) * NRED + threadIdx.x] = v;
  }
}

// point-sharded mode: CTA partials -> per-pair totals (then all-reduced across ranks before the solve)
__global__ void __launch_bounds__(256) k_icp_sum_partials(const double* __restrict__ partials, int nblk, const IcpState* __restrict__ st,
                                                          double* __restrict__ totals);

// K5b: one warp per pair combines the CTA partials in block order (deterministic) and runs the solve
// Sum of the CTA partials of one pair in a FIXED order: the 256 threads form 8 groups of 32 lanes, lane k < 17 of group g
// adds value k of blocks g, g + 8, g + 16, ... (coalesced 136-byte rows, four independent chains), the 8 group totals are
// then added in group order -- reproducible run to run.  (One warp summing thousands of partials serially cost 170 us
// per iteration on a 50 M-point pair.)
constexpr int SOLVE_T = 256;
__device__ __forceinline__ void icp_combine_partials(const double* __restrict__ P /* [nblk][NRED] */, int nblk, double* s_tmp /* [SOLVE_T] */,
                                                     double* sums /* [NRED] */) {
  const int k = threadIdx.x & 31, grp = threadIdx.x >> 5;
  constexpr int NG = SOLVE_T / 32;
  double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
  if (k < NRED) {
    int b = grp;
    for (; b + 3 * NG < nblk; b += 4 * NG) {
      v0 += P[(size_t)b * NRED + k];
      v1 += P[(size_t)(b + NG) * NRED + k];
      v2 += P[(size_t)(b + 2 * NG) * NRED + k];
      v3 += P[(size_t)(b + 3 * NG) * NRED + k];
    }
    for (; b < nblk; b += NG) v0 += P[(size_t)b * NRED + k];
  }
  s_tmp[threadIdx.x] = (v0 + v1) + (v2 + v3);
  __syncthreads();
  if (threadIdx.x < NRED) {
    double t = 0;
#pragma unroll
    for (int g = 0; g < NG; ++g) t += s_tmp[g * 32 + threadIdx.x];
    sums[threadIdx.x] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_icp_sum_partials(const double* __restrict__ partials, int nblk, const IcpState* __restrict__ st,
                                                          double* __restrict__ totals) {
  const int seg = blockIdx.x;
  __shared__ double sums[NRED];
  __shared__ double s_tmp[SOLVE_T];
  // finished pairs contribute zeros on every rank (their kernels exit early and leave stale partials)
  if (st[seg].done) {
    if (threadIdx.x < NRED) totals[seg * NRED + threadIdx.x] = 0.0;
    return;
  }
  icp_combine_partials(partials + (size_t)seg * nblk * NRED, nblk, s_tmp, sums);
  if (threadIdx.x < NRED) totals[seg * NRED + threadIdx.x] = sums[threadIdx.x];
}

__global__ void __launch_bounds__(SOLVE_T) k_icp_solve(IcpState* __restrict__ st, const double* __restrict__ partials, int nblk,
                                                       IcpDevParams prm, int* __restrict__ n_active) {
  const int seg = blockIdx.x;
  if (st[seg].done) return;
  __shared__ double sums[NRED];
  __shared__ double s_tmp[SOLVE_T];
  if (nblk <= 64) {  // few CTAs per pair (batches of small pairs): one warp's serial sum is the shortest path
    if (threadIdx.x < NRED) {
      double v = 0;
      const double* P = partials + (size_t)seg * nblk * NRED + threadIdx.x;
      for (int b = 0; b < nblk; ++b) v += P[(size_t)b * NRED];
      sums[threadIdx.x] = v;
    }
    __syncthreads();
  } else {
    icp_combine_partials(partials + (size_t)seg * nblk * NRED, nblk, s_tmp, sums);
  }
  if (threadIdx.x < 32) icp_solve_pair(&st[seg], sums, prm, n_active, threadIdx.x);
}

// point-sharded mode, peer-memory path: the same warp combines this rank's CTA partials, exchanges the 17 totals with every
// other rank through NVLink peer memory (one-shot all-reduce, comm.cu / common.cuh) and runs the solve -- one launch instead
// of {sum partials, ncclAllReduce, solve}.  Every rank sums the ranks' totals in rank order: identical bits, identical
// convergence decisions, no broadcast.
__global__ void __launch_bounds__(SOLVE_T) k_icp_solve_peer(IcpState* __restrict__ st, const double* __restrict__ partials, int nblk,
                                                            IcpDevParams prm, int* __restrict__ n_active, PeerX X) {
  const int seg = blockIdx.x;
  if (st[seg].done) return;  // (identical on every rank)
  __shared__ double sums[NRED];
  __shared__ double s_tmp[SOLVE_T];
  icp_combine_partials(partials + (size_t)seg * nblk * NRED, nblk, s_tmp, sums);
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    double v = lane < NRED ? sums[lane] : 0.0;
    v = peer_allreduce_warp(X, v, lane, NRED, seg);
    __syncwarp();
    if (lane < NRED) sums[lane] = v;
    __syncwarp();
    icp_solve_pair(&st[seg], sums, prm, n_active, lane);
  }
}

// cell key of every source point (for the spatial sort of the working cloud: neighbouring lanes then query
// neighbouring cells, which removes most of the divergence of the cell / candidate loops)
__global__ void k_source_keys(const float4* __restrict__ src, const int* __restrict__ count, int stride, int pstride,
                              float inv_cs, unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pstride; i += gridDim.x * blockDim.x) {
    // padding entries sort to the end of THEIR segment's block, so block s always holds segment s
    unsigned long long k = ((unsigned long long)(unsigned)seg << 48) | 0xFFFFFFFFFFFFull;
    if (i < n) {
      const float4 p = src[(size_t)seg * stride + i];
      int ix = 0, iy = 0, iz = 0;
      if (finite3(p.x, p.y, p.z)) {
        ix = max(-32767, min(32766, __float2int_rd(__fmul_rn(p.x, inv_cs))));
        iy = max(-32767, min(32766, __float2int_rd(__fmul_rn(p.y, inv_cs))));
        iz = max(-32767, min(32766, __float2int_rd(__fmul_rn(p.z, inv_cs))));
      }
      k = ((unsigned long long)(unsigned)seg << 48) | ((unsigned long long)(unsigned)(iz + 32768) << 32) |
          ((unsigned long long)(unsigned)(iy + 32768) << 16) | (unsigned long long)(unsigned)(ix + 32768);
    }
    keys[(size_t)seg * pstride + i] = k;
    vals[(size_t)seg * pstride + i] = i;
  }
}

// work[seg][j] = src[seg][perm[j]] with the original index kept in .w (rgba is not needed by the iterations)
__global__ void k_copy_work_perm(const float4* __restrict__ src, const int* __restrict__ count, int stride_src,
                                 const int* __restrict__ perm, int pstride, float4* __restrict__ work, int stride_work) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int i = perm[(size_t)seg * pstride + j];
    float4 p = src[(size_t)seg * stride_src + i];
    p.w = __int_as_float(i);
    work[(size_t)seg * stride_work + j] = p;
  }
}

__global__ void k_gather_final(const IcpState* __restrict__ st, float* __restrict__ T, int n_seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_seg * 16) T[i] = st[i / 16].final_T[i % 16];
}

__global__ void k_pack_i32(const int* __restrict__ v, const int* __restrict__ count, const int* __restrict__ off, int stride,
                           int* __restrict__ out) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[off[seg] + i] = v[(size_t)seg * stride + i];
}

#include "icp_persist.cuh"

// ------------------------------------------------------------------------------------------------------------------
// Certified-cache correspondence passes for clouds that do not fit the shared-memory kernel (global-memory grid).
// Per source point: ci = position of its cached match inside g.sorted (-1: none), lb = lower bound of the true distance
// to every OTHER target point (to every target point if ci < 0).  See icp_persist.cuh for why this is exact.
//   k_icp_stream  one streaming pass over the working cloud, which is a pair of 16-byte records per point:
//                 A = {x, y, z, certified bound} (rewritten every iteration) and B = {matched target x, y, z, its position
//                 in g.sorted} (rewritten only by a re-query).  Move the point, decay its bound by the distance moved,
//                 re-measure the cached match FROM THE RECORD (no gather into the target: round 1 paid a 32-byte sector
//                 for every 16-byte match) and, if the bound still decides, accumulate the 17 sums.  Points whose bound
//                 no longer decides are appended to the block's slice of a work list IN POINT ORDER (ballot + prefix
//                 counts), so every later sum is taken in a fixed order and results stay reproducible run to run.
//                 Bytes per point: 32 R + 16 W = 48, all streaming (algorithmic: 32).
//   k_icp_rescan  re-queries the listed points in the voxel-hash grid (exact NN inside the gate ball, second-best
//                 distance for the new bound) and adds its own partial sums.  The block slices are concatenated through
//                 an exclusive scan of their counts (k_wl_offsets), item t always goes to the same thread.
// Exact re-query of one source point: nearest target point inside the gate ball (grid scan) merged with the cached
// incumbent `c`, the new cache position and the new certified bound.  Returns the winner's original index or -1.
__device__ __forceinline__ int icp_requery(const DevGrid& g, int seg, const float4& p, int c, float r, float* out_bd,
                                           int* out_pos, float4* out_t, float* out_lb) {
  float bdc = INFINITY;
  float4 tc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c >= 0) {
    tc = __ldg(&g.sorted[c]);
    bdc = dist2_l2simple(p.x, p.y, p.z, tc.x, tc.y, tc.z);
  }
  float bd, d2nd;
  int pos;
  float4 t;
  int best = grid_nn_top2(g, seg, p.x, p.y, p.z, r, c, &bd, &d2nd, &pos, &t);
  // points outside the scanned cells are farther than r minus the rounding of the cell-boundary test
  float outer = r - 2e-7f * (fabsf(p.x) + fabsf(p.y) + fabsf(p.z) + r);
  if (best < 0) {
    // Nothing inside the gate ball.  Certify a whole cell size instead (3x3x3 cells), so that this point is not looked
    // at again until it has moved that far; the nearest point found (beyond the gate) becomes the cached incumbent.
    float d27, d27b;
    float4 t27;
    const int p27 = grid_scan27_top2(g, seg, p.x, p.y, p.z, &d27, &d27b, &t27);
    if (d27 >= 0.f) {
      outer = g.cs - 2e-7f * (fabsf(p.x) + fabsf(p.y) + fabsf(p.z) + g.cs);
      if (p27 >= 0) {
        best = __float_as_int(t27.w);
        bd = d27;
        d2nd = d27b;
        pos = p27;
        t = t27;
      }
      c = -1;  // the cached point, if any, was part of this scan
    }
  }
  if (c >= 0) {  // merge the cached incumbent (skipped by the ball scan; it may lie outside the scanned ball)
    const int bc = __float_as_int(tc.w);
    if (bdc < bd || (bdc == bd && bc < best)) {
      d2nd = bd;
      bd = bdc;
      best = bc;
      pos = c;
      t = tc;
    } else {
      d2nd = fminf(d2nd, bdc);
    }
  }
  float lbn = fminf(sqrt_approx(d2nd), outer) * 0.9999f;
  if (best >= 0 && !(__fmaf_rn(sqrt_approx(bd), 1.0001f, 1e-7f) < lbn)) {
    // a winner that the new bound can never confirm is not worth caching: fold it into the bound over ALL points
    lbn = fminf(lbn, sqrt_approx(bd) * 0.9999f);
    pos = -1;
  }
  *out_lb = lbn;
  *out_bd = bd;
  *out_pos = pos;
  *out_t = t;
  return best;
}

template <bool FIRST>  // FIRST: nothing is cached yet (iteration 0), every point is queried in place (no work list)
__global__ void __launch_bounds__(IT) k_icp_stream(float4* __restrict__ work, const int* __restrict__ count, int stride,
                                                   const IcpState* __restrict__ st, DevGrid g, IcpDevParams prm,
                                                   float4* __restrict__ recB, const int* __restrict__ perm, int pstride,
                                                   int* __restrict__ wl, int* __restrict__ wlcount,
                                                   double* __restrict__ partials, int* __restrict__ corr_out, int corr_iters) {
  __shared__ float M[16];
  __shared__ double s_red[IT / 32][NRED];
  constexpr int SU = 4;  // tiles in flight per thread in the streaming pass (memory-level parallelism: 2 x SU 16-byte loads)
  __shared__ int s_wcnt[2][SU][IT / 32];
  const int seg = blockIdx.y;
  if (st[seg].done) return;
  const int n = count[seg];
  const int apply = st[seg].apply_inc;
  if (threadIdx.x < 16) M[threadIdx.x] = st[seg].inc_T[threadIdx.x];
  __syncthreads();
  const bool want_corr = corr_out != nullptr && st[seg].iterations < corr_iters;
  int* first_corr = want_corr ? corr_out + (size_t)st[seg].iterations * gridDim.y * stride : nullptr;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double acc[NRED];
#pragma unroll
  for (int k = 0; k < NRED; ++k) acc[k] = 0.0;
  const int chunk = (((n + (int)gridDim.x - 1) / (int)gridDim.x + IT - 1) / IT) * IT;  // contiguous slice per block
  const int lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
  const float r = prm.search_r;
  int nlist = 0, par = 0;
  if (!FIRST) {
    // Steady-state pass: SU tiles per trip -- all 2 x SU record loads of a thread are issued before the first is consumed
    // (the pass is bound by load latency at this occupancy, not by bytes), one barrier per SU tiles.
    for (int base = lo; base < hi; base += SU * IT, par ^= 1) {  // block-uniform trip count
      float4 pA[SU], pB[SU];
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        const int i = base + u * IT + threadIdx.x;
        if (i < hi) {
          pA[u] = work[(size_t)seg * stride + i];
          pB[u] = recB[(size_t)seg * stride + i];
        } else {
          pA[u] = make_float4(NAN, NAN, NAN, 0.f);
          pB[u] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        }
      }
      unsigned fbits = 0u;
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        const int i = base + u * IT + threadIdx.x;
        if (i >= hi) continue;
        const size_t gi = (size_t)seg * stride + i;
        float4 p = pA[u];
        float lbv = p.w;
        const bool fin = finite3(p.x, p.y, p.z);
        const int orig = want_corr ? perm[(size_t)seg * pstride + i] : 0;  // original source index (dump only)
        if (!fin) {
          if (want_corr) first_corr[(size_t)seg * stride + orig] = -1;
          continue;
        }
        if (apply) {
          const float3 q = xform_point(M, p.x, p.y, p.z);
          lbv = lbv - __fmaf_rn(sqrt_approx(dist2_l2simple(q.x, q.y, q.z, p.x, p.y, p.z)), 1.00001f, 1e-9f);
          p.x = q.x;
          p.y = q.y;
          p.z = q.z;
          p.w = lbv;
          work[gi] = p;
        }
        const float4 t = pB[u];  // record B: the cached match and where it lives in g.sorted
        const int c = __float_as_int(t.w);
        bool valid;
        float bd = INFINITY;
        if (c >= 0) {
          bd = dist2_l2simple(p.x, p.y, p.z, t.x, t.y, t.z);
          valid = __fmaf_rn(sqrt_approx(bd), 1.0001f, 1e-7f) < lbv;
        } else {
          valid = lbv > r;
        }
        if (!valid) {
          fbits |= 1u << u;
        } else {
          const bool ok = c >= 0 && !((double)bd > prm.max_dist_sqr);
          if (want_corr) first_corr[(size_t)seg * stride + orig] = ok ? __float_as_int(__ldg(&g.sorted[c]).w) : -1;
          if (ok) icp_accumulate(acc, p.x, p.y, p.z, t.x, t.y, t.z, bd);
        }
      }
      // ordered compaction of the flagged points of these SU tiles (ascending point index: tile, warp, lane)
      unsigned bal[SU];
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        bal[u] = __ballot_sync(0xffffffffu, (fbits >> u) & 1u);
        if (lane == 0) s_wcnt[par][u][wid] = __popc(bal[u]);
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < IT / 32; ++w) {
          const int cw = s_wcnt[par][u][w];
          before += (w < wid) ? cw : 0;
          total += cw;
        }
        if ((fbits >> u) & 1u)
          wl[(size_t)seg * stride + lo + nlist + before + __popc(bal[u] & ((1u << lane) - 1u))] = base + u * IT + threadIdx.x;
        nlist += total;
      }
    }
  }
  for (int base = lo; FIRST && base < hi; base += IT, par ^= 1) {  // first iteration: every point is queried in place
    const int i = base + threadIdx.x;
    bool flag = false;
    if (i < hi) {
      const size_t gi = (size_t)seg * stride + i;
      float4 p = work[gi];  // record A: position + certified bound
      float lbv = p.w;
      const bool fin = finite3(p.x, p.y, p.z);
      if (apply && fin) {
        const float3 q = xform_point(M, p.x, p.y, p.z);
        lbv = lbv - __fmaf_rn(sqrt_approx(dist2_l2simple(q.x, q.y, q.z, p.x, p.y, p.z)), 1.00001f, 1e-9f);
        p.x = q.x;
        p.y = q.y;
        p.z = q.z;
      }
      const int orig = want_corr ? perm[(size_t)seg * pstride + i] : 0;  // original source index (dump only)
      if (!fin) {
        if (want_corr) first_corr[(size_t)seg * stride + orig] = -1;
        if (FIRST) {
          p.w = 0.f;
          work[gi] = p;
          recB[gi] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        }
      } else if (FIRST) {
        float bd, lbn;
        int pos;
        float4 t;
        const int best = icp_requery(g, seg, p, -1, r, &bd, &pos, &t, &lbn);
        p.w = lbn;
        work[gi] = p;
        recB[gi] = make_float4(t.x, t.y, t.z, __int_as_float(pos));
        const bool ok = best >= 0 && !((double)bd > prm.max_dist_sqr);
        if (want_corr) first_corr[(size_t)seg * stride + orig] = ok ? best : -1;
        if (ok) icp_accumulate(acc, p.x, p.y, p.z, t.x, t.y, t.z, bd);
      } else {
        const float4 t = recB[gi];  // record B: the cached match and where it lives in g.sorted
        const int c = __float_as_int(t.w);
        p.w = lbv;
        work[gi] = p;
        bool valid;
        float bd = INFINITY;
        if (c >= 0) {
          bd = dist2_l2simple(p.x, p.y, p.z, t.x, t.y, t.z);
          valid = __fmaf_rn(sqrt_approx(bd), 1.0001f, 1e-7f) < lbv;
        } else {
          valid = lbv > r;
        }
        if (!valid) {
          flag = true;
        } else {
          const bool ok = c >= 0 && !((double)bd > prm.max_dist_sqr);
          if (want_corr) first_corr[(size_t)seg * stride + orig] = ok ? __float_as_int(__ldg(&g.sorted[c]).w) : -1;
          if (ok) icp_accumulate(acc, p.x, p.y, p.z, t.x, t.y, t.z, bd);
        }
      }
    }
    // (nothing is flagged in the first iteration: the work list stays empty)
    (void)flag;
  }
  if (threadIdx.x == 0) wlcount[seg * gridDim.x + blockIdx.x] = nlist;
#pragma unroll
  for (int k = 0; k < NRED; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) s_red[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NRED) {
    double v = 0;
#pragma unroll
    for (int w = 0; w < IT / 32; ++w) v += s_red[w][threadIdx.x];
    partials[((size_t)seg * 2 * gridDim.x + blockIdx.x) * NRED + threadIdx.x] = v;
  }
}

// exclusive scan of the per-block work-list counts of one segment (one CTA per segment; nblk <= a few thousand)
__global__ void __launch_bounds__(1024) k_wl_offsets(const int* __restrict__ wlcount, int nblk, const IcpState* __restrict__ st,
                                                     int* __restrict__ wloff) {
  __shared__ int s_part[32];
  __shared__ int s_carry;
  const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (st[seg].done) return;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int b = base + tid;
    const int v = b < nblk ? wlcount[seg * nblk + b] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_part[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int w = s_part[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      s_part[lane] = wi - w;
    }
    __syncthreads();
    const int carry = s_carry;
    if (b < nblk) wloff[seg * (nblk + 1) + b] = carry + s_part[wid] + incl - v;
    __syncthreads();
    if (tid == 1023) s_carry = carry + s_part[wid] + incl;
    __syncthreads();
  }
  if (tid == 0) wloff[seg * (nblk + 1) + nblk] = s_carry;
}

__global__ void __launch_bounds__(IT, RESCAN_MINB) k_icp_rescan(float4* __restrict__ work, const int* __restrict__ count, int stride,
                                                   const IcpState* __restrict__ st, DevGrid g, IcpDevParams prm,
                                                   float4* __restrict__ recB, const int* __restrict__ perm, int pstride,
                                                   const int* __restrict__ wl, const int* __restrict__ wloff,
                                                   double* __restrict__ partials, int* __restrict__ corr_out, int corr_iters) {
  __shared__ double s_red[IT / 32][NRED];
  // the 17 running sums of every thread live in shared memory ([sum][thread]: conflict-free), not in 34 registers: the
  // kernel is latency-bound and runs at four CTAs per SM, where registers are what limits the probes in flight
  __shared__ double s_acc[NRED][IT];
  const int seg = blockIdx.y;
  if (st[seg].done) return;
  const int n = count[seg];
  const bool want_corr = corr_out != nullptr && st[seg].iterations < corr_iters;
  int* first_corr = want_corr ? corr_out + (size_t)st[seg].iterations * gridDim.y * stride : nullptr;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nblk = (int)gridDim.x;
  const int chunk = (((n + nblk - 1) / nblk + IT - 1) / IT) * IT;
  const int* off = wloff + seg * (nblk + 1);
  const int total = off[nblk];
  const float r = prm.search_r;
#pragma unroll
  for (int k = 0; k < NRED; ++k) s_acc[k][threadIdx.x] = 0.0;
  // item t of the segment's concatenated (block-ordered) work list: a fixed item -> thread map, so sums stay reproducible
  for (int t_ = blockIdx.x * IT + threadIdx.x; t_ < total; t_ += nblk * IT) {
    int lo_b = 0, hi_b = nblk;  // largest b with off[b] <= t_
    while (hi_b - lo_b > 1) {
      const int mid = (lo_b + hi_b) >> 1;
      if (__ldg(&off[mid]) <= t_) lo_b = mid; else hi_b = mid;
    }
    const int i = wl[(size_t)seg * stride + (size_t)lo_b * chunk + (t_ - __ldg(&off[lo_b]))];
    const size_t gi = (size_t)seg * stride + i;
    const float4 p = work[gi];  // already moved by the streaming pass
    float bd, lbn;
    int pos;
    float4 t;
    const int best = icp_requery(g, seg, p, __float_as_int(recB[gi].w), r, &bd, &pos, &t, &lbn);
    work[gi].w = lbn;
    recB[gi] = make_float4(t.x, t.y, t.z, __int_as_float(pos));
    const bool ok = best >= 0 && !((double)bd > prm.max_dist_sqr);
    if (want_corr) first_corr[(size_t)seg * stride + perm[(size_t)seg * pstride + i]] = ok ? best : -1;
    if (ok) {
      double acc[NRED];
#pragma unroll
      for (int k = 0; k < NRED; ++k) acc[k] = s_acc[k][threadIdx.x];
      icp_accumulate(acc, p.x, p.y, p.z, t.x, t.y, t.z, bd);
#pragma unroll
      for (int k = 0; k < NRED; ++k) s_acc[k][threadIdx.x] = acc[k];
    }
  }
#pragma unroll
  for (int k = 0; k < NRED; ++k) {
    const double v = warp_sum(s_acc[k][threadIdx.x]);
    if (lane == 0) s_red[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NRED) {
    double v = 0;
#pragma unroll
    for (int w = 0; w < IT / 32; ++w) v += s_red[w][threadIdx.x];
    partials[((size_t)seg * 2 * gridDim.x + gridDim.x + blockIdx.x) * NRED + threadIdx.x] = v;
  }
}


// ---- launch plumbing of the persistent kernel
typedef void (*PersistFn)(const PersistArgs, const IcpDevParams);
PersistFn persist_fn(bool dbg) { return dbg ? k_icp_persist<true> : k_icp_persist<false>; }

}  // namespace

extern "C" void rspcl_icp_reference_params(rspcl_icp_params* p) {
  p->max_iterations = 100;              // icp:42
  p->min_correspondences = 3;
  p->max_corr_dist = 0.01;              // icp:43
  p->transformation_epsilon = 1;        // icp:44
  p->euclidean_fitness_epsilon = 1000;  // icp:45
  p->mse_threshold_absolute = 1e-12;
}

int radix_sort_pairs(rspcl_ctx* ctx, unsigned long long* keys, int* vals, unsigned long long* tmp_keys, int* tmp_vals,
                     long long n);

// Per-pair cluster sizes for one batch (1..P_CLMAX CTAs per pair): the smallest cost bound L such that every pair can be
// given enough CTAs to stay below it and the whole batch still fits one wave, i.e. the launch time (set by the pair with
// the most expensive slice) is minimised and the chip is filled; left-over CTAs go to the most expensive slices.  The
// cost of a slice is its length, doubled when it exceeds the P_CHUNK points a CTA keeps in registers (a streamed slice
// re-reads and re-writes its points every iteration).  64 pairs of 4.5 - 12 k points on 148 SMs end up with 2, 3 or 4
// CTAs each, every slice register-resident.
static void plan_clusters(const std::vector<int>& cnt, double sm_count, const double* weight /* [P_CLMAX + 1]: SMs a cluster of c CTAs takes */,
                          std::vector<int>* cl) {
  const int S = (int)cnt.size();
  cl->assign(S, 1);
  if (S > (int)sm_count) return;  // several waves anyway: one CTA per pair
  auto cost = [](int n, int c) {
    const int len = (n + c - 1) / c;
    return len > P_CHUNK ? 2 * len : len;
  };
  auto need = [&](int L, std::vector<int>* out) {
    double tot = 0;
    for (int s = 0; s < S; ++s) {
      int c = 1;
      while (c < P_CLMAX && cost(cnt[s], c) > L) ++c;
      if (out) (*out)[s] = c;
      tot += weight[c];
    }
    return tot;
  };
  int lo = P_THREADS, hi = 1;  // a slice below one point per thread gains nothing
  for (int v : cnt) hi = 2 * v > hi ? 2 * v : hi;
  if (hi < lo) hi = lo;
  if (need(lo, nullptr) > sm_count) {
    while (lo < hi) {  // smallest L with need(L) <= sm_count
      const int mid = (lo + hi) / 2;
      if (need(mid, nullptr) <= sm_count) hi = mid; else lo = mid + 1;
    }
  }
  double used = need(lo, cl);
  while (used < sm_count) {  // hand the remaining CTAs to the most expensive slices
    int best = -1, bc = 0;
    for (int s = 0; s < S; ++s) {
      const int c = (*cl)[s];
      if (c >= P_CLMAX || cnt[s] / (c + 1) < P_THREADS || used - weight[c] + weight[c + 1] > sm_count) continue;
      const int v = cost(cnt[s], c);
      if (v > bc) {
        bc = v;
        best = s;
      }
    }
    if (best < 0) break;
    used += weight[(*cl)[best] + 1] - weight[(*cl)[best]];
    (*cl)[best] += 1;
  }
}

// Device-level align used by rspcl_icp_align and the pairwise pipeline.  d_guess: n_seg x 16 floats on the device or
// null.  h_results: host array (prev_mse read as input).  With o.h_results2 the call runs TWO aligns per pair: the second
// from identity on the first one's output (the reference's coarse + fine ICP, icp:95 + icp:111), fused into one launch of
// the persistent kernel when the pair fits it.  Synchronises before returning.
int icp_align_device(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt, const rspcl_icp_params* prm,
                     const float* d_guess, rspcl_icp_result* h_results, rspcl_cloud* aligned, const IcpAlignOpts& o) {
  const int S = src->n_seg;
  const int shared_target = (tgt->n_seg == 1 && S > 1) ? 1 : 0;
  if (!shared_target && tgt->n_seg != S) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "icp_align: src has %d segments, tgt %d", S, tgt->n_seg);
  if (aligned && (aligned->n_seg != S || aligned->stride < src->max_count_hint))
    RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "icp_align: aligned output too small");
  if (prm->max_iterations < 1) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "icp_align: max_iterations < 1");
  const bool two_stage = o.h_results2 != nullptr;
  int* d_first_corr = o.d_corr_out;
  const int corr_iters = o.d_corr_out ? (o.corr_iters > 0 ? o.corr_iters : 1) : 0;

  IcpDevParams dp;
  dp.max_iterations = prm->max_iterations;
  dp.min_corr = prm->min_correspondences;
  dp.max_dist_sqr = prm->max_corr_dist * prm->max_corr_dist;
  dp.rot_thr = 1.0 - prm->transformation_epsilon;
  dp.trans_thr = prm->transformation_epsilon;
  dp.mse_abs = prm->mse_threshold_absolute;
  dp.mse_rel = prm->euclidean_fitness_epsilon;
  // grid mode needs cells of ~2x the gate that still fit the 16-bit cell coordinates; otherwise brute force
  const bool brute = !(prm->max_corr_dist > 0.0 && prm->max_corr_dist < 5.0);
  const float cs = brute ? 1.f : (float)(prm->max_corr_dist * 2.05);
  dp.search_r = brute ? 0.f : (float)(prm->max_corr_dist * 1.01);

  const int wstride = src->stride ? src->stride : 1;
  Scratch scr(ctx);  // every stream-ordered scratch buffer of this call: released on every exit path
  float4* work = nullptr;
  IcpState *st = nullptr, *st2 = nullptr;
  double *partials = nullptr, *d_prev = nullptr;
  int *n_active = nullptr, *d_range = nullptr;
  float* d_T = nullptr;
  unsigned char* d_skip = nullptr;
  const int nblk = blocks_per_seg(ctx, S, src->max_count_hint, IT);
  CU(ctx, scr.alloc(&work, (size_t)S * wstride));
  CU(ctx, scr.alloc(&st, (size_t)S));
  if (two_stage) CU(ctx, scr.alloc(&st2, (size_t)S));
  // global-memory path: certified-cache passes (k_icp_stream + k_icp_rescan) unless RSPCL_ICP_CACHE=0
  const char* cache_env = getenv("RSPCL_ICP_CACHE");
  const bool use_cache = !brute && !(cache_env && cache_env[0] == '0');
  const int nblk_total = use_cache ? 2 * nblk : nblk;
  float4* g_rb = nullptr;  // record B of every source point (cached match)
  int *g_wl = nullptr, *g_wlcount = nullptr, *g_wloff = nullptr;
  CU(ctx, scr.alloc(&partials, (size_t)S * nblk_total * NRED));
  CU(ctx, scr.alloc(&d_prev, (size_t)S));
  CU(ctx, scr.alloc(&n_active, 1));
  CU(ctx, scr.alloc(&d_range, 1));
  CU(ctx, scr.alloc(&d_T, (size_t)S * 16));
  std::vector<double> prev(S);
  int n_todo = S;
  for (int s = 0; s < S; ++s) prev[s] = h_results[s].prev_mse;
  if (o.h_skip) {
    CU(ctx, scr.alloc(&d_skip, (size_t)((S + 3) & ~3)));
    std::vector<unsigned char> sk((size_t)((S + 3) & ~3), 0);
    for (int s = 0; s < S; ++s) {
      sk[s] = o.h_skip[s] ? 1 : 0;
      n_todo -= sk[s];
    }
    CU(ctx, small_h2d(ctx, d_skip, sk.data(), sk.size()));
  }
  CU(ctx, small_h2d(ctx, d_prev, prev.data(), S * sizeof(double)));
  CU(ctx, small_h2d(ctx, n_active, &n_todo, sizeof(int)));
  CU(ctx, cudaMemsetAsync(d_range, 0, sizeof(int), ctx->stream));
  k_icp_init<<<div_up(S, 128), 128, 0, ctx->stream>>>(st, d_guess, d_prev, d_skip, S);
  LAUNCH_CHECK(ctx);
  dim3 gcopy(blocks_per_seg(ctx, S, src->max_count_hint, 256), S);
  // Working cloud of the GLOBAL-MEMORY path = source in SPATIAL order (radix sort by cell key), original index kept in
  // .w: neighbouring lanes then probe neighbouring cells (coalesced / cached probes, far less divergence).  Sums are
  // order-independent up to fp64 rounding; correspondences and the aligned output are written in original order.
  // (The persistent kernel reads the source itself, unsorted, straight into registers.)
  const float sort_inv_cs = brute ? 10.0f : 1.0f / (float)(prm->max_corr_dist * 4.1);
  int* perm = nullptr;
  const int pstride = src->max_count_hint > 0 ? src->max_count_hint : 1;
  auto build_sorted_work = [&]() -> int {
    const long long N = (long long)S * pstride;
    unsigned long long *keys = nullptr, *tkeys = nullptr;
    int* tvals = nullptr;
    CU(ctx, scr.alloc(&keys, (size_t)N));
    CU(ctx, scr.alloc(&tkeys, (size_t)N));
    CU(ctx, scr.alloc(&perm, (size_t)N));
    CU(ctx, scr.alloc(&tvals, (size_t)N));
    dim3 gk(blocks_per_seg(ctx, S, pstride, 256), S);
    k_source_keys<<<gk, 256, 0, ctx->stream>>>(src->pts, src->count, src->stride, pstride, sort_inv_cs, keys, perm);
    LAUNCH_CHECK(ctx);
    int rcs = radix_sort_pairs(ctx, keys, perm, tkeys, tvals, N);
    if (rcs) return rcs;
    k_copy_work_perm<<<gk, 256, 0, ctx->stream>>>(src->pts, src->count, src->stride, perm, pstride, work, wstride);
    LAUNCH_CHECK(ctx);
    return RSPCL_OK;
  };
  const bool sharded = ctx->sharded_call && ctx->nccl_comm && ctx->nranks > 1;
  bool want_persist = false;
  {
    const char* env = getenv("RSPCL_ICP_PERSIST");
    // (a target above the shared-memory capacity only sends ITS pair to the global-memory path, see below)
    want_persist = !(env && env[0] == '0') && !o.no_persist && !sharded && !brute && src->max_count_hint > 0 &&
                   src->max_count_hint < 65536 && (S > 1 || tgt->max_count_hint <= P_NTMAX);
  }

  // ---- persistent shared-memory path: one cluster per pair, all iterations (of both aligns) in one launch (icp_persist.cuh)
  std::vector<unsigned char> finished(S, 0);  // pairs the persistent kernel has completed (all stages)
  std::vector<IcpState> hst(S), hst2(two_stage ? S : 0);
  int active_init = n_todo;  // pairs the global-memory path still has to run
  double* totals = nullptr;
  if (sharded) CU(ctx, scr.alloc(&totals, (size_t)S * NRED));
  if (want_persist) {
    if (!ctx->persist_ready) {
      for (int d = 0; d < 2; ++d)
        CU(ctx, cudaFuncSetAttribute(persist_fn(d != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PersistSmem)));
      for (int a = 0; a < RSPCL_AUX_STREAMS; ++a) {
        CU(ctx, cudaStreamCreateWithFlags(&ctx->aux[a], cudaStreamNonBlocking));
        CU(ctx, cudaEventCreateWithFlags(&ctx->ev_join[a], cudaEventDisableTiming));
      }
      CU(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
      for (int c = 1; c <= P_CLMAX; ++c) {  // SMs a cluster of c CTAs effectively occupies = SMs / co-resident clusters
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3((unsigned)(c * ctx->sm_count));
        q.blockDim = dim3(P_THREADS);
        q.dynamicSmemBytes = sizeof(PersistSmem);
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = (unsigned)c;
        qa[0].val.clusterDim.y = qa[0].val.clusterDim.z = 1;
        q.attrs = qa;
        q.numAttrs = 1;
        int n_act = 0;
        if (cudaOccupancyMaxActiveClusters(&n_act, persist_fn(false), &q) != cudaSuccess || n_act <= 0) {
          cudaGetLastError();
          n_act = ctx->sm_count / c;
        }
        ctx->cluster_weight[c] = (double)ctx->sm_count / n_act;
        if (ctx->cluster_weight[c] < c) ctx->cluster_weight[c] = c;
      }
      ctx->cluster_weight[0] = 0;
      if (getenv("RSPCL_PERSIST_DBG"))
        fprintf(stderr, "persist: SMs per cluster of 1..8 CTAs: %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f\n", ctx->cluster_weight[1], ctx->cluster_weight[2],
                ctx->cluster_weight[3], ctx->cluster_weight[4], ctx->cluster_weight[5], ctx->cluster_weight[6], ctx->cluster_weight[7], ctx->cluster_weight[8]);
      ctx->persist_ready = true;
    }
    int* d_status = nullptr;
    CU(ctx, scr.alloc(&d_status, (size_t)2 * S));  // [S] status, [S] cluster start times
    CU(ctx, cudaMemsetAsync(d_status, 0, (size_t)2 * S * sizeof(int), ctx->stream));
    std::vector<int> hcnt(S);
    if (o.h_src_counts) {
      for (int s = 0; s < S; ++s) hcnt[s] = o.h_src_counts[s];
    } else {
      CU(ctx, small_d2h(ctx, hcnt.data(), src->count, (size_t)S * sizeof(int)));
      CU(ctx, ctx_sync(ctx));
    }
    // pairs in decreasing size (longest-processing-time-first), cluster size per pair, one launch per cluster size
    std::vector<int> cl;
    // SMs one wave may use: clusters of 3 or more do not tile every GPC, and a mix of sizes is placed greedily, so the
    // planner charges every cluster size what the occupancy query says it takes and keeps two SMs in reserve (measured:
    // a batch planned for all 148 SMs left a few clusters waiting for a second wave, +20 % launch time)
    // The occupancy query charges a cluster of 3 what it costs when the WHOLE chip runs clusters of 3 (3.2 SMs); in a mix
    // the 2-CTA clusters fill the SMs those leave, so plans of up to ~152 "SMs" are resident on 148 (measured: 150 and 152
    // run the 64-pair batch in 1.26 ms, 146 in 1.38, 154 in 1.51 with a second wave).  The planner therefore starts at
    // sm_count + 2 and backs off, two SMs at a time down to sm_count - 2, whenever a launch reports a cluster that started
    // long after the others (see the start-time feedback after the launch).
    double budget = ctx->sm_count + ctx->persist_extra;
    bool budget_forced = false;
    if (const char* e = getenv("RSPCL_PERSIST_BUDGET"))
      if (atof(e) > 0) budget = atof(e), budget_forced = true;  // tuning knob
    // Waves.  A batch that fits the chip -- or whose pairs all fit one CTA's registers, so that the hardware can simply
    // stream single-CTA clusters through the SMs -- is one wave.  A larger batch of larger pairs (e.g. 4096 frame pairs
    // of ~7 k points) is cut, in decreasing size, into waves that each fill the chip with register-resident slices.
    cl.assign(S, 1);
    std::vector<int> todo;
    int max_cnt = 0;
    for (int s = 0; s < S; ++s)
      if (!(o.h_skip && o.h_skip[s])) {
        todo.push_back(s);
        max_cnt = hcnt[s] > max_cnt ? hcnt[s] : max_cnt;
      }
    std::stable_sort(todo.begin(), todo.end(), [&](int a, int b) { return hcnt[a] > hcnt[b]; });
    std::vector<std::vector<int>> waves;
    if ((int)todo.size() <= (int)budget || max_cnt <= P_CHUNK) {
      waves.push_back(todo);
    } else {
      double need = 0;
      std::vector<int> cur;
      for (int s : todo) {
        int cmin = (hcnt[s] + P_CHUNK - 1) / P_CHUNK;
        cmin = cmin < 1 ? 1 : (cmin > P_CLMAX ? P_CLMAX : cmin);
        const double w = ctx->cluster_weight[cmin];
        if (!cur.empty() && need + w > budget) {
          waves.push_back(cur);
          cur.clear();
          need = 0;
        }
        cur.push_back(s);
        need += w;
      }
      if (!cur.empty()) waves.push_back(cur);
    }
    // launch groups: per wave, per cluster size (largest first), pairs in decreasing size
    struct Group { int off, n, cl, wave; };
    std::vector<Group> groups;
    std::vector<int> order;
    int max_cl = 1;
    for (size_t wv = 0; wv < waves.size(); ++wv) {
      std::vector<int> wc(waves[wv].size()), wcl;
      for (size_t k = 0; k < wc.size(); ++k) wc[k] = hcnt[waves[wv][k]];
      plan_clusters(wc, budget, ctx->cluster_weight, &wcl);
      for (size_t k = 0; k < wc.size(); ++k) cl[waves[wv][k]] = wcl[k];
      for (int c = P_CLMAX; c >= 1; --c) {
        Group g{(int)order.size(), 0, c, (int)wv};
        for (int s : waves[wv])  // (already in decreasing size)
          if (cl[s] == c) order.push_back(s);
        g.n = (int)order.size() - g.off;
        if (g.n > 0) {
          groups.push_back(g);
          max_cl = c > max_cl ? c : max_cl;
        }
      }
    }
    int* d_order = nullptr;
    unsigned short* d_tidx = nullptr;  // original index of every cell-sorted target point (tie-breaks, correspondences)
    float* d_lb = nullptr;             // streamed slices only: certified bounds between iterations
    CU(ctx, scr.alloc(&d_tidx, (size_t)S * max_cl * P_NTMAX));  // one replica per CTA of a cluster
    CU(ctx, scr.alloc(&d_lb, (size_t)S * wstride));
    CU(ctx, scr.alloc(&d_order, order.size()));
    if (!order.empty()) CU(ctx, small_h2d(ctx, d_order, order.data(), order.size() * sizeof(int)));
    const bool want_dbg = getenv("RSPCL_PERSIST_DBG") != nullptr;
    long long* d_dbg = nullptr;  // RSPCL_PERSIST_DBG=1: per-CTA phase cycle counters, printed to stderr
    long long* d_dbg_iter = nullptr;
    if (want_dbg) {
      CU(ctx, scr.alloc(&d_dbg, (size_t)S * max_cl * 8));
      CU(ctx, cudaMemsetAsync(d_dbg, 0, (size_t)S * max_cl * 8 * sizeof(long long), ctx->stream));
      CU(ctx, scr.alloc(&d_dbg_iter, (size_t)3 * 256));
      CU(ctx, cudaMemsetAsync(d_dbg_iter, 0, (size_t)3 * 256 * sizeof(long long), ctx->stream));
    }
    PersistArgs pa;
    pa.src = src->pts;
    pa.count = src->count;
    pa.sstride = src->stride;
    pa.work = work;
    pa.lb = d_lb;
    pa.wstride = wstride;
    pa.st1 = st;
    pa.st2 = st2;
    pa.n_stages = two_stage ? 2 : 1;
    pa.prev2 = nullptr;
    if (two_stage) {
      double* d_prev2 = nullptr;
      std::vector<double> p2(S);
      for (int s = 0; s < S; ++s) p2[s] = o.h_results2[s].prev_mse;
      CU(ctx, scr.alloc(&d_prev2, (size_t)S));
      CU(ctx, small_h2d(ctx, d_prev2, p2.data(), S * sizeof(double)));
      pa.prev2 = d_prev2;
    }
    pa.tgt = tgt->pts;
    pa.tcount = tgt->count;
    pa.tstride = tgt->stride;
    pa.shared_target = shared_target;
    pa.inv_cs = 1.0f / (float)(prm->max_corr_dist * 4.1);
    pa.corr_out = d_first_corr;
    pa.corr_iters = corr_iters;
    pa.n_pairs_total = S;
    pa.status = d_status;
    pa.tstart = d_status + S;
    pa.order = d_order;
    pa.tidx = d_tidx;
    pa.tidx_rep = max_cl;
    pa.dbg = d_dbg;
    pa.dbg_iter = d_dbg_iter;
    {
      ProfScope prof(ctx, "k_icp_persist", 0.0);
      // one launch per cluster size; the launches of a wave run side by side (largest clusters first: they are the
      // hardest to place) on the context stream and forked streams, waves follow one another
      int used = 0, cur_wave = -1;
      for (size_t gi = 0; gi < groups.size(); ++gi) {
        const int ng = groups[gi].n, gcl_i = groups[gi].cl;
        if (groups[gi].wave != cur_wave) {  // new wave: everything forked so far has been joined into the context stream
          cur_wave = groups[gi].wave;
          used = 0;
          const bool forks = gi + 1 < groups.size() && groups[gi + 1].wave == cur_wave;
          if (forks) CU(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
        }
        cudaStream_t strm = used == 0 ? ctx->stream : ctx->aux[used - 1];
        if (used > 0) CU(ctx, cudaStreamWaitEvent(strm, ctx->ev_fork, 0));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(gcl_i * ng));
        cfg.blockDim = dim3(P_THREADS);
        cfg.dynamicSmemBytes = sizeof(PersistSmem);
        cfg.stream = strm;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)gcl_i;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        PersistArgs pg = pa;
        pg.order = d_order + groups[gi].off;
        CU(ctx, cudaLaunchKernelEx(&cfg, persist_fn(want_dbg), pg, dp));
        LAUNCH_CHECK(ctx);
        if (used > 0) {
          CU(ctx, cudaEventRecord(ctx->ev_join[used - 1], strm));
          CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[used - 1], 0));
        }
        ++used;
      }
      prof.end();
      std::vector<int> hs((size_t)2 * S);
      CU(ctx, small_d2h(ctx, hs.data(), d_status, (size_t)2 * S * sizeof(int)));
      CU(ctx, small_d2h(ctx, hst.data(), st, (size_t)S * sizeof(IcpState)));
      if (two_stage) CU(ctx, small_d2h(ctx, hst2.data(), st2, (size_t)S * sizeof(IcpState)));
      CU(ctx, ctx_sync(ctx));
      if (waves.size() == 1 && (int)todo.size() <= ctx->sm_count && !budget_forced && todo.size() > 1) {
        // one-wave plan: every cluster should have started within microseconds of the first.  A cluster that started
        // hundreds of microseconds late was waiting for SMs -- the plan was too big for this chip's GPC layout.
        int lo = 0, hi = 0;
        const int t0 = hs[S + todo[0]];
        for (int s : todo) {
          int d = (hs[S + s] - t0) & 0x7fffffff;
          if (d > 0x3fffffff) d = d - 0x7fffffff - 1;  // (31-bit wrap)
          lo = d < lo ? d : lo;
          hi = d > hi ? d : hi;
        }
        if (hi - lo > 200) {  // ~0.2 ms; twice in a row (a single late start can be the host being slow to enqueue a group)
          if (++ctx->persist_late >= 2 && ctx->persist_extra > -2) {
            ctx->persist_extra -= 2;
            ctx->persist_late = 0;
          }
          ctx->persist_clean = 0;
        } else {
          ctx->persist_late = 0;
          // late starts also happen when another stream / context keeps SMs busy (pipelined callers): after a run of
          // clean launches the planner tries the larger plan again
          if (++ctx->persist_clean >= 32 && ctx->persist_extra < 2) {
            ctx->persist_extra += 2;
            ctx->persist_clean = 0;
          }
        }
      }
      if (want_dbg) {
        std::vector<long long> hd((size_t)S * max_cl * 8);
        CU(ctx, cudaMemcpyAsync(hd.data(), d_dbg, hd.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        std::vector<long long> hi3((size_t)3 * 256);
        CU(ctx, cudaMemcpyAsync(hi3.data(), d_dbg_iter, hi3.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        fprintf(stderr, "persist dbg: per-iteration trace of the first CTA(s) [iteration: re-queried points, phase-B kcycles, iteration kcycles]\n");
        for (int it = 0; it < 256 && hi3[3 * it + 2]; ++it)
          fprintf(stderr, "  it %3d: %6lld %7.1f %7.1f\n", it, hi3[3 * it], hi3[3 * it + 1] / 1e3, hi3[3 * it + 2] / 1e3);
        fprintf(stderr, "persist dbg: S=%d  [pair cl crank ns rescans iters | phaseA scan phaseB reduce+exchange solve head | total] kcycles\n", S);
        for (int s = 0; s < S; ++s)
          for (int c = 0; c < cl[s]; ++c) {
            const long long* D = &hd[((size_t)s * max_cl + c) * 8];
            fprintf(stderr, "  %3d %d %d %6d %6lld %3d | %7.1f %7.1f %7.1f %7.1f %7.1f %7.1f | %8.1f\n", s, cl[s], c, hcnt[s], D[6],
                    hst[s].iterations + (two_stage ? hst2[s].iterations : 0), D[0] / 1e3, D[1] / 1e3, D[2] / 1e3, D[3] / 1e3,
                    D[4] / 1e3, D[5] / 1e3, D[7] / 1e3);
          }
      }
      int n_failed = 0;
      double units = 0;
      for (int s = 0; s < S; ++s) {
        if (o.h_skip && o.h_skip[s]) continue;
        if (hs[s]) {
          ++n_failed;
        } else {
          finished[s] = 1;
          units += (double)hcnt[s] * ((hst[s].iterations > 0 ? hst[s].iterations : 1) + (two_stage ? (hst2[s].iterations > 0 ? hst2[s].iterations : 1) : 0));
        }
      }
      prof.set_units(units);
      // Targets that did not fit the shared-memory grid (too many points / occupied cells) left the kernel before touching
      // their state, so they are still freshly initialised; the finished pairs carry done = 1 and are skipped by every
      // kernel of the global-memory path, which now runs for the remaining ones only.
      active_init = n_failed;
      if (n_failed) CU(ctx, small_h2d(ctx, n_active, &active_init, sizeof(int)));
    }
  }
  const bool persist_done = want_persist && active_init == 0;
  if (!persist_done) {
    const int rcs = build_sorted_work();
    if (rcs) return rcs;
  }

  DevGrid g;
  g.shared_target = shared_target;
  struct GridGuard {
    rspcl_ctx* c;
    DevGrid* g;
    bool on = false;
    ~GridGuard() { if (on) grid_free(c, g); }
  } gguard{ctx, &g};
  if (!brute && !persist_done) {
    ProfScope prof(ctx, "grid_build", (double)S * tgt->max_count_hint);
    int rc = grid_build(ctx, tgt, cs, &g, d_range);
    if (rc) return rc;
    gguard.on = true;
  }

  // iteration loop: launches are enqueued in growing chunks; the host only looks at the active-pair counter
  // between chunks (1, 1, 2, 4, 8, ... iterations), so the reference's one-iteration aligns cost one read-back.
  dim3 gstep(nblk, S);
  int done_iters = 0, chunk = 1, active = persist_done ? 0 : active_init;
  if (use_cache && !persist_done) {
    const size_t np = (size_t)S * wstride;
    CU(ctx, scr.alloc(&g_rb, np));
    CU(ctx, scr.alloc(&g_wl, np));
    CU(ctx, scr.alloc(&g_wlcount, (size_t)S * nblk));
    CU(ctx, scr.alloc(&g_wloff, (size_t)S * (nblk + 1)));
  }
  double prof_units = 0;  // source points per launch (all pairs; converged pairs exit early)
  if (ctx->prof_on && !persist_done) {
    std::vector<int> c(S);
    CU(ctx, small_d2h(ctx, c.data(), src->count, S * sizeof(int)));
    CU(ctx, ctx_sync(ctx));
    for (int v : c) prof_units += v;
  }
  while (done_iters < prm->max_iterations && active > 0) {
    const int todo = (chunk < prm->max_iterations - done_iters) ? chunk : prm->max_iterations - done_iters;
    for (int k = 0; k < todo; ++k) {
      if (brute || !use_cache) {
        ProfScope prof(ctx, "k_icp_step", prof_units);
        if (brute)
          k_icp_step<true><<<gstep, IT, 0, ctx->stream>>>(work, src->count, wstride, st, g, dp, tgt->pts, tgt->count,
                                                         tgt->stride, partials, d_first_corr, corr_iters);
        else if (!use_cache)
          k_icp_step<false><<<gstep, IT, 0, ctx->stream>>>(work, src->count, wstride, st, g, dp, tgt->pts, tgt->count,
                                                          tgt->stride, partials, d_first_corr, corr_iters);
        LAUNCH_CHECK(ctx);
      }
      if (use_cache) {
        {
          ProfScope prof(ctx, (done_iters == 0 && k == 0) ? "k_icp_stream_first" : "k_icp_stream", prof_units);
          if (done_iters == 0 && k == 0)
            k_icp_stream<true><<<gstep, IT, 0, ctx->stream>>>(work, src->count, wstride, st, g, dp, g_rb, perm, pstride, g_wl, g_wlcount,
                                                              partials, d_first_corr, corr_iters);
          else
            k_icp_stream<false><<<gstep, IT, 0, ctx->stream>>>(work, src->count, wstride, st, g, dp, g_rb, perm, pstride, g_wl, g_wlcount,
                                                               partials, d_first_corr, corr_iters);
          LAUNCH_CHECK(ctx);
        }
        ProfScope prof(ctx, "k_icp_rescan", prof_units);
        k_wl_offsets<<<S, 1024, 0, ctx->stream>>>(g_wlcount, nblk, st, g_wloff);
        LAUNCH_CHECK(ctx);
        k_icp_rescan<<<gstep, IT, 0, ctx->stream>>>(work, src->count, wstride, st, g, dp, g_rb, perm, pstride, g_wl, g_wloff,
                                                    partials, d_first_corr, corr_iters);
        LAUNCH_CHECK(ctx);
      }
      ProfScope prof2(ctx, "k_icp_solve", (double)S);
      if (sharded && comm_peer_ready(ctx, S, NRED)) {
        // partial sums of this rank's shard -> one-shot exchange through peer memory, fused with the solve
        k_icp_solve_peer<<<S, SOLVE_T, 0, ctx->stream>>>(st, partials, nblk_total, dp, n_active, comm_peer_next(ctx));
      } else if (sharded) {
        // partial sums of this rank's shard -> ncclAllReduce -> identical solve on every rank
        k_icp_sum_partials<<<S, SOLVE_T, 0, ctx->stream>>>(partials, nblk_total, st, totals);
        LAUNCH_CHECK(ctx);
        int rcc;
        {
          ProfScope prof_ar(ctx, "allreduce", (double)S * NRED * sizeof(double));
          rcc = comm_allreduce_f64(ctx, totals, (size_t)S * NRED);
        }
        if (rcc) return rcc;
        k_icp_solve<<<S, SOLVE_T, 0, ctx->stream>>>(st, totals, 1, dp, n_active);
      } else {
        k_icp_solve<<<S, SOLVE_T, 0, ctx->stream>>>(st, partials, nblk_total, dp, n_active);
      }
      LAUNCH_CHECK(ctx);
    }
    done_iters += todo;
    CU(ctx, small_d2h(ctx, &active, n_active, sizeof(int)));
    CU(ctx, ctx_sync(ctx));
    if (done_iters > 1) chunk *= 2;
  }

  if (use_cache && !persist_done && getenv("RSPCL_ICP_DBG")) {  // debug: size of the last iteration's rescan list
    for (int sg = 0; sg < S && sg < 4; ++sg) {
      int tot = -1;
      CU(ctx, cudaMemcpyAsync(&tot, g_wloff + (size_t)sg * (nblk + 1) + nblk, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CU(ctx, cudaStreamSynchronize(ctx->stream));
      fprintf(stderr, "icp dbg: seg %d: %d points re-queried in the last iteration (nblk %d)\n", sg, tot, nblk);
    }
  }
  // results + aligned output (Registration::align: output = final applied to the original source)
  int range = 0;
  if (!persist_done) CU(ctx, small_d2h(ctx, hst.data(), st, S * sizeof(IcpState)));
  CU(ctx, small_d2h(ctx, &range, d_range, sizeof(int)));
  int rc = RSPCL_OK;
  // pairs of a two-stage call that the persistent kernel could not take: their first align has just run on the
  // global-memory path; the second one follows below as a separate align on the first one's output
  bool need_stage2 = false;
  if (two_stage)
    for (int s = 0; s < S; ++s) need_stage2 = need_stage2 || (!finished[s] && !(o.h_skip && o.h_skip[s]));
  TmpCloud mid(ctx);
  if (aligned || need_stage2) {
    k_gather_final<<<div_up(S * 16, 256), 256, 0, ctx->stream>>>(st, d_T, S);
    LAUNCH_CHECK(ctx);
    if (aligned) rc = transform_device(ctx, src, d_T, 0, aligned);
    if (!rc && need_stage2) {
      rc = mid.init(S, wstride);
      if (!rc) rc = transform_device(ctx, src, d_T, 0, &mid.c);
    }
  }
  CU(ctx, ctx_sync(ctx));
  if (rc) return rc;
  if (sharded) {
    rc = comm_peer_check(ctx);
    if (rc) return rc;
  }
  if (range) RSPCL_FAIL(ctx, RSPCL_ERR_RANGE, "icp_align: target coordinates exceed the grid key range (+-32767 cells of %g m)", cs);
  auto put = [](rspcl_icp_result& r, const IcpState& h) {
    memcpy(r.T, h.final_T, sizeof(float) * 16);
    r.converged = h.converged;
    r.state = h.state;
    r.iterations = h.iterations;
    r.n_corr = h.n_corr;
    r.mse = h.mse;
    r.prev_mse = h.prev_mse;
  };
  for (int s = 0; s < S; ++s) {
    if (o.h_skip && o.h_skip[s]) continue;
    put(h_results[s], hst[s]);
    if (two_stage && finished[s]) put(o.h_results2[s], hst2[s]);
  }
  if (need_stage2) {
    std::vector<rspcl_icp_result> r2(S);
    std::vector<unsigned char> skip2(S);
    for (int s = 0; s < S; ++s) {
      r2[s] = o.h_results2[s];  // prev_mse input
      skip2[s] = (finished[s] || (o.h_skip && o.h_skip[s])) ? 1 : 0;
    }
    IcpAlignOpts o2;
    o2.h_skip = skip2.data();
    o2.no_persist = true;  // these pairs did not fit the shared-memory kernel
    rc = icp_align_device(ctx, &mid.c, tgt, prm, nullptr, r2.data(), nullptr, o2);
    if (rc) return rc;
    for (int s = 0; s < S; ++s)
      if (!skip2[s]) o.h_results2[s] = r2[s];
  }
  scr.ok();
  return RSPCL_OK;
}

int refresh_count_hint(rspcl_ctx* ctx, const rspcl_cloud* c, std::vector<int>* counts_out) {
  std::vector<int> cnt(c->n_seg);
  CU(ctx, small_d2h(ctx, cnt.data(), c->count, c->n_seg * sizeof(int)));
  CU(ctx, ctx_sync(ctx));
  int m = 0;
  for (int v : cnt) m = v > m ? v : m;
  const_cast<rspcl_cloud*>(c)->max_count_hint = m;
  if (counts_out) counts_out->swap(cnt);
  return RSPCL_OK;
}

// shared body of rspcl_icp_align / rspcl_icp_align_dump: host_corr receives n_dump iterations of correspondences, each
// packed like a download of src
static int icp_align_host(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt, const rspcl_icp_params* prm,
                          const float* guess, rspcl_icp_result* results, rspcl_cloud* aligned, int n_dump, int32_t* host_corr) {
  if (!ctx || !src || !tgt || !prm || !results) return RSPCL_ERR_ARG;
  if (aligned == tgt) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "icp_align: aligned must not alias the target");
  if (host_corr && n_dump < 1) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "icp_align: n_dump_iterations < 1");
  CU(ctx, cudaSetDevice(ctx->device));
  std::vector<int> scnt;
  int rc = refresh_count_hint(ctx, src, &scnt);
  if (rc) return rc;
  rc = refresh_count_hint(ctx, tgt, nullptr);
  if (rc) return rc;
  const int S = src->n_seg;
  const size_t wstride = src->stride ? src->stride : 1;
  Scratch scr(ctx);
  float* d_guess = nullptr;
  int* d_fc = nullptr;
  if (guess) {
    CU(ctx, scr.alloc(&d_guess, (size_t)S * 16));
    CU(ctx, small_h2d(ctx, d_guess, guess, (size_t)S * 16 * sizeof(float)));
  }
  if (host_corr) {
    CU(ctx, scr.alloc(&d_fc, (size_t)n_dump * S * wstride));
    // iterations that are never executed (early convergence) read as -1
    CU(ctx, cudaMemsetAsync(d_fc, 0xFF, (size_t)n_dump * S * wstride * sizeof(int), ctx->stream));
  }
  IcpAlignOpts o;
  o.d_corr_out = d_fc;
  o.corr_iters = n_dump;
  o.h_src_counts = scnt.data();
  rc = icp_align_device(ctx, src, tgt, prm, d_guess, results, aligned, o);
  if (!rc && host_corr) {
    std::vector<int> off(S);
    long long total = 0;
    for (int s = 0; s < S; ++s) {
      off[s] = (int)total;
      total += scnt[s];
    }
    if (total) {
      int *d_off = nullptr, *packed = nullptr;
      CU(ctx, scr.alloc(&d_off, (size_t)S));
      CU(ctx, scr.alloc(&packed, (size_t)total * n_dump));
      CU(ctx, small_h2d(ctx, d_off, off.data(), S * sizeof(int)));
      dim3 grid(blocks_per_seg(ctx, S, src->max_count_hint, 256), S);
      for (int it = 0; it < n_dump; ++it) {
        k_pack_i32<<<grid, 256, 0, ctx->stream>>>(d_fc + (size_t)it * S * wstride, src->count, d_off, src->stride,
                                                  packed + (size_t)it * total);
        LAUNCH_CHECK(ctx);
      }
      CU(ctx, small_d2h(ctx, host_corr, packed, (size_t)total * n_dump * sizeof(int)));
      CU(ctx, ctx_sync(ctx));
    }
  }
  if (!rc) scr.ok();
  return rc;
}

extern "C" int rspcl_icp_align(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt, const rspcl_icp_params* prm,
                               const float* guess, rspcl_icp_result* results, rspcl_cloud* aligned, int32_t* first_corr) {
  return icp_align_host(ctx, src, tgt, prm, guess, results, aligned, 1, first_corr);
}

extern "C" int rspcl_icp_align_dump(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt, const rspcl_icp_params* prm,
                                    const float* guess, rspcl_icp_result* results, int n_dump_iterations, int32_t* host_corr) {
  if (!host_corr) return RSPCL_ERR_ARG;
  return icp_align_host(ctx, src, tgt, prm, guess, results, nullptr, n_dump_iterations, host_corr);
}

// Host-only: the cluster-size plan of the persistent ICP kernel for one wave of pairs (exported for the CPU test-suite;
// weights = SMs a cluster of c CTAs occupies, c = 1..8, nullptr: c itself).  Returns the resident-slice capacity.
extern "C" int rspcl_debug_plan_clusters(const int32_t* counts, int n, double budget, const double* weights, int32_t* out_cl) {
  if (!counts || !out_cl || n < 0) return -1;
  double w[P_CLMAX + 1];
  for (int c = 0; c <= P_CLMAX; ++c) w[c] = weights ? (c ? weights[c - 1] : 0.0) : (double)c;
  std::vector<int> cnt(counts, counts + n), cl;
  plan_clusters(cnt, budget, w, &cl);
  for (int i = 0; i < n; ++i) out_cl[i] = cl[i];
  return P_CHUNK;
}

extern "C" int rspcl_icp_align_sharded(rspcl_ctx* ctx, const rspcl_cloud* src_shard, const rspcl_cloud* tgt,
                                       const rspcl_icp_params* prm, const float* guess, rspcl_icp_result* results,
                                       rspcl_cloud* aligned) {
  if (!ctx) return RSPCL_ERR_ARG;
  ctx->sharded_call = true;
  const int rc = rspcl_icp_align(ctx, src_shard, tgt, prm, guess, results, aligned, nullptr);
  ctx->sharded_call = false;
  return rc;
}
