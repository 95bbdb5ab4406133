// orc_ndt.cpp -- CPU oracle (test infrastructure only): pcl::NormalDistributionsTransform as the reference
// drives it (ndt:38-43,71-72,83,92,104).
//
// PCL 1.9.1 sources followed (not vendored -- PARITY UNPINNED, see orc.h):
//   registration/impl/ndt.hpp   computeTransformation, computeDerivatives, computeAngleDerivatives,
//                               computePointDerivatives, updateDerivatives, computeHessian, updateHessian,
//                               computeStepLengthMT, trialValueSelectionMT, updateIntervalMT
//   filters/impl/voxel_grid_covariance.hpp  applyFilter (single-pass covariance, eigenvalue inflation)
//   Magnusson 2009 eq. 6.8-6.21; More & Thuente 1994.
// The covariance normalisation line `cov *= (n-1)/n` is the 1.9.x form (confidence M).
#include "orc.h"
#include "orc_linalg.h"
#include <cmath>
#include <cfloat>
#include <climits>
#include <cstring>
#include <map>
#include <unordered_map>
#include <vector>
#include <algorithm>

extern "C" void orc_ndt_reference_params(OrcNdtParams* p) {
  p->max_iterations = 50;            // ndt:43
  p->transformation_epsilon = 0.01;  // ndt:39
  p->step_size = 0.1;                // ndt:40
  p->resolution = 1.0f;              // ndt:41
  p->outlier_ratio = 0.55;
  p->min_points_per_voxel = 6;
  p->min_covar_eigvalue_mult = 0.01;
}

extern "C" void orc_ndt_gauss_constants(float resolution, double outlier_ratio, double* d1, double* d2) {
  double gauss_c1 = 10.0 * (1 - outlier_ratio);
  double gauss_c2 = outlier_ratio / pow(double(resolution), 3);
  double gauss_d3 = -log(gauss_c2);
  *d1 = -log(gauss_c1 + gauss_c2) - gauss_d3;
  *d2 = -2 * log((-log(gauss_c1 * exp(-0.5) + gauss_c2) - gauss_d3) / *d1);
}

namespace {
inline int floor_to_int(float v) {
  float f = std::floor(v);
  if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT_MIN;
  return static_cast<int>(f);
}
struct Leaf {
  int ijk[3];
  int n = 0;
  double sum[3] = {0, 0, 0};
  double sxx[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  float csum[3] = {0, 0, 0};
};
struct Key {
  int x, y, z;
  bool operator<(const Key& o) const {  // std::map<size_t,Leaf> order of idx = i + j*dx + k*dx*dy
    if (z != o.z) return z < o.z;
    if (y != o.y) return y < o.y;
    return x < o.x;
  }
};
struct CellHash {
  size_t operator()(const Key& k) const {
    return size_t(uint32_t(k.x)) * 73856093u ^ size_t(uint32_t(k.y)) * 19349663u ^ size_t(uint32_t(k.z)) * 83492791u;
  }
};
struct KeyEq {
  bool operator()(const Key& a, const Key& b) const { return a.x == b.x && a.y == b.y && a.z == b.z; }
};
}  // namespace

struct OrcNdtGrid {
  std::vector<OrcNdtVoxel> vox;  // sorted by leaf index
  std::unordered_map<Key, int, CellHash, KeyEq> lookup;
  float inv_leaf;
  float resolution;
};

extern "C" OrcNdtGrid* orc_ndt_grid_build(const OrcPoint* tgt, int nt, const OrcNdtParams* prm) {
  OrcNdtGrid* g = new OrcNdtGrid;
  g->resolution = prm->resolution;
  g->inv_leaf = 1.0f / prm->resolution;
  std::map<Key, Leaf> leaves;
  for (int i = 0; i < nt; ++i) {
    const OrcPoint& p = tgt[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    Key k{floor_to_int(p.x * g->inv_leaf), floor_to_int(p.y * g->inv_leaf), floor_to_int(p.z * g->inv_leaf)};
    Leaf& L = leaves[k];
    L.ijk[0] = k.x;
    L.ijk[1] = k.y;
    L.ijk[2] = k.z;
    double v[3] = {p.x, p.y, p.z};
    for (int a = 0; a < 3; ++a) {
      L.sum[a] += v[a];
      L.csum[a] += float(v[a]);
      for (int b = 0; b < 3; ++b) L.sxx[a * 3 + b] += v[a] * v[b];
    }
    L.n++;
  }
  for (auto& kv : leaves) {
    Leaf& L = kv.second;
    if (L.n < prm->min_points_per_voxel) continue;
    OrcNdtVoxel V;
    memset(&V, 0, sizeof(V));
    for (int a = 0; a < 3; ++a) {
      V.ijk[a] = L.ijk[a];
      V.centroid[a] = L.csum[a] / float(L.n);
      V.mean[a] = L.sum[a] / L.n;
    }
    V.npts = L.n;
    // cov = (sxx - 2 * (pt_sum * mean^T)) / n + mean * mean^T ; cov *= (n - 1.0) / n
    double cov[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        cov[a * 3 + b] = (L.sxx[a * 3 + b] - 2 * (L.sum[a] * V.mean[b])) / L.n + V.mean[a] * V.mean[b];
    for (int k = 0; k < 9; ++k) cov[k] *= (L.n - 1.0) / L.n;
    // SelfAdjointEigenSolver reads the lower triangle only: symmetrise from it
    double sym[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) sym[a * 3 + b] = (a >= b) ? cov[a * 3 + b] : cov[b * 3 + a];
    double w[3], E[9];
    orc::jacobi_eigh<double, 3>(sym, w, E);
    for (int a = 0; a < 3; ++a) V.evals[a] = w[a];
    if (w[0] < 0 || w[1] < 0 || w[2] <= 0) {
      V.npts = -1;
      memcpy(V.cov, cov, sizeof(cov));
      g->vox.push_back(V);
      continue;
    }
    double min_ev = prm->min_covar_eigvalue_mult * w[2];
    if (w[0] < min_ev) {
      w[0] = min_ev;
      if (w[1] < min_ev) w[1] = min_ev;
      // cov = evecs * diag * evecs^-1 (orthonormal -> transpose)
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          double s = 0;
          for (int k = 0; k < 3; ++k) s += E[a * 3 + k] * w[k] * E[b * 3 + k];
          cov[a * 3 + b] = s;
        }
    }
    memcpy(V.cov, cov, sizeof(cov));
    orc::inv3<double>(cov, V.icov);
    bool bad = false;
    for (int k = 0; k < 9; ++k)
      if (!std::isfinite(V.icov[k])) bad = true;
    if (bad) V.npts = -1;
    g->vox.push_back(V);
  }
  for (size_t i = 0; i < g->vox.size(); ++i)
    g->lookup[Key{g->vox[i].ijk[0], g->vox[i].ijk[1], g->vox[i].ijk[2]}] = int(i);
  return g;
}

extern "C" void orc_ndt_grid_free(OrcNdtGrid* g) { delete g; }
extern "C" int orc_ndt_grid_size(const OrcNdtGrid* g) { return int(g->vox.size()); }
extern "C" void orc_ndt_grid_get(const OrcNdtGrid* g, OrcNdtVoxel* out) {
  memcpy(out, g->vox.data(), g->vox.size() * sizeof(OrcNdtVoxel));
}

namespace {

// Translation * AngleAxis(rx, X) * AngleAxis(ry, Y) * AngleAxis(rz, Z) in float (ndt.hpp)
void pose_to_matrix(const double p[6], float T[16]) {
  float rx = float(p[3]), ry = float(p[4]), rz = float(p[5]);
  float cx = std::cos(rx), sx = std::sin(rx), cy = std::cos(ry), sy = std::sin(ry), cz = std::cos(rz), sz = std::sin(rz);
  float Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
  float Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
  float Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  float A[9], R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      float s = 0;
      for (int k = 0; k < 3; ++k) s += Rx[r * 3 + k] * Ry[k * 3 + c];
      A[r * 3 + c] = s;
    }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      float s = 0;
      for (int k = 0; k < 3; ++k) s += A[r * 3 + k] * Rz[k * 3 + c];
      R[r * 3 + c] = s;
    }
  for (int i = 0; i < 16; ++i) T[i] = 0.f;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T[c * 4 + r] = R[r * 3 + c];
  T[12] = float(p[0]);
  T[13] = float(p[1]);
  T[14] = float(p[2]);
  T[15] = 1.f;
}

// Eigen 3.3 MatrixBase::eulerAngles(0,1,2) on the rotation block, float
void matrix_to_pose(const float T[16], double p[6]) {
  auto M = [&](int r, int c) { return T[c * 4 + r]; };
  const int i = 0, j = 1, k = 2;
  float res[3];
  res[0] = std::atan2(M(j, k), M(k, k));
  float c2 = std::sqrt(M(i, i) * M(i, i) + M(i, j) * M(i, j));
  if (res[0] > 0.f) {  // !odd && res[0] > 0
    res[0] -= float(M_PI);
    res[1] = std::atan2(-M(i, k), -c2);
  } else {
    res[1] = std::atan2(-M(i, k), c2);
  }
  float s1 = std::sin(res[0]), c1 = std::cos(res[0]);
  res[2] = std::atan2(s1 * M(k, i) - c1 * M(j, i), c1 * M(j, j) - s1 * M(k, j));
  p[0] = T[12];
  p[1] = T[13];
  p[2] = T[14];
  p[3] = -res[0];
  p[4] = -res[1];
  p[5] = -res[2];
}

struct AngleDeriv {
  double ja[3], jb[3], jc[3], jd[3], je[3], jf[3], jg[3], jh[3];
  double ha2[3], ha3[3], hb2[3], hb3[3], hc2[3], hc3[3], hd1[3], hd2[3], hd3[3], he1[3], he2[3], he3[3], hf1[3], hf2[3],
      hf3[3];
};

void angle_derivatives(const double p[6], AngleDeriv& A) {
  double cx, cy, cz, sx, sy, sz;
  if (fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = cos(p[3]); sx = sin(p[3]); }
  if (fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = cos(p[4]); sy = sin(p[4]); }
  if (fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = cos(p[5]); sz = sin(p[5]); }
  auto set = [](double* v, double a, double b, double c) { v[0] = a; v[1] = b; v[2] = c; };
  set(A.ja, (-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy));
  set(A.jb, (cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy));
  set(A.jc, (-sy * cz), sy * sz, cy);
  set(A.jd, sx * cy * cz, (-sx * cy * sz), sx * sy);
  set(A.je, (-cx * cy * cz), cx * cy * sz, (-cx * sy));
  set(A.jf, (-cy * sz), (-cy * cz), 0);
  set(A.jg, (cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0);
  set(A.jh, (sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0);
  set(A.ha2, (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy);
  set(A.ha3, (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy));
  set(A.hb2, (cx * cy * cz), (-cx * cy * sz), (cx * sy));
  set(A.hb3, (sx * cy * cz), (-sx * cy * sz), (sx * sy));
  set(A.hc2, (-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0);
  set(A.hc3, (cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0);
  set(A.hd1, (-cy * cz), (cy * sz), (sy));  // PCL literal (+sy); the exact second derivative is -sy
  set(A.hd2, (-sx * sy * cz), (sx * sy * sz), (sx * cy));
  set(A.hd3, (cx * sy * cz), (-cx * sy * sz), (-cx * cy));
  set(A.he1, (sy * sz), (sy * cz), 0);
  set(A.he2, (-sx * cy * sz), (-sx * cy * cz), 0);
  set(A.he3, (cx * cy * sz), (cx * cy * cz), 0);
  set(A.hf1, (-cy * cz), (cy * sz), 0);
  set(A.hf2, (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0);
  set(A.hf3, (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0);
}

inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

struct NdtCtx {
  const OrcNdtGrid* grid;
  const OrcPoint* src;
  int ns;
  double d1, d2;
  long long n_pairs;
};

// computeDerivatives (compute_hessian) / computeHessian (hessian_only: score & gradient untouched)
double derivatives(NdtCtx& C, const std::vector<OrcPoint>& trans, const double p[6], double g[6], double H[36],
                   bool compute_hessian, bool hessian_only) {
  AngleDeriv A;
  angle_derivatives(p, A);
  if (!hessian_only)
    for (int i = 0; i < 6; ++i) g[i] = 0;
  if (compute_hessian)
    for (int i = 0; i < 36; ++i) H[i] = 0;
  double score = 0;
  C.n_pairs = 0;
  const OrcNdtGrid* G = C.grid;
  const float r2 = float(double(G->resolution) * double(G->resolution));
  for (int idx = 0; idx < C.ns; ++idx) {
    const OrcPoint& xt = trans[idx];
    if (!std::isfinite(xt.x) || !std::isfinite(xt.y) || !std::isfinite(xt.z)) continue;
    int cx = floor_to_int(xt.x * G->inv_leaf), cy = floor_to_int(xt.y * G->inv_leaf), cz = floor_to_int(xt.z * G->inv_leaf);
    // radiusSearch(x_trans, resolution) over voxel centroids == 27-cell scan + strict float distance test
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          auto it = G->lookup.find(Key{cx + dx, cy + dy, cz + dz});
          if (it == G->lookup.end()) continue;
          const OrcNdtVoxel& V = G->vox[it->second];
          float ddx = xt.x - V.centroid[0], ddy = xt.y - V.centroid[1], ddz = xt.z - V.centroid[2];
          float dist = 0.f;
          dist += ddx * ddx;
          dist += ddy * ddy;
          dist += ddz * ddz;
          if (!(dist < r2)) continue;
          C.n_pairs++;
          double x[3] = {C.src[idx].x, C.src[idx].y, C.src[idx].z};
          double xtr[3] = {double(xt.x) - V.mean[0], double(xt.y) - V.mean[1], double(xt.z) - V.mean[2]};
          const double* ci = V.icov;
          // point_gradient_ (3x6): identity | angular columns
          double J[3][6] = {{1, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0}};
          J[1][3] = dot3(x, A.ja);
          J[2][3] = dot3(x, A.jb);
          J[0][4] = dot3(x, A.jc);
          J[1][4] = dot3(x, A.jd);
          J[2][4] = dot3(x, A.je);
          J[0][5] = dot3(x, A.jf);
          J[1][5] = dot3(x, A.jg);
          J[2][5] = dot3(x, A.jh);
          double PH[18][6];
          if (compute_hessian) {
            memset(PH, 0, sizeof(PH));
            double a[3] = {0, dot3(x, A.ha2), dot3(x, A.ha3)};
            double b[3] = {0, dot3(x, A.hb2), dot3(x, A.hb3)};
            double c[3] = {0, dot3(x, A.hc2), dot3(x, A.hc3)};
            double d[3] = {dot3(x, A.hd1), dot3(x, A.hd2), dot3(x, A.hd3)};
            double e[3] = {dot3(x, A.he1), dot3(x, A.he2), dot3(x, A.he3)};
            double f[3] = {dot3(x, A.hf1), dot3(x, A.hf2), dot3(x, A.hf3)};
            for (int r = 0; r < 3; ++r) {
              PH[9 + r][3] = a[r];
              PH[12 + r][3] = b[r];
              PH[15 + r][3] = c[r];
              PH[9 + r][4] = b[r];
              PH[12 + r][4] = d[r];
              PH[15 + r][4] = e[r];
              PH[9 + r][5] = c[r];
              PH[12 + r][5] = e[r];
              PH[15 + r][5] = f[r];
            }
          }
          // updateDerivatives
          double cix[3] = {ci[0] * xtr[0] + ci[1] * xtr[1] + ci[2] * xtr[2], ci[3] * xtr[0] + ci[4] * xtr[1] + ci[5] * xtr[2],
                           ci[6] * xtr[0] + ci[7] * xtr[1] + ci[8] * xtr[2]};
          double e_x_cov_x = exp(-C.d2 * dot3(xtr, cix) / 2);
          double score_inc = -C.d1 * e_x_cov_x;
          e_x_cov_x = C.d2 * e_x_cov_x;
          if (e_x_cov_x > 1 || e_x_cov_x < 0 || e_x_cov_x != e_x_cov_x) continue;  // returns 0: no score either
          e_x_cov_x *= C.d1;
          double cJ[6][3];  // c_inv * J.col(i)
          double xcJ[6];    // x_trans . (c_inv * J.col(i))
          for (int i = 0; i < 6; ++i) {
            for (int r = 0; r < 3; ++r) cJ[i][r] = ci[r * 3 + 0] * J[0][i] + ci[r * 3 + 1] * J[1][i] + ci[r * 3 + 2] * J[2][i];
            xcJ[i] = dot3(xtr, cJ[i]);
          }
          for (int i = 0; i < 6; ++i) {
            if (!hessian_only) g[i] += xcJ[i] * e_x_cov_x;
            if (compute_hessian) {
              for (int j = 0; j < 6; ++j) {
                double ph[3] = {PH[3 * i + 0][j], PH[3 * i + 1][j], PH[3 * i + 2][j]};
                double cph[3] = {ci[0] * ph[0] + ci[1] * ph[1] + ci[2] * ph[2], ci[3] * ph[0] + ci[4] * ph[1] + ci[5] * ph[2],
                                 ci[6] * ph[0] + ci[7] * ph[1] + ci[8] * ph[2]};
                double Jj[3] = {J[0][j], J[1][j], J[2][j]};
                H[i * 6 + j] += e_x_cov_x * (-C.d2 * xcJ[i] * xcJ[j] + dot3(xtr, cph) + dot3(Jj, cJ[i]));
              }
            }
          }
          if (!hessian_only) score += score_inc;
        }
  }
  return score;
}

// JacobiSVD(H).solve(-g): least-squares solve through the symmetric eigen-decomposition (H is symmetric
// up to rounding); singular values below 6*eps*max are treated as zero like Eigen's default threshold.
void solve_newton(const double H[36], const double g[6], double dp[6]) {
  double S[36], w[6], V[36];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) S[i * 6 + j] = 0.5 * (H[i * 6 + j] + H[j * 6 + i]);
  orc::jacobi_eigh<double, 6>(S, w, V);
  double wmax = 0;
  for (int i = 0; i < 6; ++i) wmax = std::max(wmax, std::fabs(w[i]));
  double thr = 6 * DBL_EPSILON * wmax;
  for (int i = 0; i < 6; ++i) dp[i] = 0;
  for (int k = 0; k < 6; ++k) {
    if (!(std::fabs(w[k]) > thr)) continue;
    double c = 0;
    for (int i = 0; i < 6; ++i) c += V[i * 6 + k] * (-g[i]);
    c /= w[k];
    for (int i = 0; i < 6; ++i) dp[i] += V[i * 6 + k] * c;
  }
}

inline double psiMT(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
inline double dpsiMT(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

bool updateIntervalMT(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t,
                      double f_t, double g_t) {
  if (f_t > f_l) {
    a_u = a_t; f_u = f_t; g_u = g_t;
    return false;
  } else if (g_t * (a_l - a_t) > 0) {
    a_l = a_t; f_l = f_t; g_l = g_t;
    return false;
  } else if (g_t * (a_l - a_t) < 0) {
    a_u = a_l; f_u = f_l; g_u = g_l;
    a_l = a_t; f_l = f_t; g_l = g_t;
    return false;
  }
  return true;
}

double trialValueSelectionMT(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t,
                             double f_t, double g_t) {
  if (f_t > f_l) {  // Case 1
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    if (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) return a_c;
    return 0.5 * (a_q + a_c);
  } else if (g_t * g_l < 0) {  // Case 2
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    if (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) return a_c;
    return a_s;
  } else if (std::fabs(g_t) <= std::fabs(g_l)) {  // Case 3
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    double a_t_next = (std::fabs(a_c - a_t) < std::fabs(a_s - a_t)) ? a_c : a_s;
    if (a_t > a_l) return std::min(a_t + 0.66 * (a_u - a_t), a_t_next);
    return std::max(a_t + 0.66 * (a_u - a_t), a_t_next);
  } else {  // Case 4
    double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
    double w = std::sqrt(z * z - g_t * g_u);
    return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
  }
}

struct NdtState {
  NdtCtx C;
  std::vector<OrcPoint> trans;
  float final_T[16];
  int n_deriv = 0, n_hess = 0;
};

double stepLengthMT(NdtState& S, const double x[6], double step_dir[6], double step_init, double step_max,
                    double step_min, double& score, double g[6], double H[36]) {
  double phi_0 = -score;
  double d_phi_0 = 0;
  for (int i = 0; i < 6; ++i) d_phi_0 -= g[i] * step_dir[i];
  double x_t[6];
  if (d_phi_0 >= 0) {
    if (d_phi_0 == 0) return 0;
    d_phi_0 *= -1;
    for (int i = 0; i < 6; ++i) step_dir[i] *= -1;
  }
  const int max_step_iterations = 10;
  int step_iterations = 0;
  const double mu = 1.e-4, nu = 0.9;
  double a_l = 0, a_u = 0;
  double f_l = psiMT(a_l, phi_0, phi_0, d_phi_0, mu);
  double g_l = dpsiMT(d_phi_0, d_phi_0, mu);
  double f_u = psiMT(a_u, phi_0, phi_0, d_phi_0, mu);
  double g_u = dpsiMT(d_phi_0, d_phi_0, mu);
  bool interval_converged = (step_max - step_min) < 0, open_interval = true;
  double a_t = step_init;
  a_t = std::min(a_t, step_max);
  a_t = std::max(a_t, step_min);
  for (int i = 0; i < 6; ++i) x_t[i] = x[i] + step_dir[i] * a_t;
  pose_to_matrix(x_t, S.final_T);
  orc_transform(S.C.src, S.C.ns, S.final_T, S.trans.data());
  score = derivatives(S.C, S.trans, x_t, g, H, true, false);
  S.n_deriv++;
  S.n_hess++;
  double phi_t = -score;
  double d_phi_t = 0;
  for (int i = 0; i < 6; ++i) d_phi_t -= g[i] * step_dir[i];
  double psi_t = psiMT(a_t, phi_t, phi_0, d_phi_0, mu);
  double d_psi_t = dpsiMT(d_phi_t, d_phi_0, mu);
  while (!interval_converged && step_iterations < max_step_iterations && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
    if (open_interval)
      a_t = trialValueSelectionMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
    else
      a_t = trialValueSelectionMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
    a_t = std::min(a_t, step_max);
    a_t = std::max(a_t, step_min);
    for (int i = 0; i < 6; ++i) x_t[i] = x[i] + step_dir[i] * a_t;
    pose_to_matrix(x_t, S.final_T);
    orc_transform(S.C.src, S.C.ns, S.final_T, S.trans.data());
    score = derivatives(S.C, S.trans, x_t, g, H, false, false);
    S.n_deriv++;
    phi_t = -score;
    d_phi_t = 0;
    for (int i = 0; i < 6; ++i) d_phi_t -= g[i] * step_dir[i];
    psi_t = psiMT(a_t, phi_t, phi_0, d_phi_0, mu);
    d_psi_t = dpsiMT(d_phi_t, d_phi_0, mu);
    if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
      open_interval = false;
      f_l = f_l + phi_0 - mu * d_phi_0 * a_l;
      g_l = g_l + mu * d_phi_0;
      f_u = f_u + phi_0 - mu * d_phi_0 * a_u;
      g_u = g_u + mu * d_phi_0;
    }
    if (open_interval)
      interval_converged = updateIntervalMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
    else
      interval_converged = updateIntervalMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
    step_iterations++;
  }
  if (step_iterations) {
    double gdummy[6];
    derivatives(S.C, S.trans, x_t, gdummy, H, true, true);
    S.n_hess++;
  }
  return a_t;
}

}  // namespace

extern "C" void orc_pose_to_matrix(const double p[6], float T[16]) { pose_to_matrix(p, T); }
extern "C" void orc_matrix_to_pose(const float T[16], double p[6]) { matrix_to_pose(T, p); }

extern "C" double orc_ndt_derivatives(const OrcNdtGrid* grid, const OrcPoint* src, int ns, const OrcNdtParams* prm,
                                      const double p[6], double g[6], double H[36], int compute_hessian,
                                      long long* n_pairs) {
  NdtCtx C;
  C.grid = grid;
  C.src = src;
  C.ns = ns;
  orc_ndt_gauss_constants(prm->resolution, prm->outlier_ratio, &C.d1, &C.d2);
  float T[16];
  pose_to_matrix(p, T);
  std::vector<OrcPoint> trans(ns);
  orc_transform(src, ns, T, trans.data());
  double score = derivatives(C, trans, p, g, H, compute_hessian != 0, false);
  if (n_pairs) *n_pairs = C.n_pairs;
  return score;
}

extern "C" void orc_ndt_align(const OrcPoint* src, int ns, const OrcPoint* tgt, int nt, const OrcNdtParams* prm,
                              const float guess[16], OrcNdtResult* res, OrcPoint* aligned) {
  OrcNdtGrid* grid = orc_ndt_grid_build(tgt, nt, prm);
  NdtState S;
  S.C.grid = grid;
  S.C.src = src;
  S.C.ns = ns;
  orc_ndt_gauss_constants(prm->resolution, prm->outlier_ratio, &S.C.d1, &S.C.d2);
  S.trans.assign(src, src + ns);
  float ident[16];
  for (int i = 0; i < 16; ++i) ident[i] = (i % 5 == 0) ? 1.f : 0.f;
  const float* gs = guess ? guess : ident;
  memcpy(S.final_T, gs, sizeof(S.final_T));
  bool is_ident = true;
  for (int i = 0; i < 16; ++i)
    if (gs[i] != ident[i]) is_ident = false;
  if (!is_ident) orc_transform(src, ns, gs, S.trans.data());

  double p[6], delta_p[6], g[6], H[36];
  matrix_to_pose(S.final_T, p);
  double score = derivatives(S.C, S.trans, p, g, H, true, false);
  S.n_deriv++;
  S.n_hess++;
  int nr_iterations = 0;
  bool converged = false;
  while (!converged) {
    solve_newton(H, g, delta_p);
    double nrm = 0;
    for (int i = 0; i < 6; ++i) nrm += delta_p[i] * delta_p[i];
    nrm = std::sqrt(nrm);
    if (nrm == 0 || nrm != nrm) {
      converged = (nrm == nrm);
      break;
    }
    for (int i = 0; i < 6; ++i) delta_p[i] /= nrm;
    nrm = stepLengthMT(S, p, delta_p, nrm, prm->step_size, prm->transformation_epsilon / 2, score, g, H);
    for (int i = 0; i < 6; ++i) {
      delta_p[i] *= nrm;
      p[i] += delta_p[i];
    }
    if (nr_iterations > prm->max_iterations || (nr_iterations && (std::fabs(nrm) < prm->transformation_epsilon)))
      converged = true;
    nr_iterations++;
  }
  memcpy(res->T, S.final_T, sizeof(S.final_T));
  res->converged = converged ? 1 : 0;
  res->iterations = nr_iterations;
  res->score = score;
  res->trans_probability = score / double(ns);
  memcpy(res->p, p, sizeof(p));
  res->n_derivative_evals = S.n_deriv;
  res->n_hessian_evals = S.n_hess;
  if (aligned) memcpy(aligned, S.trans.data(), sizeof(OrcPoint) * ns);
  orc_ndt_grid_free(grid);
}
