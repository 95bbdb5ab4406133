// orc_linalg.h -- small dense linear algebra for the CPU oracle (test infrastructure only).
// Stand-ins for the Eigen routines PCL calls (JacobiSVD, SelfAdjointEigenSolver, inverse); Eigen's
// exact rounding cannot be reproduced offline, so results are pinned against numpy in tests/.
#ifndef ORC_LINALG_H
#define ORC_LINALG_H
#include <cmath>
#include <algorithm>

namespace orc {

// Cyclic Jacobi eigen-decomposition of a symmetric NxN matrix (row-major).  On return w[] ascending,
// V columns = eigenvectors (V row-major, V[r*N+c] = component r of eigenvector c).
template <typename S, int N>
inline void jacobi_eigh(const S* A_in, S* w, S* V) {
  S A[N * N];
  for (int i = 0; i < N * N; ++i) A[i] = A_in[i];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) V[i * N + j] = (i == j) ? S(1) : S(0);
  for (int sweep = 0; sweep < 64; ++sweep) {
    S off = 0;
    for (int p = 0; p < N; ++p)
      for (int q = p + 1; q < N; ++q) off += A[p * N + q] * A[p * N + q];
    if (off == S(0)) break;
    for (int p = 0; p < N; ++p) {
      for (int q = p + 1; q < N; ++q) {
        S apq = A[p * N + q];
        if (apq == S(0)) continue;
        S app = A[p * N + p], aqq = A[q * N + q];
        // an off-diagonal entry that no longer registers against either diagonal entry is set to zero instead of being
        // rotated (Rutishauser's test); without it rounding noise keeps `off` from ever reaching 0 and all 64 sweeps run
        const S g = S(100) * std::fabs(apq);
        if (std::fabs(app) + g == std::fabs(app) && std::fabs(aqq) + g == std::fabs(aqq)) {
          A[p * N + q] = A[q * N + p] = S(0);
          continue;
        }
        S theta = (aqq - app) / (S(2) * apq);
        S t = (theta >= S(0) ? S(1) : S(-1)) / (std::fabs(theta) + std::sqrt(theta * theta + S(1)));
        S c = S(1) / std::sqrt(t * t + S(1));
        S s = t * c;
        for (int k = 0; k < N; ++k) {  // A <- A * J
          S akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {  // A <- J^T * A
          S apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < N; ++k) {
          S vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = c * vkp - s * vkq;
          V[k * N + q] = s * vkp + c * vkq;
        }
      }
    }
  }
  for (int i = 0; i < N; ++i) w[i] = A[i * N + i];
  // sort ascending (selection sort, swapping eigenvector columns)
  for (int i = 0; i < N - 1; ++i) {
    int m = i;
    for (int j = i + 1; j < N; ++j)
      if (w[j] < w[m]) m = j;
    if (m != i) {
      std::swap(w[i], w[m]);
      for (int k = 0; k < N; ++k) std::swap(V[k * N + i], V[k * N + m]);
    }
  }
}

template <typename S>
inline S det3(const S* M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

template <typename S>
inline void inv3(const S* M, S* I) {
  S c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
  S det = M[0] * c00 + M[1] * c01 + M[2] * c02;
  S id = S(1) / det;
  I[0] = c00 * id;
  I[1] = (M[2] * M[7] - M[1] * M[8]) * id;
  I[2] = (M[1] * M[5] - M[2] * M[4]) * id;
  I[3] = c01 * id;
  I[4] = (M[0] * M[8] - M[2] * M[6]) * id;
  I[5] = (M[2] * M[3] - M[0] * M[5]) * id;
  I[6] = c02 * id;
  I[7] = (M[1] * M[6] - M[0] * M[7]) * id;
  I[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}

// 3x3 SVD by one-sided (Hestenes) Jacobi: A = U diag(s) V^T, s descending, U,V orthogonal (row-major).
template <typename S>
inline void svd3(const S* A_in, S* U, S* s, S* V) {
  S B[9];
  for (int i = 0; i < 9; ++i) B[i] = A_in[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i * 3 + j] = (i == j) ? S(1) : S(0);
  const S eps = std::numeric_limits<S>::epsilon();
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 3; ++p) {
      for (int q = p + 1; q < 3; ++q) {
        S alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 3; ++k) {
          alpha += B[k * 3 + p] * B[k * 3 + p];
          beta += B[k * 3 + q] * B[k * 3 + q];
          gamma += B[k * 3 + p] * B[k * 3 + q];
        }
        if (gamma == S(0) || std::fabs(gamma) <= eps * std::sqrt(alpha * beta)) continue;
        rotated = true;
        S zeta = (beta - alpha) / (S(2) * gamma);
        S t = (zeta >= S(0) ? S(1) : S(-1)) / (std::fabs(zeta) + std::sqrt(S(1) + zeta * zeta));
        S c = S(1) / std::sqrt(S(1) + t * t);
        S sn = c * t;
        for (int k = 0; k < 3; ++k) {
          S bp = B[k * 3 + p], bq = B[k * 3 + q];
          B[k * 3 + p] = c * bp - sn * bq;
          B[k * 3 + q] = sn * bp + c * bq;
          S vp = V[k * 3 + p], vq = V[k * 3 + q];
          V[k * 3 + p] = c * vp - sn * vq;
          V[k * 3 + q] = sn * vp + c * vq;
        }
      }
    }
    if (!rotated) break;
  }
  S nrm[3];
  for (int j = 0; j < 3; ++j) nrm[j] = std::sqrt(B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j]);
  int ord[3] = {0, 1, 2};
  std::sort(ord, ord + 3, [&](int a, int b) { return nrm[a] > nrm[b]; });
  S Bs[9], Vs[9];
  for (int j = 0; j < 3; ++j) {
    s[j] = nrm[ord[j]];
    for (int k = 0; k < 3; ++k) {
      Bs[k * 3 + j] = B[k * 3 + ord[j]];
      Vs[k * 3 + j] = V[k * 3 + ord[j]];
    }
  }
  for (int i = 0; i < 9; ++i) V[i] = Vs[i];
  // U columns: normalised B columns; complete rank-deficient columns to an orthonormal basis
  const S tiny = s[0] * eps * S(8);
  int rank = 0;
  for (int j = 0; j < 3; ++j) {
    if (s[j] > tiny && s[j] > S(0)) {
      for (int k = 0; k < 3; ++k) U[k * 3 + j] = Bs[k * 3 + j] / s[j];
      rank = j + 1;
    } else {
      break;
    }
  }
  if (rank == 0) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) U[i * 3 + j] = (i == j) ? S(1) : S(0);
  } else if (rank == 1) {
    S u0[3] = {U[0], U[3], U[6]};
    int m = 0;
    if (std::fabs(u0[1]) < std::fabs(u0[m])) m = 1;
    if (std::fabs(u0[2]) < std::fabs(u0[m])) m = 2;
    S e[3] = {0, 0, 0};
    e[m] = 1;
    S u1[3] = {u0[1] * e[2] - u0[2] * e[1], u0[2] * e[0] - u0[0] * e[2], u0[0] * e[1] - u0[1] * e[0]};
    S n1 = std::sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    for (int k = 0; k < 3; ++k) u1[k] /= n1;
    S u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
    for (int k = 0; k < 3; ++k) {
      U[k * 3 + 1] = u1[k];
      U[k * 3 + 2] = u2[k];
    }
  } else if (rank == 2) {
    S u0[3] = {U[0], U[3], U[6]}, u1[3] = {U[1], U[4], U[7]};
    S u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
    S n2 = std::sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    for (int k = 0; k < 3; ++k) U[k * 3 + 2] = u2[k] / n2;
  }
}

}  // namespace orc
#endif
