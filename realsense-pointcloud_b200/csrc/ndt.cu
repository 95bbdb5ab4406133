// ndt.cu -- pcl::NormalDistributionsTransform<PointXYZRGB,PointXYZRGB>::align for a batch of independent pairs.
//
// Reference call sites: ndt_edge_based_registration.hpp:38-43 (parameters), :71-72 (setInputSource/Target),
// :83,92 (align with the initial guess), :104 (getFinalTransformation).
// Kernels:
//   K6  voxel Gaussians (VoxelGridCovariance::filter): radix sort of the target by (segment, iz, iy, ix) (sort.cu),
//       run heads -> cells, one thread per voxel accumulates sum p / sum p p^T in fp64 IN INPUT ORDER (bit-identical
//       to a sequential pass), then mean, covariance (PCL 1.9 normalisation), Jacobi eigen-decomposition, eigenvalue
//       inflation, inverse; accepted voxels go into an open-addressing hash keyed by the cell.
//   K7  k_ndt_eval: fused transform + 27-cell neighbourhood (radiusSearch(resolution) over voxel centroids) +
//       score / 6-gradient / 21-entry symmetric Hessian accumulation in fp64, reduced per CTA with warp shuffles;
//       CTA partials are combined in a fixed order (deterministic).
//   K7b k_ndt_control: one warp per pair -- Newton step (6x6 symmetric Jacobi pseudo-inverse, i.e. JacobiSVD::solve
//       on a symmetric matrix) and the More-Thuente line search as a device-resident state machine; the host only
//       polls an active-pair counter.
// Roofline: K7 reads 16 B per source point and is bounded by the FP64 pipe (hundreds of fp64 FMAs per
// (point, voxel) pair), not by HBM; both are reported by bench.py.
#include "grid.cuh"
#include <cooperative_groups.h>
#include <float.h>
#include <math.h>
namespace cg = cooperative_groups;

int radix_sort_pairs(rspcl_ctx* ctx, unsigned long long* keys, int* vals, unsigned long long* tmp_keys, int* tmp_vals,
                     long long n);
int refresh_count_hint(rspcl_ctx* ctx, const rspcl_cloud* c, std::vector<int>* counts_out);

namespace {

struct VoxRec {  // == oracle OrcNdtVoxel == rspcl_ndt_voxels record (224 B)
  int ijk[3];
  int npts;
  float centroid[3];
  int pad;
  double mean[3];
  double cov[9];
  double icov[9];
  double evals[3];
};
static_assert(sizeof(VoxRec) == 224, "voxel record layout");

constexpr unsigned long long KEY_INVALID = 0xFFFFFFFFFFFFFFFFull;
constexpr int NT = 128;     // threads per CTA in k_ndt_eval
constexpr int NACC = 28;    // score, g[6], H upper triangle [21]

__device__ __forceinline__ unsigned long long ndt_key(int seg, int ix, int iy, int iz) {
  // (seg, iz, iy, ix) major -> minor: sorted order == std::map<leaf index> order of VoxelGridCovariance
  return ((unsigned long long)(unsigned)seg << 48) | ((unsigned long long)(unsigned)(iz + 32768) << 32) |
         ((unsigned long long)(unsigned)(iy + 32768) << 16) | (unsigned long long)(unsigned)(ix + 32768);
}

__global__ void k_ndt_keys(const float4* __restrict__ pts, const int* __restrict__ count, int stride, int pstride,
                           float inv_leaf, unsigned long long* __restrict__ keys, int* __restrict__ vals,
                           int* __restrict__ range_flag) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pstride; i += gridDim.x * blockDim.x) {
    unsigned long long k = KEY_INVALID;
    if (i < n) {
      const float4 p = pts[(size_t)seg * stride + i];
      if (finite3(p.x, p.y, p.z)) {
        const int ix = floor_to_int_x86(fmul(p.x, inv_leaf)), iy = floor_to_int_x86(fmul(p.y, inv_leaf)),
                  iz = floor_to_int_x86(fmul(p.z, inv_leaf));
        if (grid_in_range(ix, iy, iz))
          k = ndt_key(seg, ix, iy, iz);
        else
          atomicExch(range_flag, 1);
      }
    }
    keys[(size_t)seg * pstride + i] = k;
    vals[(size_t)seg * pstride + i] = i;
  }
}

__global__ void k_cell_heads(const unsigned long long* __restrict__ keys, long long n, int* __restrict__ flags,
                             int* __restrict__ n_valid) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[j];
    const bool valid = k != KEY_INVALID;
    flags[j] = (valid && (j == 0 || keys[j - 1] != k)) ? 1 : 0;
    if (valid && (j == n - 1 || keys[j + 1] == KEY_INVALID)) *n_valid = (int)(j + 1);
  }
}

__global__ void k_cell_starts(const int* __restrict__ flags, const int* __restrict__ ord, long long n,
                              const int* __restrict__ n_cells, const int* __restrict__ n_valid, int* __restrict__ cell_start) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x)
    if (flags[j]) cell_start[ord[j]] = (int)j;
  if (blockIdx.x == 0 && threadIdx.x == 0) cell_start[*n_cells] = *n_valid;
}

__global__ void k_leaf_flags(const int* __restrict__ cell_start, const int* __restrict__ n_cells, int min_points,
                             int* __restrict__ leaf_flag, long long cap) {
  const int nc = *n_cells;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cap; c += (long long)gridDim.x * blockDim.x)
    leaf_flag[c] = (c < nc && cell_start[c + 1] - cell_start[c] >= min_points) ? 1 : 0;
}

// ---- one-CTA voxel build for a single target of <= NB1_MAX points (configs[0] / configs[2] / the facade's per-frame aligns).
// Replaces ~35 launches (keys, 8-bit radix passes over 64-bit keys, two scans, heads, starts, leaf flags) with ONE launch
// that leaves the very same arrays behind for k_voxel_stats: sorted point indices, cell starts, leaf flags / ordinals,
// and the 64-bit key at every cell head.  Keys are packed TIGHTLY inside the occupied cell box ((iz,iy,ix) major -> minor,
// the order of ndt_key), so a 5 x 3 x 5 m scene at 1 m leaves sorts in 2-3 four-bit passes.  Stable LSD radix sort with a
// contiguous chunk per thread: per-thread digit counts in shared memory, one block scan per pass, ordered scatter.
constexpr int NB1_T = 1024;
constexpr int NB1_MAX = 32 * NB1_T;   // <= 32 items per thread (keys / indices ping-pong in global memory, L1/L2-resident)
constexpr int NB1_SMEM_MAX = 16384;   // up to here the sort runs entirely in shared memory (32-bit keys, 16-bit indices)

struct Nb1Smem {
  unsigned short hist[16][NB1_T];  // per-thread digit counts -> scatter bases (<= 32768 fits)
  int wsum[32];
  int box[6];
  int tot;
  int pad;
};
constexpr size_t NB1_SMEM_BYTES = sizeof(Nb1Smem) + (size_t)NB1_SMEM_MAX * 12;

__device__ __forceinline__ int nb1_block_excl_scan(int v, Nb1Smem& S, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // wsum free
  if (lane == 31) S.wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = S.wsum[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    S.wsum[lane] = winc - w;
    if (lane == 31) S.tot = winc;
  }
  __syncthreads();
  if (total) *total = S.tot;
  return S.wsum[warp] + inc - v;
}

__device__ __forceinline__ int bits_for(int range) {  // bits to hold 0..range
  return range <= 0 ? 0 : 32 - __clz(range);
}

// the phases after the keys exist; K / V = the key / index types of the storage the sort runs in
template <typename K, typename V>
__device__ __forceinline__ void nb1_sort_emit(K* ks, K* kd, V* vs, V* vd, Nb1Smem& S, int b0, int b1, int kbits, int n_valid,
                                              int bx, int by, int mx, int my, int mz, int min_points,
                                              unsigned long long* __restrict__ keys_out, int* __restrict__ vals_out,
                                              int* __restrict__ cell_start, int* __restrict__ leaf_flag,
                                              int* __restrict__ leaf_ord, int* __restrict__ n_cells_out,
                                              int* __restrict__ n_leaves_out) {
  const int t = threadIdx.x;
  // ---- stable LSD radix sort, 4 bits per pass, over kbits + 1 bits (the extra bit sorts the invalid keys last)
  for (int shift = 0; shift <= kbits; shift += 4) {
#pragma unroll
    for (int d = 0; d < 16; ++d) S.hist[d][t] = 0;
    for (int i = b0; i < b1; ++i) S.hist[(unsigned)(ks[i] >> shift) & 15][t] += 1;  // column t is private to the thread
    __syncthreads();
    // exclusive scan of the (digit-major, thread-minor) counts: thread t owns 16 consecutive entries (two 128-bit words)
    uint4* flat = reinterpret_cast<uint4*>(&S.hist[0][0]) + 2 * t;
    uint4 q0 = flat[0], q1 = flat[1];
    unsigned w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    int sum = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) sum += (int)(w[q] & 0xFFFFu) + (int)(w[q] >> 16);
    int run = nb1_block_excl_scan(sum, S, nullptr);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c0 = (int)(w[q] & 0xFFFFu), c1 = (int)(w[q] >> 16);
      w[q] = (unsigned)run | ((unsigned)(run + c0) << 16);
      run += c0 + c1;
    }
    flat[0] = make_uint4(w[0], w[1], w[2], w[3]);
    flat[1] = make_uint4(w[4], w[5], w[6], w[7]);
    __syncthreads();
    for (int i = b0; i < b1; ++i) {
      const K k = ks[i];
      const unsigned pos = S.hist[(unsigned)(k >> shift) & 15][t]++;
      kd[pos] = k;
      vd[pos] = vs[i];
    }
    __syncthreads();
    K* tk = ks; ks = kd; kd = tk;
    V* tv = vs; vs = vd; vd = tv;
  }
  // ---- cell heads -> cell ids, cell starts, 64-bit keys at the heads, sorted point indices
  int heads = 0;
  for (int j = b0; j < b1; ++j) heads += (j < n_valid && (j == 0 || ks[j - 1] != ks[j])) ? 1 : 0;
  int n_cells;
  int cell = nb1_block_excl_scan(heads, S, &n_cells);
  for (int j = b0; j < b1; ++j) {
    vals_out[j] = (int)vs[j];
    if (j < n_valid && (j == 0 || ks[j - 1] != ks[j])) {
      const unsigned long long k = (unsigned long long)ks[j];
      const int ix = (int)(k & ((1ull << bx) - 1)) + mx, iy = (int)((k >> bx) & ((1ull << by) - 1)) + my,
                iz = (int)(k >> (bx + by)) + mz;
      keys_out[j] = ndt_key(0, ix, iy, iz);
      cell_start[cell++] = j;
    }
  }
  if (t == 0) cell_start[n_cells] = n_valid, *n_cells_out = n_cells;
  __syncthreads();
  // ---- leaves: cells with >= min_points points, numbered in key order
  const int Cc = (n_cells + NB1_T - 1) / NB1_T;
  const int c0 = min(t * Cc, n_cells), c1 = min(c0 + Cc, n_cells);
  int leaves = 0;
  for (int c = c0; c < c1; ++c) {
    const int f = (cell_start[c + 1] - cell_start[c] >= min_points) ? 1 : 0;
    leaf_flag[c] = f;
    leaves += f;
  }
  int n_leaves;
  int ord = nb1_block_excl_scan(leaves, S, &n_leaves);
  for (int c = c0; c < c1; ++c) {
    leaf_ord[c] = ord;
    ord += leaf_flag[c];
  }
  if (t == 0) *n_leaves_out = n_leaves;
}

__global__ void __launch_bounds__(NB1_T, 1)
k_ndt_build_one(const float4* __restrict__ pts, const int* __restrict__ count, int N, float inv_leaf, int min_points,
                unsigned long long* __restrict__ kA, unsigned long long* __restrict__ kB, int* __restrict__ vA,
                int* __restrict__ vB, unsigned long long* __restrict__ keys_out, int* __restrict__ vals_out,
                int* __restrict__ cell_start, int* __restrict__ leaf_flag, int* __restrict__ leaf_ord,
                int* __restrict__ n_cells_out, int* __restrict__ n_leaves_out, int* __restrict__ range_flag) {
  extern __shared__ __align__(16) unsigned char nb1_raw[];
  Nb1Smem& S = *reinterpret_cast<Nb1Smem*>(nb1_raw);
  const int t = threadIdx.x, lane = t & 31;
  int n = count[0];
  if (n > N) n = N;
  // an odd chunk length keeps the chunked shared-memory walks (lane stride = C words) free of bank conflicts
  const int C = ((N + NB1_T - 1) / NB1_T) | 1;
  const int b0 = min(t * C, N), b1 = min(b0 + C, N);
  // ---- occupied cell box
  if (t < 3) S.box[t] = INT_MAX;
  else if (t < 6) S.box[t] = INT_MIN;
  __syncthreads();
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (int i = t; i < n; i += NB1_T) {  // coalesced walk
    const float4 p = pts[i];
    if (!finite3(p.x, p.y, p.z)) continue;
    const int c[3] = {floor_to_int_x86(fmul(p.x, inv_leaf)), floor_to_int_x86(fmul(p.y, inv_leaf)),
                      floor_to_int_x86(fmul(p.z, inv_leaf))};
    if (!grid_in_range(c[0], c[1], c[2])) {
      atomicExch(range_flag, 1);
      continue;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) lo[a] = min(lo[a], c[a]), hi[a] = max(hi[a], c[a]);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if (lane == 0) atomicMin(&S.box[a], lo[a]), atomicMax(&S.box[3 + a], hi[a]);
  }
  __syncthreads();
  const int mx = S.box[0], my = S.box[1], mz = S.box[2];
  const bool any = S.box[3] >= mx;
  const int bx = any ? bits_for(S.box[3] - mx) : 0, by = any ? bits_for(S.box[4] - my) : 0,
            bz = any ? bits_for(S.box[5] - mz) : 0;
  const int kbits = bx + by + bz;  // <= 48
  const unsigned long long kinv = 1ull << kbits;
  const bool in_smem = (N <= NB1_SMEM_MAX) && (kbits < 32);
  unsigned* sk = reinterpret_cast<unsigned*>(nb1_raw + sizeof(Nb1Smem));  // [2][NB1_SMEM_MAX] keys
  unsigned short* sv = reinterpret_cast<unsigned short*>(sk + 2 * NB1_SMEM_MAX);  // [2][NB1_SMEM_MAX] point indices
  // ---- keys.  Coalesced walk over the points; item i lands in slot i of the (chunked) sort storage.
  int nvalid_t = 0;
  for (int i = t; i < N; i += NB1_T) {
    unsigned long long k = kinv;
    if (i < n) {
      const float4 p = pts[i];
      if (finite3(p.x, p.y, p.z)) {
        const int ix = floor_to_int_x86(fmul(p.x, inv_leaf)), iy = floor_to_int_x86(fmul(p.y, inv_leaf)),
                  iz = floor_to_int_x86(fmul(p.z, inv_leaf));
        if (grid_in_range(ix, iy, iz)) {
          k = ((unsigned long long)(unsigned)(iz - mz) << (bx + by)) | ((unsigned long long)(unsigned)(iy - my) << bx) |
              (unsigned long long)(unsigned)(ix - mx);
          ++nvalid_t;
        }
      }
    }
    if (in_smem) {
      sk[i] = (unsigned)k;
      sv[i] = (unsigned short)i;
    } else {
      kA[i] = k;
      vA[i] = i;
    }
  }
  int n_valid;
  nb1_block_excl_scan(nvalid_t, S, &n_valid);
  if (in_smem)
    nb1_sort_emit<unsigned, unsigned short>(sk, sk + NB1_SMEM_MAX, sv, sv + NB1_SMEM_MAX, S, b0, b1, kbits, n_valid, bx, by, mx, my,
                                            mz, min_points, keys_out, vals_out, cell_start, leaf_flag, leaf_ord, n_cells_out,
                                            n_leaves_out);
  else
    nb1_sort_emit<unsigned long long, int>(kA, kB, vA, vB, S, b0, b1, kbits, n_valid, bx, by, mx, my, mz, min_points, keys_out,
                                           vals_out, cell_start, leaf_flag, leaf_ord, n_cells_out, n_leaves_out);
}

// ---- symmetric 3x3 / 6x6 cyclic Jacobi (same rotation formulas as the oracle's stand-in for Eigen)
template <int N>
__device__ void jacobi_eigh(const double* A_in, double* w, double* V) {
  double A[N * N];
  for (int i = 0; i < N * N; ++i) A[i] = A_in[i];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) V[i * N + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0;
    for (int p = 0; p < N; ++p)
      for (int q = p + 1; q < N; ++q) off = __dadd_rn(off, __dmul_rn(A[p * N + q], A[p * N + q]));
    if (off == 0.0) break;
    for (int p = 0; p < N; ++p)
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        if (apq == 0.0) continue;
        const double app = A[p * N + p], aqq = A[q * N + q];
        // negligible against both diagonal entries: zero it instead of rotating (same test as the oracle; without it
        // rounding noise keeps `off` above 0 and all 64 sweeps run -- 0.37 ms of one thread per voxel)
        const double g = __dmul_rn(100.0, fabs(apq));
        if (__dadd_rn(fabs(app), g) == fabs(app) && __dadd_rn(fabs(aqq), g) == fabs(aqq)) {
          A[p * N + q] = A[q * N + p] = 0.0;
          continue;
        }
        const double theta = __ddiv_rn(__dsub_rn(aqq, app), __dmul_rn(2.0, apq));
        const double t = __ddiv_rn(theta >= 0.0 ? 1.0 : -1.0,
                                   __dadd_rn(fabs(theta), __dsqrt_rn(__dadd_rn(__dmul_rn(theta, theta), 1.0))));
        const double c = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(__dmul_rn(t, t), 1.0)));
        const double s = __dmul_rn(t, c);
        for (int k = 0; k < N; ++k) {
          const double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = __dsub_rn(__dmul_rn(c, akp), __dmul_rn(s, akq));
          A[k * N + q] = __dadd_rn(__dmul_rn(s, akp), __dmul_rn(c, akq));
        }
        for (int k = 0; k < N; ++k) {
          const double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = __dsub_rn(__dmul_rn(c, apk), __dmul_rn(s, aqk));
          A[q * N + k] = __dadd_rn(__dmul_rn(s, apk), __dmul_rn(c, aqk));
        }
        for (int k = 0; k < N; ++k) {
          const double vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = __dsub_rn(__dmul_rn(c, vkp), __dmul_rn(s, vkq));
          V[k * N + q] = __dadd_rn(__dmul_rn(s, vkp), __dmul_rn(c, vkq));
        }
      }
  }
  for (int i = 0; i < N; ++i) w[i] = A[i * N + i];
  for (int i = 0; i < N - 1; ++i) {
    int m = i;
    for (int j = i + 1; j < N; ++j)
      if (w[j] < w[m]) m = j;
    if (m != i) {
      double tw = w[i];
      w[i] = w[m];
      w[m] = tw;
      for (int k = 0; k < N; ++k) {
        double tv = V[k * N + i];
        V[k * N + i] = V[k * N + m];
        V[k * N + m] = tv;
      }
    }
  }
}

__device__ void inv3(const double* M, double* I) {
  const double c00 = __dsub_rn(__dmul_rn(M[4], M[8]), __dmul_rn(M[5], M[7]));
  const double c01 = __dsub_rn(__dmul_rn(M[5], M[6]), __dmul_rn(M[3], M[8]));
  const double c02 = __dsub_rn(__dmul_rn(M[3], M[7]), __dmul_rn(M[4], M[6]));
  const double det = __dadd_rn(__dadd_rn(__dmul_rn(M[0], c00), __dmul_rn(M[1], c01)), __dmul_rn(M[2], c02));
  const double id = __ddiv_rn(1.0, det);
  I[0] = __dmul_rn(c00, id);
  I[1] = __dmul_rn(__dsub_rn(__dmul_rn(M[2], M[7]), __dmul_rn(M[1], M[8])), id);
  I[2] = __dmul_rn(__dsub_rn(__dmul_rn(M[1], M[5]), __dmul_rn(M[2], M[4])), id);
  I[3] = __dmul_rn(c01, id);
  I[4] = __dmul_rn(__dsub_rn(__dmul_rn(M[0], M[8]), __dmul_rn(M[2], M[6])), id);
  I[5] = __dmul_rn(__dsub_rn(__dmul_rn(M[2], M[3]), __dmul_rn(M[0], M[5])), id);
  I[6] = __dmul_rn(c02, id);
  I[7] = __dmul_rn(__dsub_rn(__dmul_rn(M[1], M[6]), __dmul_rn(M[0], M[7])), id);
  I[8] = __dmul_rn(__dsub_rn(__dmul_rn(M[0], M[4]), __dmul_rn(M[1], M[3])), id);
}

// One thread per accepted cell: sequential fp64 sums in input order, then the VoxelGridCovariance second pass.
// One WARP per cell.  The per-voxel sums are sequential in point order (VoxelGridCovariance accumulates mean_ and cov_
// point by point in double; the oracle does the same), so the order of every accumulator's additions is fixed -- but the
// 3 + 9 double accumulators and the 3 float centroid sums are independent chains.  The warp loads 32 points at a time
// (coalesced indices, gathered points), every lane forms the 12 addends of ITS point (the products are order-free) and
// parks them in shared memory; lanes 0..11 then run one dependent DADD chain each over the 32 rows, lanes 12..14 the float
// chains.  A cell of 2000 points costs 2000 chained DADDs, not 2000 chained global-memory round trips.
constexpr int VS_WARPS = 8, VS_PAD = 17;
__global__ void __launch_bounds__(VS_WARPS * 32)
k_voxel_stats(const float4* __restrict__ pts, int stride, int pstride, const unsigned long long* __restrict__ keys,
              const int* __restrict__ vals, const int* __restrict__ cell_start, const int* __restrict__ leaf_flag,
              const int* __restrict__ leaf_ord, const int* __restrict__ n_cells, double eig_mult,
              VoxRec* __restrict__ vox, unsigned long long* __restrict__ hkeys, int* __restrict__ hvals,
              unsigned cap_mask, int* __restrict__ nvox_seg) {
  __shared__ double s_add[VS_WARPS][32][VS_PAD];  // [point][addend], padded: conflict-free 64-bit rows and columns
  __shared__ float s_xyz[VS_WARPS][32][3];
  const int nc = *n_cells;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int c = blockIdx.x * VS_WARPS + wib; c < nc; c += gridDim.x * VS_WARPS) {
    if (!leaf_flag[c]) continue;
    const int b = cell_start[c], e = cell_start[c + 1], n = e - b;
    const unsigned long long key = keys[b];
    const int seg = (int)(key >> 48);
    double acc = 0.0;
    float csf = 0.f;
    for (int base = b; base < e; base += 32) {
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);  // rows past the end add +0 (exact)
      if (base + lane < e) p = pts[(size_t)seg * stride + vals[base + lane]];
      const double v[3] = {(double)p.x, (double)p.y, (double)p.z};
      __syncwarp();
      double* row = s_add[wib][lane];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        row[a] = v[a];
        s_xyz[wib][lane][a] = a == 0 ? p.x : (a == 1 ? p.y : p.z);
#pragma unroll
        for (int q = 0; q < 3; ++q) row[3 + a * 3 + q] = __dmul_rn(v[a], v[q]);
      }
      __syncwarp();
      if (lane < 12) {
#pragma unroll 8
        for (int q = 0; q < 32; ++q) acc = __dadd_rn(acc, s_add[wib][q][lane]);
      } else if (lane < 15) {
#pragma unroll 8
        for (int q = 0; q < 32; ++q) csf = fadd(csf, s_xyz[wib][q][lane - 12]);
      }
    }
    double sum[3], sxx[9];
    float cs[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      sum[a] = __shfl_sync(0xffffffffu, acc, a);
      cs[a] = __shfl_sync(0xffffffffu, csf, 12 + a);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) sxx[k] = __shfl_sync(0xffffffffu, acc, 3 + k);
    if (lane != 0) continue;
    VoxRec R;
    R.ijk[0] = (int)(key & 0xFFFF) - 32768;
    R.ijk[1] = (int)((key >> 16) & 0xFFFF) - 32768;
    R.ijk[2] = (int)((key >> 32) & 0xFFFF) - 32768;
    R.npts = n;
    R.pad = 0;
    const double dn = (double)n;
    for (int a = 0; a < 3; ++a) {
      R.centroid[a] = __fdiv_rn(cs[a], (float)n);
      R.mean[a] = __ddiv_rn(sum[a], dn);
    }
    double cov[9];
    for (int a = 0; a < 3; ++a)
      for (int q = 0; q < 3; ++q)
        cov[a * 3 + q] = __dadd_rn(__ddiv_rn(__dsub_rn(sxx[a * 3 + q], __dmul_rn(2.0, __dmul_rn(sum[a], R.mean[q]))), dn),
                                   __dmul_rn(R.mean[a], R.mean[q]));
    const double nf = __ddiv_rn(__dsub_rn(dn, 1.0), dn);
    for (int k = 0; k < 9; ++k) cov[k] = __dmul_rn(cov[k], nf);
    double sym[9];
    for (int a = 0; a < 3; ++a)
      for (int q = 0; q < 3; ++q) sym[a * 3 + q] = (a >= q) ? cov[a * 3 + q] : cov[q * 3 + a];
    double w[3], E[9];
    jacobi_eigh<3>(sym, w, E);
    for (int a = 0; a < 3; ++a) R.evals[a] = w[a];
    for (int k = 0; k < 9; ++k) R.icov[k] = 0.0;
    if (w[0] < 0 || w[1] < 0 || w[2] <= 0) {
      R.npts = -1;
    } else {
      const double min_ev = __dmul_rn(eig_mult, w[2]);
      if (w[0] < min_ev) {
        w[0] = min_ev;
        if (w[1] < min_ev) w[1] = min_ev;
        for (int a = 0; a < 3; ++a)
          for (int q = 0; q < 3; ++q) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s = __dadd_rn(s, __dmul_rn(__dmul_rn(E[a * 3 + k], w[k]), E[q * 3 + k]));
            cov[a * 3 + q] = s;
          }
      }
      inv3(cov, R.icov);
      bool bad = false;
      for (int k = 0; k < 9; ++k)
        if (!isfinite(R.icov[k])) bad = true;
      if (bad) R.npts = -1;
    }
    for (int k = 0; k < 9; ++k) R.cov[k] = cov[k];
    const int li = leaf_ord[c];
    vox[li] = R;
    // publish in the lookup table (keys are unique: one thread per cell)
    unsigned s = grid_hash(key) & cap_mask;
    while (atomicCAS(&hkeys[s], KEY_INVALID, key) != KEY_INVALID) s = (s + 1) & cap_mask;
    hvals[s] = li;
    atomicAdd(&nvox_seg[seg], 1);
  }
}

struct NdtGridDev {
  VoxRec* vox = nullptr;
  unsigned long long* hkeys = nullptr;
  int* hvals = nullptr;
  unsigned cap_mask = 0;
  float inv_leaf = 1.f;
  float r2 = 1.f;
  int shared_target = 0;
  int* nvox_seg = nullptr;
  int* n_leaves = nullptr;
  long long vox_cap = 0;
};

__device__ __forceinline__ int ndt_lookup(const NdtGridDev& G, unsigned long long key) {
  unsigned s = grid_hash(key) & G.cap_mask;
  while (true) {
    const unsigned long long k = __ldg(&G.hkeys[s]);
    if (k == key) return __ldg(&G.hvals[s]);
    if (k == KEY_INVALID) return -1;
    s = (s + 1) & G.cap_mask;
  }
}

// ------------------------------------------------------------------------------------------------ evaluation
struct NdtEval {       // per pair, written by the controller, read by k_ndt_eval
  float T[16];         // transform applied to the ORIGINAL source for this evaluation
  double ja[8][3];     // j_ang_a .. j_ang_h
  double ha[15][3];    // h_ang_a2,a3,b2,b3,c2,c3,d1,d2,d3,e1,e2,e3,f1,f2,f3
  int want_hessian;
  int hessian_only;
  int active;          // 0: pair finished, skip
  int pad;
};

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// x_trans . (c_inv * h)
__device__ __forceinline__ double xCh(const double* xr, const double* C, const double* h) {
  return xr[0] * (C[0] * h[0] + C[1] * h[1] + C[2] * h[2]) + xr[1] * (C[3] * h[0] + C[4] * h[1] + C[5] * h[2]) +
         xr[2] * (C[6] * h[0] + C[7] * h[1] + C[8] * h[2]);
}

// One source point against its 3x3x3 voxel neighbourhood: computeDerivatives' loop body (transform, radiusSearch over the
// voxel centroids, updateDerivatives / updateHessian).  Shared by the per-evaluation kernel and the persistent kernel.
// n_pairs counts the (point, voxel) pairs that contributed (roofline accounting).
// LISTED: the (point, voxel) pair was already found by a candidate pass (k_ndt_persist), `listed_vi` is its voxel.
template <bool LISTED>
__device__ __forceinline__ void ndt_point_eval(const float4 p, const NdtEval& E, const NdtGridDev& G, int tseg, bool wantH,
                                               double d1, double d2, double* acc, int& n_pairs, int listed_vi = -1) {
  do {
    if (!finite3(p.x, p.y, p.z)) break;
    const float3 xt = xform_point(E.T, p.x, p.y, p.z);
    if (!finite3(xt.x, xt.y, xt.z)) break;
    const int cx = floor_to_int_x86(fmul(xt.x, G.inv_leaf)), cy = floor_to_int_x86(fmul(xt.y, G.inv_leaf)),
              cz = floor_to_int_x86(fmul(xt.z, G.inv_leaf));
    if (!grid_in_range(cx, cy, cz)) break;
    const double x[3] = {(double)p.x, (double)p.y, (double)p.z};
    // computePointDerivatives: J = [I | angular columns]
    const double J3[3] = {0.0, dot3(x, E.ja[0]), dot3(x, E.ja[1])};
    const double J4[3] = {dot3(x, E.ja[2]), dot3(x, E.ja[3]), dot3(x, E.ja[4])};
    const double J5[3] = {dot3(x, E.ja[5]), dot3(x, E.ja[6]), dot3(x, E.ja[7])};
    double ha[3] = {0, 0, 0}, hb[3] = {0, 0, 0}, hc[3] = {0, 0, 0}, hd[3] = {0, 0, 0}, he[3] = {0, 0, 0}, hf[3] = {0, 0, 0};
    if (wantH) {
      ha[1] = dot3(x, E.ha[0]); ha[2] = dot3(x, E.ha[1]);
      hb[1] = dot3(x, E.ha[2]); hb[2] = dot3(x, E.ha[3]);
      hc[1] = dot3(x, E.ha[4]); hc[2] = dot3(x, E.ha[5]);
      hd[0] = dot3(x, E.ha[6]); hd[1] = dot3(x, E.ha[7]); hd[2] = dot3(x, E.ha[8]);
      he[0] = dot3(x, E.ha[9]); he[1] = dot3(x, E.ha[10]); he[2] = dot3(x, E.ha[11]);
      hf[0] = dot3(x, E.ha[12]); hf[1] = dot3(x, E.ha[13]); hf[2] = dot3(x, E.ha[14]);
    }
    for (int dz = -1; dz <= (LISTED ? -1 : 1); ++dz)
      for (int dy = -1; dy <= (LISTED ? -1 : 1); ++dy)
        for (int dx = -1; dx <= (LISTED ? -1 : 1); ++dx) {
          const int vi = LISTED ? listed_vi : ndt_lookup(G, ndt_key(tseg, cx + dx, cy + dy, cz + dz));
          if (vi < 0) continue;
          const VoxRec* V = &G.vox[vi];
          if (!LISTED) {
            const float ddx = __fsub_rn(xt.x, V->centroid[0]), ddy = __fsub_rn(xt.y, V->centroid[1]),
                        ddz = __fsub_rn(xt.z, V->centroid[2]);
            const float dist = fadd(fadd(fmul(ddx, ddx), fmul(ddy, ddy)), fmul(ddz, ddz));
            if (!(dist < G.r2)) continue;  // FLANN radius search: strict
          }
          const double xr[3] = {(double)xt.x - V->mean[0], (double)xt.y - V->mean[1], (double)xt.z - V->mean[2]};
          double C[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) C[k] = V->icov[k];
          const double q[3] = {C[0] * xr[0] + C[1] * xr[1] + C[2] * xr[2], C[3] * xr[0] + C[4] * xr[1] + C[5] * xr[2],
                               C[6] * xr[0] + C[7] * xr[1] + C[8] * xr[2]};
          double e = exp(-d2 * dot3(xr, q) / 2);
          const double score_inc = -d1 * e;
          e = d2 * e;
          if (e > 1 || e < 0 || e != e) continue;  // updateDerivatives returns 0: no score either
          e *= d1;
          ++n_pairs;
          // c_inv * J.col(i) and x_trans . (c_inv * J.col(i))
          double cJ[6][3];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            cJ[0][r] = C[r * 3 + 0];
            cJ[1][r] = C[r * 3 + 1];
            cJ[2][r] = C[r * 3 + 2];
            cJ[3][r] = C[r * 3 + 1] * J3[1] + C[r * 3 + 2] * J3[2];
            cJ[4][r] = C[r * 3 + 0] * J4[0] + C[r * 3 + 1] * J4[1] + C[r * 3 + 2] * J4[2];
            cJ[5][r] = C[r * 3 + 0] * J5[0] + C[r * 3 + 1] * J5[1] + C[r * 3 + 2] * J5[2];
          }
          double xcJ[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) xcJ[k] = dot3(xr, cJ[k]);
          if (!E.hessian_only) {
            acc[0] += score_inc;
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[1 + k] += xcJ[k] * e;
          }
          if (wantH) {
            const double* Jc[6] = {nullptr, nullptr, nullptr, J3, J4, J5};
            // x_trans . (c_inv * point_hessian block): only the angular 3x3 corner is non-zero
            const double qa = xCh(xr, C, ha), qb = xCh(xr, C, hb), qc = xCh(xr, C, hc), qd = xCh(xr, C, hd),
                         qe = xCh(xr, C, he), qf = xCh(xr, C, hf);
            int h = 7;
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
              for (int j = i; j < 6; ++j) {
                // J.col(j) . (c_inv * J.col(i))
                double jcj;
                if (j < 3)
                  jcj = cJ[i][j];
                else
                  jcj = dot3(Jc[j], cJ[i]);
                double ph = 0.0;
                if (i == 3 && j == 3) ph = qa;
                if (i == 3 && j == 4) ph = qb;
                if (i == 3 && j == 5) ph = qc;
                if (i == 4 && j == 4) ph = qd;
                if (i == 4 && j == 5) ph = qe;
                if (i == 5 && j == 5) ph = qf;
                acc[h++] += e * (-d2 * xcJ[i] * xcJ[j] + ph + jcj);
              }
          }
        }
    } while (false);
}

__global__ void __launch_bounds__(NT) k_ndt_eval(const float4* __restrict__ src, const int* __restrict__ count, int stride,
                                                 const NdtEval* __restrict__ evals, NdtGridDev G, double d1, double d2,
                                                 double* __restrict__ partials) {
  __shared__ NdtEval E;
  __shared__ double s_red[NT / 32][NACC];
  const int seg = blockIdx.y;
  if (!evals[seg].active) return;
  for (int k = threadIdx.x; k < (int)(sizeof(NdtEval) / 4); k += NT) ((int*)&E)[k] = ((const int*)&evals[seg])[k];
  __syncthreads();
  const int n = count[seg];
  const int tseg = G.shared_target ? 0 : seg;
  const bool wantH = E.want_hessian != 0;
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;

  int n_pairs_unused = 0;
  for (int i = blockIdx.x * NT + threadIdx.x; i < n; i += gridDim.x * NT)
    ndt_point_eval<false>(src[(size_t)seg * stride + i], E, G, tseg, wantH, d1, d2, acc, n_pairs_unused);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) s_red[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    double v = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) v += s_red[w][threadIdx.x];
    partials[((size_t)seg * gridDim.x + blockIdx.x) * NACC + threadIdx.x] = v;
  }
}

// ------------------------------------------------------------------------------------------------ controller
enum { PH_INIT = 0, PH_MT_FIRST = 1, PH_MT_LOOP = 2, PH_MT_HESS = 3 };

struct NdtState {
  double p[6], x_t[6], dir[6], g[6], H[36];
  double score;
  double phi_0, d_phi_0, a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t, psi_t, d_psi_t;
  float final_T[16];
  int phase, nr_iterations, converged, done, n_deriv, n_hess, step_iterations, open_interval, interval_converged, pad;
};

struct NdtCtl {
  int max_iterations;
  double eps, step_size;
};

__device__ void pose_to_matrix(const double* p, float* T) {
  const float rx = (float)p[3], ry = (float)p[4], rz = (float)p[5];
  const float cx = cosf(rx), sx = sinf(rx), cy = cosf(ry), sy = sinf(ry), cz = cosf(rz), sz = sinf(rz);
  const float Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
  const float Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
  const float Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  float A[9], R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k) s = fadd(s, fmul(Rx[r * 3 + k], Ry[k * 3 + c]));
      A[r * 3 + c] = s;
    }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k) s = fadd(s, fmul(A[r * 3 + k], Rz[k * 3 + c]));
      R[r * 3 + c] = s;
    }
  for (int i = 0; i < 16; ++i) T[i] = 0.f;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T[c * 4 + r] = R[r * 3 + c];
  T[12] = (float)p[0];
  T[13] = (float)p[1];
  T[14] = (float)p[2];
  T[15] = 1.f;
}

// Eigen 3.3 eulerAngles(0,1,2) on the rotation block (float)
__device__ void matrix_to_pose(const float* T, double* p) {
#define MM(r, c) T[(c)*4 + (r)]
  float res0 = atan2f(MM(1, 2), MM(2, 2)), res1, res2;
  const float c2 = sqrtf(fadd(fmul(MM(0, 0), MM(0, 0)), fmul(MM(0, 1), MM(0, 1))));
  if (res0 > 0.f) {
    res0 -= 3.14159265358979323846f;
    res1 = atan2f(-MM(0, 2), -c2);
  } else {
    res1 = atan2f(-MM(0, 2), c2);
  }
  const float s1 = sinf(res0), c1 = cosf(res0);
  res2 = atan2f(fadd(fmul(s1, MM(2, 0)), -fmul(c1, MM(1, 0))), fadd(fmul(c1, MM(1, 1)), -fmul(s1, MM(2, 1))));
#undef MM
  p[0] = T[12];
  p[1] = T[13];
  p[2] = T[14];
  p[3] = -res0;
  p[4] = -res1;
  p[5] = -res2;
}

__device__ void angle_derivatives(const double* p, NdtEval* E) {
  double cx, cy, cz, sx, sy, sz;
  if (fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = cos(p[3]); sx = sin(p[3]); }
  if (fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = cos(p[4]); sy = sin(p[4]); }
  if (fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = cos(p[5]); sz = sin(p[5]); }
#define SET3(v, a, b, c) { (v)[0] = (a); (v)[1] = (b); (v)[2] = (c); }
  SET3(E->ja[0], (-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy));
  SET3(E->ja[1], (cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy));
  SET3(E->ja[2], (-sy * cz), sy * sz, cy);
  SET3(E->ja[3], sx * cy * cz, (-sx * cy * sz), sx * sy);
  SET3(E->ja[4], (-cx * cy * cz), cx * cy * sz, (-cx * sy));
  SET3(E->ja[5], (-cy * sz), (-cy * cz), 0);
  SET3(E->ja[6], (cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0);
  SET3(E->ja[7], (sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0);
  SET3(E->ha[0], (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy);
  SET3(E->ha[1], (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy));
  SET3(E->ha[2], (cx * cy * cz), (-cx * cy * sz), (cx * sy));
  SET3(E->ha[3], (sx * cy * cz), (-sx * cy * sz), (sx * sy));
  SET3(E->ha[4], (-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0);
  SET3(E->ha[5], (cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0);
  SET3(E->ha[6], (-cy * cz), (cy * sz), (sy));  // PCL literal (+sy)
  SET3(E->ha[7], (-sx * sy * cz), (sx * sy * sz), (sx * cy));
  SET3(E->ha[8], (cx * sy * cz), (-cx * sy * sz), (-cx * cy));
  SET3(E->ha[9], (sy * sz), (sy * cz), 0);
  SET3(E->ha[10], (-sx * cy * sz), (-sx * cy * cz), 0);
  SET3(E->ha[11], (cx * cy * sz), (cx * cy * cz), 0);
  SET3(E->ha[12], (-cy * cz), (cy * sz), 0);
  SET3(E->ha[13], (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0);
  SET3(E->ha[14], (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0);
#undef SET3
}

// Newton step H dp = -g.  Near the optimum the NDT Hessian is symmetric and definite (negative definite for PCL's
// score, which is maximised): an unrolled LDL^T of +-H in registers (6 divisions) gives the same dp as the
// pseudo-inverse to rounding.  Anything else (a pivot that is not clearly of the common sign) takes the Jacobi
// pseudo-inverse that restates PCL's JacobiSVD solve, rank threshold included.
__device__ bool solve6_ldlt(const double* A_in, const double* g_in, double* x) {
  double L[36], D[6], A[36], g[6];
  const double sgn = A_in[0] < 0.0 ? -1.0 : 1.0;  // solve (sgn H) dp = -(sgn g)
#pragma unroll
  for (int i = 0; i < 36; ++i) A[i] = sgn * A_in[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) g[i] = sgn * g_in[i];
  double maxd = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) maxd = fmax(maxd, fabs(A[i * 6 + i]));
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = A[j * 6 + j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
    if (!(d > 1e-9 * maxd)) return false;
    D[j] = d;
    const double inv = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double v = A[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= L[i * 6 + k] * L[j * 6 + k] * D[k];
      L[i * 6 + j] = v * inv;
    }
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double v = -g[i];
#pragma unroll
    for (int k = 0; k < i; ++k) v -= L[i * 6 + k] * y[k];
    y[i] = v;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) y[i] /= D[i];
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double v = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) v -= L[k * 6 + i] * x[k];
    x[i] = v;
  }
  return true;
}

// Indefinite but well-conditioned Hessians (far from the optimum, fine voxels): Gaussian elimination with partial
// pivoting, rows swapped by predicated moves so that every index stays static (registers).  Refuses (-> pseudo-inverse)
// when a pivot falls below 1e-7 of the largest entry.
__device__ bool solve6_gepp(const double* A_in, const double* g, double* x) {
  double A[36], b[6];
  double amax = 0;
#pragma unroll
  for (int i = 0; i < 36; ++i) {
    A[i] = A_in[i];
    amax = fmax(amax, fabs(A[i]));
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) b[i] = -g[i];
  if (!(amax > 0.0)) return false;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    int piv = j;
    double pv = fabs(A[j * 6 + j]);
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      const double v = fabs(A[i * 6 + j]);
      if (v > pv) {
        pv = v;
        piv = i;
      }
    }
    if (!(pv > 1e-7 * amax)) return false;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      if (i == piv) {
#pragma unroll
        for (int k = j; k < 6; ++k) {
          const double t = A[j * 6 + k];
          A[j * 6 + k] = A[i * 6 + k];
          A[i * 6 + k] = t;
        }
        const double t = b[j];
        b[j] = b[i];
        b[i] = t;
      }
    }
    const double inv = 1.0 / A[j * 6 + j];
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      const double f = A[i * 6 + j] * inv;
#pragma unroll
      for (int k = j + 1; k < 6; ++k) A[i * 6 + k] -= f * A[j * 6 + k];
      b[i] -= f * b[j];
    }
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double v = b[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) v -= A[i * 6 + k] * x[k];
    x[i] = v / A[i * 6 + i];
  }
  return true;
}

__device__ void solve_newton(const double* H, const double* g, double* dp) {
  double S[36], w[6], V[36];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) S[i * 6 + j] = 0.5 * (H[i * 6 + j] + H[j * 6 + i]);
  if (solve6_ldlt(S, g, dp)) return;
  if (solve6_gepp(S, g, dp)) return;
  jacobi_eigh<6>(S, w, V);
  double wmax = 0;
  for (int i = 0; i < 6; ++i) wmax = fmax(wmax, fabs(w[i]));
  const double thr = 6 * DBL_EPSILON * wmax;
  for (int i = 0; i < 6; ++i) dp[i] = 0;
  for (int k = 0; k < 6; ++k) {
    if (!(fabs(w[k]) > thr)) continue;
    double c = 0;
    for (int i = 0; i < 6; ++i) c += V[i * 6 + k] * (-g[i]);
    c /= w[k];
    for (int i = 0; i < 6; ++i) dp[i] += V[i * 6 + k] * c;
  }
}

__device__ __forceinline__ double psiMT(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
__device__ __forceinline__ double dpsiMT(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

__device__ bool updateIntervalMT(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t,
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

__device__ double trialValueSelectionMT(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t,
                                        double f_t, double g_t) {
  if (f_t > f_l) {
    const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    const double w = sqrt(z * z - g_t * g_l);
    const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    const double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    if (fabs(a_c - a_l) < fabs(a_q - a_l)) return a_c;
    return 0.5 * (a_q + a_c);
  } else if (g_t * g_l < 0) {
    const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    const double w = sqrt(z * z - g_t * g_l);
    const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    if (fabs(a_c - a_t) >= fabs(a_s - a_t)) return a_c;
    return a_s;
  } else if (fabs(g_t) <= fabs(g_l)) {
    const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    const double w = sqrt(z * z - g_t * g_l);
    const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    const double a_t_next = (fabs(a_c - a_t) < fabs(a_s - a_t)) ? a_c : a_s;
    if (a_t > a_l) return fmin(a_t + 0.66 * (a_u - a_t), a_t_next);
    return fmax(a_t + 0.66 * (a_u - a_t), a_t_next);
  } else {
    const double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
    const double w = sqrt(z * z - g_t * g_u);
    return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
  }
}

__global__ void k_ndt_init(NdtState* __restrict__ st, NdtEval* __restrict__ ev, const float* __restrict__ guess, int n_seg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  NdtState* S = &st[s];
  NdtEval* E = &ev[s];
  for (int i = 0; i < 16; ++i) S->final_T[i] = guess ? guess[s * 16 + i] : ((i % 5 == 0) ? 1.f : 0.f);
  matrix_to_pose(S->final_T, S->p);
  for (int i = 0; i < 6; ++i) S->x_t[i] = S->p[i];
  // first computeDerivatives: the cloud moved by the GUESS MATRIX (not by the matrix rebuilt from p)
  for (int i = 0; i < 16; ++i) E->T[i] = S->final_T[i];
  angle_derivatives(S->p, E);
  E->want_hessian = 1;
  E->hessian_only = 0;
  E->active = 1;
  E->pad = 0;
  S->phase = PH_INIT;
  S->nr_iterations = 0;
  S->converged = 0;
  S->done = 0;
  S->n_deriv = 0;
  S->n_hess = 0;
  S->score = 0;
  S->step_iterations = 0;
}

// Newton / More-Thuente controller for one pair (one thread); returns true when the pair has finished.
__device__ bool ndt_control_step(NdtState* S, NdtEval* E, const double* sums, const NdtCtl& ctl) {
  const double mu = 1.e-4, nu = 0.9;
  const double step_max = ctl.step_size, step_min = ctl.eps / 2;
  // ---- consume the evaluation that just ran
  if (E->want_hessian) {
    int h = 7;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) {
        S->H[i * 6 + j] = sums[h];
        S->H[j * 6 + i] = sums[h];
        ++h;
      }
    S->n_hess++;
  }
  if (!E->hessian_only) {
    S->score = sums[0];
    for (int i = 0; i < 6; ++i) S->g[i] = sums[1 + i];
    S->n_deriv++;
  }
  bool goto_newton = false, goto_after = false, goto_check = false;
  switch (S->phase) {
    case PH_INIT:
      goto_newton = true;
      break;
    case PH_MT_FIRST: {
      S->phi_t = -S->score;
      double d = 0;
      for (int i = 0; i < 6; ++i) d -= S->g[i] * S->dir[i];
      S->d_phi_t = d;
      S->psi_t = psiMT(S->a_t, S->phi_t, S->phi_0, S->d_phi_0, mu);
      S->d_psi_t = dpsiMT(S->d_phi_t, S->d_phi_0, mu);
      goto_check = true;
      break;
    }
    case PH_MT_LOOP: {
      S->phi_t = -S->score;
      double d = 0;
      for (int i = 0; i < 6; ++i) d -= S->g[i] * S->dir[i];
      S->d_phi_t = d;
      S->psi_t = psiMT(S->a_t, S->phi_t, S->phi_0, S->d_phi_0, mu);
      S->d_psi_t = dpsiMT(S->d_phi_t, S->d_phi_0, mu);
      if (S->open_interval && (S->psi_t <= 0 && S->d_psi_t >= 0)) {
        S->open_interval = 0;
        S->f_l = S->f_l + S->phi_0 - mu * S->d_phi_0 * S->a_l;
        S->g_l = S->g_l + mu * S->d_phi_0;
        S->f_u = S->f_u + S->phi_0 - mu * S->d_phi_0 * S->a_u;
        S->g_u = S->g_u + mu * S->d_phi_0;
      }
      if (S->open_interval)
        S->interval_converged = updateIntervalMT(S->a_l, S->f_l, S->g_l, S->a_u, S->f_u, S->g_u, S->a_t, S->psi_t, S->d_psi_t);
      else
        S->interval_converged = updateIntervalMT(S->a_l, S->f_l, S->g_l, S->a_u, S->f_u, S->g_u, S->a_t, S->phi_t, S->d_phi_t);
      S->step_iterations++;
      goto_check = true;
      break;
    }
    case PH_MT_HESS:
      goto_after = true;
      break;
  }

  for (int guard = 0; guard < 8; ++guard) {
    if (goto_check) {
      goto_check = false;
      if (!S->interval_converged && S->step_iterations < 10 && !(S->psi_t <= 0 && S->d_phi_t <= -nu * S->d_phi_0)) {
        if (S->open_interval)
          S->a_t = trialValueSelectionMT(S->a_l, S->f_l, S->g_l, S->a_u, S->f_u, S->g_u, S->a_t, S->psi_t, S->d_psi_t);
        else
          S->a_t = trialValueSelectionMT(S->a_l, S->f_l, S->g_l, S->a_u, S->f_u, S->g_u, S->a_t, S->phi_t, S->d_phi_t);
        S->a_t = fmin(S->a_t, step_max);
        S->a_t = fmax(S->a_t, step_min);
        for (int i = 0; i < 6; ++i) S->x_t[i] = S->p[i] + S->dir[i] * S->a_t;
        pose_to_matrix(S->x_t, S->final_T);
        for (int i = 0; i < 16; ++i) E->T[i] = S->final_T[i];
        angle_derivatives(S->x_t, E);
        E->want_hessian = 0;
        E->hessian_only = 0;
        S->phase = PH_MT_LOOP;
        return false;
      }
      if (S->step_iterations) {  // computeHessian at the accepted step (angle derivatives are already those of x_t)
        E->want_hessian = 1;
        E->hessian_only = 1;
        S->phase = PH_MT_HESS;
        return false;
      }
      goto_after = true;
    }
    if (goto_after) {
      goto_after = false;
      for (int i = 0; i < 6; ++i) S->p[i] += S->dir[i] * S->a_t;  // delta_p = normalised direction * step length
      if (S->nr_iterations > ctl.max_iterations || (S->nr_iterations && (fabs(S->a_t) < ctl.eps))) S->converged = 1;
      S->nr_iterations++;
      if (S->converged) break;
      goto_newton = true;
    }
    if (goto_newton) {
      goto_newton = false;
      double dp[6];
      solve_newton(S->H, S->g, dp);
      double nrm = 0;
      for (int i = 0; i < 6; ++i) nrm += dp[i] * dp[i];
      nrm = sqrt(nrm);
      if (nrm == 0 || nrm != nrm) {
        S->converged = (nrm == nrm) ? 1 : 0;
        break;
      }
      for (int i = 0; i < 6; ++i) S->dir[i] = dp[i] / nrm;
      // computeStepLengthMT prologue
      S->phi_0 = -S->score;
      double d0 = 0;
      for (int i = 0; i < 6; ++i) d0 -= S->g[i] * S->dir[i];
      if (d0 >= 0) {
        if (d0 == 0) {  // "return 0": no step, no new evaluation
          S->a_t = 0;
          goto_after = true;
          continue;
        }
        d0 *= -1;
        for (int i = 0; i < 6; ++i) S->dir[i] *= -1;
      }
      S->d_phi_0 = d0;
      S->step_iterations = 0;
      S->a_l = 0;
      S->a_u = 0;
      S->f_l = psiMT(0, S->phi_0, S->phi_0, d0, mu);
      S->g_l = dpsiMT(d0, d0, mu);
      S->f_u = S->f_l;
      S->g_u = S->g_l;
      S->interval_converged = (step_max - step_min) < 0 ? 1 : 0;
      S->open_interval = 1;
      double a_t = nrm;
      a_t = fmin(a_t, step_max);
      a_t = fmax(a_t, step_min);
      S->a_t = a_t;
      for (int i = 0; i < 6; ++i) S->x_t[i] = S->p[i] + S->dir[i] * a_t;
      pose_to_matrix(S->x_t, S->final_T);
      for (int i = 0; i < 16; ++i) E->T[i] = S->final_T[i];
      angle_derivatives(S->x_t, E);
      E->want_hessian = 1;
      E->hessian_only = 0;
      S->phase = PH_MT_FIRST;
      return false;
    }
  }
  // finished (converged or NaN step)
  S->done = 1;
  E->active = 0;
  return true;
}

// One warp per pair: the lanes combine the CTA partials (fixed block order) and stage the pair's state in shared
// memory, lane 0 runs the controller on the shared copy (no dependent global round trips), the lanes write it back.
__global__ void __launch_bounds__(32) k_ndt_control(NdtState* __restrict__ st, NdtEval* __restrict__ ev,
                                                    const double* __restrict__ partials, int nblk, NdtCtl ctl,
                                                    int* __restrict__ n_active) {
  const int seg = blockIdx.x;
  if (st[seg].done) return;
  const int lane = threadIdx.x;
  __shared__ double sums[NACC];
  __shared__ NdtState sS;
  static_assert(sizeof(NdtState) % 4 == 0, "NdtState is copied as 32-bit words");
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(&st[seg]);
    unsigned* dst = reinterpret_cast<unsigned*>(&sS);
    for (int i = lane; i < (int)(sizeof(NdtState) / 4); i += 32) dst[i] = src[i];
  }
  if (lane < NACC) {
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0;  // four independent chains; combined in a fixed order
    int b = 0;
    for (; b + 3 < nblk; b += 4) {
      v0 += partials[((size_t)seg * nblk + b) * NACC + lane];
      v1 += partials[((size_t)seg * nblk + b + 1) * NACC + lane];
      v2 += partials[((size_t)seg * nblk + b + 2) * NACC + lane];
      v3 += partials[((size_t)seg * nblk + b + 3) * NACC + lane];
    }
    for (; b < nblk; ++b) v0 += partials[((size_t)seg * nblk + b) * NACC + lane];
    sums[lane] = (v0 + v1) + (v2 + v3);
  }
  __syncwarp();
  bool finished = false;
  if (lane == 0) finished = ndt_control_step(&sS, &ev[seg], sums, ctl);
  __syncwarp();
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(&sS);
    unsigned* dst = reinterpret_cast<unsigned*>(&st[seg]);
    for (int i = lane; i < (int)(sizeof(NdtState) / 4); i += 32) dst[i] = src[i];
  }
  if (finished) atomicSub(n_active, 1);
}


// ------------------------------------------------------------------------------------------------ persistent align
// The WHOLE align of a few pairs -- every derivative evaluation of the Newton iterations and of the More-Thuente line
// search, the 28-sum reduction, the Newton solve and the line-search state machine -- inside ONE cooperative launch that
// spans the chip: no launch per evaluation, no host polling (the per-evaluation path costs ~35 us of launch + control per
// evaluation for ~10 us of arithmetic on a 6 k-point pair, and polls the host every few evaluations).
//   * the CTAs of the grid are divided evenly among the pairs; a pair's CTAs share its source points round-robin;
//   * one evaluation = candidate pass (27 hash probes per point; hits compacted IN POINT ORDER into a shared-memory pair
//     list) -> derivative bodies spread evenly over the threads (fixed pair -> thread map: reproducible fp64 sums) -> CTA
//     partials to global memory -> ONE grid barrier -> every CTA of the pair sums the partials in CTA order and runs the
//     identical controller on its own copy of the state, so nothing is broadcast.  The partial buffers alternate by
//     evaluation parity (a CTA can be at most one evaluation ahead of another).
// FP64-bound work with a long dependent chain per (point, voxel) pair: the latency of one evaluation is one body, which is
// why the pairs get as many CTAs as the chip has rather than one cluster each.
constexpr int NPT = 256;            // threads per CTA
constexpr int NP_CAP = 27 * NPT;    // (point, voxel) pairs of one round of NPT points

// radiusSearch of one point over the 27 neighbouring voxels (the test ndt_point_eval<false> applies), cell by cell:
// one (point, neighbour cell) probe of that search: the voxel index, or -1
struct NdtPointCell {
  float3 xt;
  int cx, cy, cz;
  bool ok;
};
__device__ __forceinline__ NdtPointCell ndt_point_cell(const float4 p, const NdtEval& E, const NdtGridDev& G) {
  NdtPointCell r;
  r.ok = false;
  r.xt = make_float3(0.f, 0.f, 0.f);
  r.cx = r.cy = r.cz = 0;
  if (!finite3(p.x, p.y, p.z)) return r;
  r.xt = xform_point(E.T, p.x, p.y, p.z);
  if (!finite3(r.xt.x, r.xt.y, r.xt.z)) return r;
  r.cx = floor_to_int_x86(fmul(r.xt.x, G.inv_leaf));
  r.cy = floor_to_int_x86(fmul(r.xt.y, G.inv_leaf));
  r.cz = floor_to_int_x86(fmul(r.xt.z, G.inv_leaf));
  r.ok = grid_in_range(r.cx, r.cy, r.cz);
  return r;
}
__device__ __forceinline__ int ndt_probe(const NdtPointCell& pc, int c, const NdtGridDev& G, int tseg) {
  const int dz = c / 9 - 1, dy = (c / 3) % 3 - 1, dx = c % 3 - 1;  // c = ((dz+1)*3 + (dy+1))*3 + (dx+1), as ndt_candidates
  const int vi = ndt_lookup(G, ndt_key(tseg, pc.cx + dx, pc.cy + dy, pc.cz + dz));
  if (vi < 0) return -1;
  const VoxRec* V = &G.vox[vi];
  const float ddx = __fsub_rn(pc.xt.x, __ldg(&V->centroid[0])), ddy = __fsub_rn(pc.xt.y, __ldg(&V->centroid[1])),
              ddz = __fsub_rn(pc.xt.z, __ldg(&V->centroid[2]));
  const float dist = fadd(fadd(fmul(ddx, ddx), fmul(ddy, ddy)), fmul(ddz, ddz));
  return dist < G.r2 ? vi : -1;  // FLANN radius search: strict
}

struct NdtPersistSmem {
  NdtEval E;
  NdtState S;
  int pl_vi[NP_CAP];                 // pair list of the current round: voxel ...
  unsigned char pl_pt[NP_CAP];       // ... and the round-local index of its source point
  int wtot[NPT / 32];
  double red[NPT / 32][NACC];
  double sums[NACC];
  int finished;
};

__global__ void __launch_bounds__(NPT, 1)
k_ndt_persist(const float4* __restrict__ src, const int* __restrict__ count, int stride, NdtState* __restrict__ st,
              NdtEval* __restrict__ ev, NdtGridDev G, double d1, double d2, NdtCtl ctl, int max_evals, int n_seg, int cpp /* CTAs per pair */,
              double* __restrict__ gpart /* [2][n_seg][cpp][NACC] */, int* __restrict__ fin_eval /* [n_seg], preset to INT_MAX */,
              unsigned long long* __restrict__ work_count /* [0] (point, voxel) pairs of gradient evaluations, [1] of Hessian ones */) {
  __shared__ NdtPersistSmem M;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int seg = blockIdx.x / cpp, crank = blockIdx.x % cpp;  // (gridDim.x == n_seg * cpp)
  static_assert(sizeof(NdtState) % 4 == 0 && sizeof(NdtEval) % 4 == 0, "copied as 32-bit words");
  for (int k = tid; k < (int)(sizeof(NdtEval) / 4); k += NPT) ((unsigned*)&M.E)[k] = ((const unsigned*)&ev[seg])[k];
  for (int k = tid; k < (int)(sizeof(NdtState) / 4); k += NPT) ((unsigned*)&M.S)[k] = ((const unsigned*)&st[seg])[k];
  if (tid == 0) M.finished = 0;
  __syncthreads();
  const int n = count[seg];
  const int tseg = G.shared_target ? 0 : seg;
  const float4* P = src + (size_t)seg * stride;
  bool done = false;  // this pair has finished; its CTAs keep attending the grid barrier until every pair has
  unsigned long long pairs_g = 0, pairs_h = 0;
  for (int evals = 0; evals < max_evals; ++evals) {
    double* part = gpart + (((size_t)(evals & 1) * n_seg + seg) * cpp) * NACC;
    if (!done) {
      const bool wantH = M.E.want_hessian != 0;
      double acc[NACC];
#pragma unroll
      for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
      int np = 0;
      // Rounds of NPT points per CTA.  Pass 1 finds every point's voxels: the 27 x (points of the round) hash probes are
      // dealt out in contiguous chunks over ALL the threads -- a round of 40 points is 1080 probes, 5 per thread, not 27
      // dependent probe chains on each of 40 threads -- and the hits are compacted, in (point, cell) order, into the pair
      // list; pass 2 spreads the pairs evenly over the threads (running the ~500-operation derivative body inside the probe
      // loop would execute it once per probe and warp).  Pair j always goes to thread j % NPT: the fp64 sums keep a fixed
      // order.
      for (int base = crank; base < n; base += NPT * cpp) {
        const int npl = min(NPT, (n - base + cpp - 1) / cpp);  // the round's points: i = base + pl * cpp, pl < npl
        const int items = npl * 27;
        const int C = (items + NPT - 1) / NPT;  // <= 27
        const int j0 = min(tid * C, items), j1 = min(j0 + C, items);
        unsigned mask = 0u;  // bit q: probe j0 + q found a voxel within the resolution
        {
          int pl_cur = -1;
          NdtPointCell pc;
          pc.ok = false;
          for (int j = j0; j < j1; ++j) {
            const int pl = j / 27;
            if (pl != pl_cur) {
              pl_cur = pl;
              pc = ndt_point_cell(P[base + pl * cpp], M.E, G);
            }
            if (pc.ok && ndt_probe(pc, j - pl * 27, G, tseg) >= 0) mask |= 1u << (j - j0);
          }
        }
        const int mycnt = __popc(mask);
        int incl = mycnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        if (lane == 31) M.wtot[wid] = incl;
        __syncthreads();
        int woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NPT / 32; ++w) {
          const int c = M.wtot[w];
          woff += (w < wid) ? c : 0;
          total += c;
        }
        if (mask) {  // the hits again (cached lines), straight into the list
          int pos = woff + incl - mycnt;
          int pl_cur = -1;
          NdtPointCell pc;
          pc.ok = false;
          for (int q = 0; q < j1 - j0; ++q) {
            if (!(mask & (1u << q))) continue;
            const int j = j0 + q, pl = j / 27;
            if (pl != pl_cur) {
              pl_cur = pl;
              pc = ndt_point_cell(P[base + pl * cpp], M.E, G);
            }
            M.pl_vi[pos] = ndt_probe(pc, j - pl * 27, G, tseg);
            M.pl_pt[pos] = (unsigned char)pl;
            ++pos;
          }
        }
        __syncthreads();
        for (int j = tid; j < total; j += NPT) {
          const int ii = base + (int)M.pl_pt[j] * cpp;
          ndt_point_eval<true>(P[ii], M.E, G, tseg, wantH, d1, d2, acc, np, M.pl_vi[j]);
        }
        __syncthreads();  // the list is rewritten by the next round
      }
      if (wantH) pairs_h += np; else pairs_g += np;
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) M.red[wid][k] = v;
      }
      __syncthreads();
      if (tid < NACC) {
        double v = 0;
#pragma unroll
        for (int w = 0; w < NPT / 32; ++w) v += M.red[w][tid];
        part[(size_t)crank * NACC + tid] = v;
      }
      __threadfence();
    }
    grid.sync();
    {
      bool all = true;
      for (int sg = 0; sg < n_seg; ++sg) all = all && (__ldcg(&fin_eval[sg]) < evals);
      if (all) break;  // (identical on every thread of the grid)
    }
    if (!done) {
      {  // the pair's CTA partials, summed by 8 groups of lanes in a fixed order (identical on every CTA)
        const int k = tid & 31, grp = tid >> 5;
        double v = 0;
        if (k < NACC)
          for (int c = grp; c < cpp; c += NPT / 32) v += __ldcg(&part[(size_t)c * NACC + k]);
        __syncthreads();  // (M.red was last read before the grid barrier)
        if (k < NACC) M.red[grp][k] = v;
        __syncthreads();
        if (tid < NACC) {
          double t = 0;
#pragma unroll
          for (int w = 0; w < NPT / 32; ++w) t += M.red[w][tid];
          M.sums[tid] = t;
        }
      }
      __syncthreads();
      if (tid == 0) M.finished = ndt_control_step(&M.S, &M.E, M.sums, ctl) ? 1 : 0;
      __syncthreads();
      if (M.finished) {
        done = true;
        if (crank == 0) {
          for (int k = tid; k < (int)(sizeof(NdtState) / 4); k += NPT) ((unsigned*)&st[seg])[k] = ((const unsigned*)&M.S)[k];
          for (int k = tid; k < (int)(sizeof(NdtEval) / 4); k += NPT) ((unsigned*)&ev[seg])[k] = ((const unsigned*)&M.E)[k];
          if (tid == 0) {
            __threadfence();
            fin_eval[seg] = evals;
          }
        }
      }
    }
    // Termination: a pair's first CTA records the evaluation at which the pair finished; everybody reads the records after
    // the NEXT barrier and only trusts entries <= the previous evaluation, which are complete by then -- every CTA takes
    // the same decision and leaves at the same barrier (one barrier per evaluation, one extra at the end).
  }
  if (!done && crank == 0) {  // evaluation budget exhausted: hand the state back as it stands
    for (int k = tid; k < (int)(sizeof(NdtState) / 4); k += NPT) ((unsigned*)&st[seg])[k] = ((const unsigned*)&M.S)[k];
    for (int k = tid; k < (int)(sizeof(NdtEval) / 4); k += NPT) ((unsigned*)&ev[seg])[k] = ((const unsigned*)&M.E)[k];
  }
  for (int o = 16; o > 0; o >>= 1) {
    pairs_g += __shfl_down_sync(0xffffffffu, pairs_g, o);
    pairs_h += __shfl_down_sync(0xffffffffu, pairs_h, o);
  }
  if (lane == 0) {
    atomicAdd(&work_count[0], pairs_g);
    atomicAdd(&work_count[1], pairs_h);
  }
}

// point-sharded mode, peer-memory path (see k_icp_solve_peer): CTA partials -> one-shot exchange of the 28 totals -> controller
__global__ void __launch_bounds__(32) k_ndt_control_peer(NdtState* __restrict__ st, NdtEval* __restrict__ ev,
                                                         const double* __restrict__ partials, int nblk, NdtCtl ctl,
                                                         int* __restrict__ n_active, PeerX X) {
  const int seg = blockIdx.x;
  if (st[seg].done) return;
  const int lane = threadIdx.x;
  __shared__ double sums[NACC];
  double v = 0;
  if (lane < NACC)
    for (int b = 0; b < nblk; ++b) v += partials[((size_t)seg * nblk + b) * NACC + lane];
  v = peer_allreduce_warp(X, v, lane, NACC, seg);
  if (lane < NACC) sums[lane] = v;
  __syncwarp();
  bool finished = false;
  if (lane == 0) finished = ndt_control_step(&st[seg], &ev[seg], sums, ctl);
  if (finished) atomicSub(n_active, 1);
}

__global__ void k_ndt_gather_T(const NdtState* __restrict__ st, float* __restrict__ T, int n_seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_seg * 16) T[i] = st[i / 16].final_T[i % 16];
}

__global__ void k_ndt_eval_setup(NdtEval* __restrict__ ev, const double* __restrict__ poses, int n_seg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  NdtEval* E = &ev[s];
  pose_to_matrix(poses + 6 * s, E->T);
  angle_derivatives(poses + 6 * s, E);
  E->want_hessian = 1;
  E->hessian_only = 0;
  E->active = 1;
  E->pad = 0;
}

// point-sharded mode: per-pair totals of this rank's shard (zeros for finished pairs), all-reduced before the controller
__global__ void k_ndt_sum_partials_active(const double* __restrict__ partials, int nblk, const NdtEval* __restrict__ ev,
                                          double* __restrict__ totals) {
  const int seg = blockIdx.x, lane = threadIdx.x;
  if (lane >= NACC) return;
  double v = 0;
  if (ev[seg].active)
    for (int b = 0; b < nblk; ++b) v += partials[((size_t)seg * nblk + b) * NACC + lane];
  totals[seg * NACC + lane] = v;
}

__global__ void k_ndt_sum_partials(const double* __restrict__ partials, int nblk, int n_seg, double* __restrict__ out) {
  const int seg = blockIdx.x, lane = threadIdx.x;
  if (lane >= NACC) return;
  double v = 0;
  for (int b = 0; b < nblk; ++b) v += partials[((size_t)seg * nblk + b) * NACC + lane];
  out[seg * NACC + lane] = v;
}

void gauss_constants(float resolution, double outlier_ratio, double* d1, double* d2) {
  const double c1 = 10.0 * (1 - outlier_ratio);
  const double c2 = outlier_ratio / pow((double)resolution, 3);
  const double d3 = -log(c2);
  *d1 = -log(c1 + c2) - d3;
  *d2 = -2 * log((-log(c1 * exp(-0.5) + c2) - d3) / *d1);
}

int ndt_grid_build(rspcl_ctx* ctx, const rspcl_cloud* tgt, const rspcl_ndt_params* prm, NdtGridDev* G, int* d_range) {
  const int S = tgt->n_seg;
  const int pstride = tgt->max_count_hint > 0 ? tgt->max_count_hint : 1;
  const long long N = (long long)S * pstride;
  G->inv_leaf = 1.0f / prm->resolution;
  G->r2 = (float)((double)prm->resolution * (double)prm->resolution);
  Scratch scr(ctx);  // the temporaries; the grid itself (G->...) is owned by the caller (ndt_grid_free)
  unsigned long long *keys = nullptr, *tkeys = nullptr;
  int *vals = nullptr, *tvals = nullptr, *flags = nullptr, *ord = nullptr, *cell_start = nullptr, *leaf_flag = nullptr,
      *leaf_ord = nullptr, *n_cells = nullptr, *n_valid = nullptr;
  CU(ctx, scr.alloc(&keys, (size_t)N));
  CU(ctx, scr.alloc(&tkeys, (size_t)N));
  CU(ctx, scr.alloc(&vals, (size_t)N));
  CU(ctx, scr.alloc(&tvals, (size_t)N));
  unsigned long long* flags_keys = nullptr;  // the one-CTA build's 64-bit keys at the cell heads
  CU(ctx, scr.alloc(&flags_keys, (size_t)N));
  CU(ctx, scr.alloc(&flags, (size_t)N));
  CU(ctx, scr.alloc(&ord, (size_t)N));
  CU(ctx, scr.alloc(&cell_start, (size_t)N + 1));
  CU(ctx, scr.alloc(&leaf_flag, (size_t)N));
  CU(ctx, scr.alloc(&leaf_ord, (size_t)N));
  CU(ctx, scr.alloc(&n_cells, 1));
  CU(ctx, scr.alloc(&n_valid, 1));
  CU(ctx, scratch_alloc(ctx, &G->n_leaves, 1));
  int nb = div_up(N, 256);
  if (nb > 16 * ctx->sm_count) nb = 16 * ctx->sm_count;
  int rc = RSPCL_OK;
  const char* benv = getenv("RSPCL_NDT_BUILD_ONE");
  unsigned long long* head_keys = keys;  // where k_voxel_stats finds the key of a cell's first point
  int* sorted_vals = vals;
  if (S == 1 && N <= NB1_MAX && !(benv && benv[0] == '0')) {
    // one target of a few ten thousand points: the whole build up to the voxel statistics in ONE launch
    CU(ctx, cudaFuncSetAttribute(k_ndt_build_one, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NB1_SMEM_BYTES));
    k_ndt_build_one<<<1, NB1_T, NB1_SMEM_BYTES, ctx->stream>>>(tgt->pts, tgt->count, (int)N, G->inv_leaf,
                                                                prm->min_points_per_voxel, keys, tkeys, vals, tvals,
                                                                reinterpret_cast<unsigned long long*>(flags_keys), ord,
                                                                cell_start, leaf_flag, leaf_ord, n_cells, G->n_leaves, d_range);
    LAUNCH_CHECK(ctx);
    head_keys = reinterpret_cast<unsigned long long*>(flags_keys);
    sorted_vals = ord;
  } else {
    CU(ctx, cudaMemsetAsync(n_valid, 0, sizeof(int), ctx->stream));
    dim3 gk(blocks_per_seg(ctx, S, pstride, 256), S);
    k_ndt_keys<<<gk, 256, 0, ctx->stream>>>(tgt->pts, tgt->count, tgt->stride, pstride, G->inv_leaf, keys, vals, d_range);
    LAUNCH_CHECK(ctx);
    rc = radix_sort_pairs(ctx, keys, vals, tkeys, tvals, N);
    if (rc) return rc;
    k_cell_heads<<<nb, 256, 0, ctx->stream>>>(keys, N, flags, n_valid);
    LAUNCH_CHECK(ctx);
    rc = rspcl_exclusive_scan_i32(ctx, flags, ord, N, n_cells);
    if (rc) return rc;
    k_cell_starts<<<nb, 256, 0, ctx->stream>>>(flags, ord, N, n_cells, n_valid, cell_start);
    LAUNCH_CHECK(ctx);
    k_leaf_flags<<<nb, 256, 0, ctx->stream>>>(cell_start, n_cells, prm->min_points_per_voxel, leaf_flag, N);
    LAUNCH_CHECK(ctx);
    rc = rspcl_exclusive_scan_i32(ctx, leaf_flag, leaf_ord, N, G->n_leaves);
    if (rc) return rc;
  }
  const int minp = prm->min_points_per_voxel > 0 ? prm->min_points_per_voxel : 1;
  G->vox_cap = N / minp + 1;
  unsigned cap = 1024;
  while ((long long)cap < 2 * G->vox_cap && cap < (1u << 30)) cap <<= 1;
  G->cap_mask = cap - 1;
  CU(ctx, scratch_alloc(ctx, &G->vox, (size_t)G->vox_cap));
  CU(ctx, scratch_alloc(ctx, &G->hkeys, (size_t)cap));
  CU(ctx, scratch_alloc(ctx, &G->hvals, (size_t)cap));
  CU(ctx, scratch_alloc(ctx, &G->nvox_seg, (size_t)S));
  CU(ctx, cudaMemsetAsync(G->hkeys, 0xFF, (size_t)cap * sizeof(unsigned long long), ctx->stream));
  CU(ctx, cudaMemsetAsync(G->nvox_seg, 0, (size_t)S * sizeof(int), ctx->stream));
  int nbw = div_up(N, 8);  // one warp per cell, grid-stride over the cells
  if (nbw > 8 * ctx->sm_count) nbw = 8 * ctx->sm_count;
  k_voxel_stats<<<nbw, 256, 0, ctx->stream>>>(tgt->pts, tgt->stride, pstride, head_keys, sorted_vals, cell_start, leaf_flag, leaf_ord, n_cells,
                                             prm->min_covar_eigvalue_mult, G->vox, G->hkeys, G->hvals, G->cap_mask, G->nvox_seg);
  LAUNCH_CHECK(ctx);
  scr.ok();
  return RSPCL_OK;
}

void ndt_grid_free(rspcl_ctx* ctx, NdtGridDev* G) {
  scratch_free(ctx, G->vox);
  scratch_free(ctx, G->hkeys);
  scratch_free(ctx, G->hvals);
  scratch_free(ctx, G->nvox_seg);
  scratch_free(ctx, G->n_leaves);
}

}  // namespace

__global__ void k_counts_f64(const int* __restrict__ count, double* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)count[i];
}

extern "C" void rspcl_ndt_reference_params(rspcl_ndt_params* p) {
  p->max_iterations = 50;            // ndt:43
  p->min_points_per_voxel = 6;
  p->transformation_epsilon = 0.01;  // ndt:39
  p->step_size = 0.1;                // ndt:40
  p->outlier_ratio = 0.55;
  p->min_covar_eigvalue_mult = 0.01;
  p->resolution = 1.0f;              // ndt:41
  p->_pad = 0;
}

int ndt_align_device(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt, const rspcl_ndt_params* prm,
                     const float* d_guess, rspcl_ndt_result* h_results, rspcl_cloud* aligned) {
  const int S = src->n_seg;
  const int shared_target = (tgt->n_seg == 1 && S > 1) ? 1 : 0;
  if (!shared_target && tgt->n_seg != S) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "ndt_align: src has %d segments, tgt %d", S, tgt->n_seg);
  if (aligned && (aligned->n_seg != S || aligned->stride < src->max_count_hint))
    RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "ndt_align: aligned output too small");
  if (!(prm->resolution > 0.f)) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "ndt_align: resolution must be positive");
  double d1, d2;
  gauss_constants(prm->resolution, prm->outlier_ratio, &d1, &d2);
  int* d_range = nullptr;
  Scratch scr(ctx);  // every temporary below; released on every exit path
  CU(ctx, scr.alloc(&d_range, 1));
  CU(ctx, cudaMemsetAsync(d_range, 0, sizeof(int), ctx->stream));
  NdtGridDev G;
  struct GridGuard {
    rspcl_ctx* c;
    NdtGridDev* g;
    ~GridGuard() { ndt_grid_free(c, g); }
  } grid_guard{ctx, &G};
  G.shared_target = shared_target;
  int rc;
  {
    ProfScope prof(ctx, "ndt_voxel_build", (double)tgt->n_seg * tgt->max_count_hint);
    rc = ndt_grid_build(ctx, tgt, prm, &G, d_range);
  }
  if (rc) return rc;

  NdtState* st = nullptr;
  NdtEval* ev = nullptr;
  double* partials = nullptr;
  int* n_active = nullptr;
  float* d_T = nullptr;
  const int nblk = blocks_per_seg(ctx, S, src->max_count_hint, NT);
  CU(ctx, scr.alloc(&st, (size_t)S));
  CU(ctx, scr.alloc(&ev, (size_t)S));
  CU(ctx, scr.alloc(&partials, (size_t)S * nblk * NACC));
  CU(ctx, scr.alloc(&n_active, 1));
  CU(ctx, scr.alloc(&d_T, (size_t)S * 16));
  const bool sharded = ctx->sharded_call && ctx->nccl_comm && ctx->nranks > 1;
  double* totals = nullptr;
  if (sharded) CU(ctx, scr.alloc(&totals, (size_t)S * NACC));
  CU(ctx, small_h2d(ctx, n_active, &S, sizeof(int)));
  k_ndt_init<<<div_up(S, 64), 64, 0, ctx->stream>>>(st, ev, d_guess, S);
  LAUNCH_CHECK(ctx);
  NdtCtl ctl;
  ctl.max_iterations = prm->max_iterations;
  ctl.eps = prm->transformation_epsilon;
  ctl.step_size = prm->step_size;
  // worst case per Newton iteration: 1 + 10 line-search evaluations + 1 Hessian; (max_iterations + 2) iterations
  const long long max_evals = 1 + (long long)(prm->max_iterations + 2) * 12;
  long long done_evals = 0;
  int chunk = 4, active = S;
  dim3 ge(nblk, S);
  const char* penv = getenv("RSPCL_NDT_PERSIST");
  // a few pairs: the whole align in ONE cooperative launch that spans the chip (k_ndt_persist); batches of many pairs keep
  // the per-evaluation launches below, which fill the chip with two CTAs per SM and amortise their launches over the batch
  if (!sharded && !(penv && penv[0] == '0') && S * 8 <= ctx->sm_count) {
    int per_sm = 0;
    CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ndt_persist, NPT, 0));
    int cpp = per_sm > 0 ? (ctx->sm_count * per_sm) / S : 0;
    const int want = div_up(src->max_count_hint > 0 ? src->max_count_hint : 1, 32);  // >= 32 points per CTA
    if (cpp > want) cpp = want;
    if (cpp >= 1) {
      unsigned long long* d_work = nullptr;
      double* d_gpart = nullptr;
      int* d_fin = nullptr;
      CU(ctx, scr.alloc(&d_work, 2));
      CU(ctx, scr.alloc(&d_gpart, (size_t)2 * S * cpp * NACC));
      CU(ctx, scr.alloc(&d_fin, (size_t)S));
      CU(ctx, cudaMemsetAsync(d_work, 0, 2 * sizeof(unsigned long long), ctx->stream));
      CU(ctx, cudaMemsetAsync(d_fin, 0x7F, (size_t)S * sizeof(int), ctx->stream));  // 0x7F7F7F7F: "not finished"
      const float4* a_src = src->pts;
      const int* a_cnt = src->count;
      int a_stride = src->stride, a_max = (int)max_evals, a_S = S;
      void* args[] = {(void*)&a_src, (void*)&a_cnt, (void*)&a_stride, (void*)&st, (void*)&ev, (void*)&G, (void*)&d1, (void*)&d2,
                      (void*)&ctl, (void*)&a_max, (void*)&a_S, (void*)&cpp, (void*)&d_gpart, (void*)&d_fin, (void*)&d_work};
      unsigned long long hw[2] = {0, 0};
      {
        ProfScope prof(ctx, "k_ndt_persist", 0.0);
        CU(ctx, cudaLaunchCooperativeKernel((const void*)k_ndt_persist, dim3((unsigned)(S * cpp)), dim3(NPT), args, 0, ctx->stream));
        LAUNCH_CHECK(ctx);
        prof.end();
        if (ctx->prof_on) {
          CU(ctx, small_d2h(ctx, hw, d_work, sizeof(hw)));
          CU(ctx, ctx_sync(ctx));
          // FP64 operations: per (point, voxel) pair ~135 for a gradient evaluation, ~456 with the Hessian (counted from
          // ndt_point_eval); reported as "equivalent gradient pairs" so that one number carries both
          prof.set_units((double)hw[0] + (double)hw[1] * (456.0 / 135.0));
        }
      }
      active = 0;
    }
  }
  while (active > 0 && done_evals < max_evals) {
    for (int k = 0; k < chunk; ++k) {
      {
        ProfScope prof(ctx, "k_ndt_eval", (double)S * src->max_count_hint);
        k_ndt_eval<<<ge, NT, 0, ctx->stream>>>(src->pts, src->count, src->stride, ev, G, d1, d2, partials);
        LAUNCH_CHECK(ctx);
      }
      ProfScope prof_c(ctx, "k_ndt_control", (double)S);
      if (sharded && comm_peer_ready(ctx, S, NACC)) {
        k_ndt_control_peer<<<S, 32, 0, ctx->stream>>>(st, ev, partials, nblk, ctl, n_active, comm_peer_next(ctx));
      } else if (sharded) {
        k_ndt_sum_partials_active<<<S, 32, 0, ctx->stream>>>(partials, nblk, ev, totals);
        LAUNCH_CHECK(ctx);
        int rcc = comm_allreduce_f64(ctx, totals, (size_t)S * NACC);
        if (rcc) return rcc;
        k_ndt_control<<<S, 32, 0, ctx->stream>>>(st, ev, totals, 1, ctl, n_active);
      } else {
        k_ndt_control<<<S, 32, 0, ctx->stream>>>(st, ev, partials, nblk, ctl, n_active);
      }
      LAUNCH_CHECK(ctx);
    }
    done_evals += chunk;
    CU(ctx, small_d2h(ctx, &active, n_active, sizeof(int)));
    CU(ctx, ctx_sync(ctx));
    if (chunk < 32) chunk *= 2;
  }
  std::vector<NdtState> hst(S);
  int range = 0;
  CU(ctx, small_d2h(ctx, hst.data(), st, (size_t)S * sizeof(NdtState)));
  CU(ctx, small_d2h(ctx, &range, d_range, sizeof(int)));
  if (aligned) {
    k_ndt_gather_T<<<div_up(S * 16, 256), 256, 0, ctx->stream>>>(st, d_T, S);
    LAUNCH_CHECK(ctx);
    rc = transform_device(ctx, src, d_T, 0, aligned);
  }
  // source sizes for trans_probability = score / input_->size() (ndt.hpp); in point-sharded mode that is the size of the
  // WHOLE source, i.e. the shard counts summed over the ranks -- every rank then reports the same value
  std::vector<int> scnt(S);
  if (!sharded) CU(ctx, small_d2h(ctx, scnt.data(), src->count, S * sizeof(int)));  // rides on the same synchronisation
  CU(ctx, ctx_sync(ctx));
  if (sharded) {
    std::vector<double> cd(S);
    double* d_cnt = nullptr;
    CU(ctx, scr.alloc(&d_cnt, (size_t)S));
    k_counts_f64<<<div_up(S, 128), 128, 0, ctx->stream>>>(src->count, d_cnt, S);
    LAUNCH_CHECK(ctx);
    int rcc = comm_allreduce_f64(ctx, d_cnt, (size_t)S);
    if (rcc) return rcc;
    CU(ctx, small_d2h(ctx, cd.data(), d_cnt, (size_t)S * sizeof(double)));
    CU(ctx, ctx_sync(ctx));
    for (int s = 0; s < S; ++s) scnt[s] = (int)(cd[s] + 0.5);
  }
  for (int s = 0; s < S; ++s) {
    memcpy(h_results[s].T, hst[s].final_T, 64);
    h_results[s].converged = hst[s].converged;
    h_results[s].iterations = hst[s].nr_iterations;
    h_results[s].n_derivative_evals = hst[s].n_deriv;
    h_results[s].n_hessian_evals = hst[s].n_hess;
    h_results[s].score = hst[s].score;
    h_results[s].trans_probability = hst[s].score / (double)(scnt[s] > 0 ? scnt[s] : 1);
    memcpy(h_results[s].p, hst[s].p, sizeof(double) * 6);
  }
  if (rc) return rc;
  if (range) RSPCL_FAIL(ctx, RSPCL_ERR_RANGE, "ndt_align: target coordinates exceed the voxel key range");
  scr.ok();
  return RSPCL_OK;
}

extern "C" int rspcl_ndt_align(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt, const rspcl_ndt_params* prm,
                               const float* guess, rspcl_ndt_result* results, rspcl_cloud* aligned) {
  if (!ctx || !src || !tgt || !prm || !results) return RSPCL_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  int rc = refresh_count_hint(ctx, src, nullptr);
  if (rc) return rc;
  rc = refresh_count_hint(ctx, tgt, nullptr);
  if (rc) return rc;
  float* d_guess = nullptr;
  if (guess) {
    CU(ctx, scratch_alloc(ctx, &d_guess, (size_t)src->n_seg * 16));
    CU(ctx, small_h2d(ctx, d_guess, guess, (size_t)src->n_seg * 16 * sizeof(float)));
  }
  rc = ndt_align_device(ctx, src, tgt, prm, d_guess, results, aligned);
  scratch_free(ctx, d_guess);
  return rc;
}

extern "C" int rspcl_ndt_voxels(rspcl_ctx* ctx, const rspcl_cloud* tgt, const rspcl_ndt_params* prm, void* host_records,
                                long long capacity, int32_t* n_vox) {
  if (!ctx || !tgt || !prm || !host_records || !n_vox) return RSPCL_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  int rc = refresh_count_hint(ctx, tgt, nullptr);
  if (rc) return rc;
  int* d_range = nullptr;
  CU(ctx, scratch_alloc(ctx, &d_range, 1));
  CU(ctx, cudaMemsetAsync(d_range, 0, sizeof(int), ctx->stream));
  NdtGridDev G;
  rc = ndt_grid_build(ctx, tgt, prm, &G, d_range);
  if (rc) return rc;
  int total = 0;
  CU(ctx, small_d2h(ctx, &total, G.n_leaves, sizeof(int)));
  CU(ctx, small_d2h(ctx, n_vox, G.nvox_seg, tgt->n_seg * sizeof(int)));
  CU(ctx, ctx_sync(ctx));
  if (total > capacity) {
    ndt_grid_free(ctx, &G);
    scratch_free(ctx, d_range);
    RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "ndt_voxels: %d voxels, capacity %lld", total, capacity);
  }
  if (total) CU(ctx, small_d2h(ctx, host_records, G.vox, (size_t)total * sizeof(VoxRec)));
  CU(ctx, ctx_sync(ctx));
  ndt_grid_free(ctx, &G);
  scratch_free(ctx, d_range);
  return RSPCL_OK;
}

extern "C" int rspcl_ndt_derivatives(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt,
                                     const rspcl_ndt_params* prm, const double* p, double* score, double* g, double* H) {
  if (!ctx || !src || !tgt || !prm || !p || !score || !g || !H) return RSPCL_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  int rc = refresh_count_hint(ctx, src, nullptr);
  if (rc) return rc;
  rc = refresh_count_hint(ctx, tgt, nullptr);
  if (rc) return rc;
  const int S = src->n_seg;
  double d1, d2;
  gauss_constants(prm->resolution, prm->outlier_ratio, &d1, &d2);
  int* d_range = nullptr;
  CU(ctx, scratch_alloc(ctx, &d_range, 1));
  CU(ctx, cudaMemsetAsync(d_range, 0, sizeof(int), ctx->stream));
  NdtGridDev G;
  G.shared_target = (tgt->n_seg == 1 && S > 1) ? 1 : 0;
  rc = ndt_grid_build(ctx, tgt, prm, &G, d_range);
  if (rc) return rc;
  NdtEval* ev = nullptr;
  double *d_p = nullptr, *partials = nullptr, *d_out = nullptr;
  const int nblk = blocks_per_seg(ctx, S, src->max_count_hint, NT);
  CU(ctx, scratch_alloc(ctx, &ev, (size_t)S));
  CU(ctx, scratch_alloc(ctx, &d_p, (size_t)S * 6));
  CU(ctx, scratch_alloc(ctx, &partials, (size_t)S * nblk * NACC));
  CU(ctx, scratch_alloc(ctx, &d_out, (size_t)S * NACC));
  CU(ctx, small_h2d(ctx, d_p, p, (size_t)S * 6 * sizeof(double)));
  k_ndt_eval_setup<<<div_up(S, 64), 64, 0, ctx->stream>>>(ev, d_p, S);
  LAUNCH_CHECK(ctx);
  k_ndt_eval<<<dim3(nblk, S), NT, 0, ctx->stream>>>(src->pts, src->count, src->stride, ev, G, d1, d2, partials);
  LAUNCH_CHECK(ctx);
  k_ndt_sum_partials<<<S, 32, 0, ctx->stream>>>(partials, nblk, S, d_out);
  LAUNCH_CHECK(ctx);
  std::vector<double> out((size_t)S * NACC);
  CU(ctx, small_d2h(ctx, out.data(), d_out, out.size() * sizeof(double)));
  CU(ctx, ctx_sync(ctx));
  for (int s = 0; s < S; ++s) {
    const double* o = &out[(size_t)s * NACC];
    score[s] = o[0];
    for (int i = 0; i < 6; ++i) g[s * 6 + i] = o[1 + i];
    int h = 7;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) {
        H[s * 36 + i * 6 + j] = o[h];
        H[s * 36 + j * 6 + i] = o[h];
        ++h;
      }
  }
  ndt_grid_free(ctx, &G);
  scratch_free(ctx, ev);
  scratch_free(ctx, d_p);
  scratch_free(ctx, partials);
  scratch_free(ctx, d_out);
  scratch_free(ctx, d_range);
  return RSPCL_OK;
}

extern "C" int rspcl_ndt_align_sharded(rspcl_ctx* ctx, const rspcl_cloud* src_shard, const rspcl_cloud* tgt,
                                       const rspcl_ndt_params* prm, const float* guess, rspcl_ndt_result* results,
                                       rspcl_cloud* aligned) {
  if (!ctx) return RSPCL_ERR_ARG;
  ctx->sharded_call = true;
  const int rc = rspcl_ndt_align(ctx, src_shard, tgt, prm, guess, results, aligned);
  ctx->sharded_call = false;
  return rc;
}
