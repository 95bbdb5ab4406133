// sort.cu -- stable LSD radix sort of (uint64 key, int32 value) pairs, 8-bit digits, hand-written for sm_100a.
//
// Used for the "radix sort by cell key" of the voxel structures (BASELINE north_star item 2): NDT voxel build
// (VoxelGridCovariance) sorts target points by (segment, iz, iy, ix) so every voxel is a contiguous, input-ordered
// run -- which makes the fp64 voxel sums deterministic and bit-identical to a sequential pass.
// Per pass: CTA digit histograms (digit-major) -> exclusive scan -> stable scatter.  Stability inside a CTA comes
// from warp-private cursors: warp w owns the w-th contiguous slice of the CTA's tile and walks it in order, ranking
// equal digits with __match_any_sync.  Passes whose digit is constant over the whole input are skipped (one upfront
// histogram of all 8 digit positions), so a 64-bit key costs only as many passes as it has varying bytes.
// HBM traffic per executed pass: 12 B read (histogram re-reads keys: +8 B) + 12 B written per element.
#include "common.cuh"

namespace {

constexpr int ST = 256, SW = ST / 32, TILE = 4096, RADIX = 256;

__global__ void __launch_bounds__(ST) k_digit_presence(const unsigned long long* __restrict__ keys, long long n,
                                                       unsigned* __restrict__ hist /* [8][256] */) {
  __shared__ unsigned h[8][RADIX];
  for (int k = threadIdx.x; k < 8 * RADIX; k += ST) (&h[0][0])[k] = 0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * ST + threadIdx.x; i < n; i += (long long)gridDim.x * ST) {
    const unsigned long long k = keys[i];
#pragma unroll
    for (int d = 0; d < 8; ++d) h[d][(k >> (8 * d)) & 255] = 1;  // presence only (benign races)
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 8 * RADIX; k += ST)
    if ((&h[0][0])[k]) hist[k] = 1;
}

__global__ void __launch_bounds__(ST) k_block_hist(const unsigned long long* __restrict__ keys, long long n, int shift,
                                                   int* __restrict__ bh /* [256][nblk] */, int nblk) {
  __shared__ int h[RADIX];
  h[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * TILE;
  for (int k = threadIdx.x; k < TILE; k += ST) {
    const long long i = base + k;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255], 1);
  }
  __syncthreads();
  bh[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(ST) k_radix_scatter(const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                                      long long n, int shift, const int* __restrict__ bh_scanned, int nblk,
                                                      unsigned long long* __restrict__ keys_out, int* __restrict__ vals_out) {
  __shared__ int cur[SW][RADIX];  // warp-private write cursors
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < SW * RADIX; k += ST) (&cur[0][0])[k] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * TILE;
  const int tile_n = (int)min((long long)TILE, n - base);
  const int L = TILE / SW;
  const int w_lo = min(wid * L, tile_n), w_hi = min(w_lo + L, tile_n);
  for (int k = w_lo + lane; k < w_hi; k += 32) atomicAdd(&cur[wid][(keys[base + k] >> shift) & 255], 1);
  __syncthreads();
  {  // digit d = threadIdx.x: global offset of (d, this CTA) then prefix over the warps
    int run = bh_scanned[(size_t)threadIdx.x * nblk + blockIdx.x];
    for (int w = 0; w < SW; ++w) {
      const int t = cur[w][threadIdx.x];
      cur[w][threadIdx.x] = run;
      run += t;
    }
  }
  __syncthreads();
  for (int b = w_lo; b < w_hi; b += 32) {
    const int k = b + lane;
    const bool act = k < w_hi;
    const unsigned amask = __ballot_sync(0xffffffffu, act);
    if (act) {
      const unsigned long long key = keys[base + k];
      const unsigned d = (unsigned)(key >> shift) & 255u;
      const unsigned peers = __match_any_sync(amask, d);
      const int leader = __ffs(peers) - 1;
      const int rank = __popc(peers & ((1u << lane) - 1u));
      int old = 0;
      if (lane == leader) {
        old = cur[wid][d];
        cur[wid][d] = old + __popc(peers);
      }
      old = __shfl_sync(peers, old, leader);
      keys_out[old + rank] = key;
      vals_out[old + rank] = vals[base + k];
    }
    __syncwarp();
  }
}

}  // namespace

// Sorts n pairs by key (stable).  keys/vals are overwritten with the sorted result; tmp_* are n-element scratch.
int radix_sort_pairs(rspcl_ctx* ctx, unsigned long long* keys, int* vals, unsigned long long* tmp_keys, int* tmp_vals,
                     long long n) {
  if (n <= 1) return RSPCL_OK;
  unsigned* d_presence = nullptr;
  CU(ctx, scratch_alloc(ctx, &d_presence, (size_t)8 * RADIX));
  CU(ctx, cudaMemsetAsync(d_presence, 0, 8 * RADIX * sizeof(unsigned), ctx->stream));
  int pb = div_up(n, ST);
  if (pb > 8 * ctx->sm_count) pb = 8 * ctx->sm_count;
  k_digit_presence<<<pb, ST, 0, ctx->stream>>>(keys, n, d_presence);
  LAUNCH_CHECK(ctx);
  unsigned h_presence[8 * RADIX];
  CU(ctx, small_d2h(ctx, h_presence, d_presence, sizeof(h_presence)));
  CU(ctx, ctx_sync(ctx));
  scratch_free(ctx, d_presence);
  const int nblk = div_up(n, TILE);
  int* bh = nullptr;
  CU(ctx, scratch_alloc(ctx, &bh, (size_t)RADIX * nblk));
  unsigned long long *src_k = keys, *dst_k = tmp_keys;
  int *src_v = vals, *dst_v = tmp_vals;
  for (int d = 0; d < 8; ++d) {
    int distinct = 0;
    for (int b = 0; b < RADIX; ++b) distinct += h_presence[d * RADIX + b] ? 1 : 0;
    if (distinct <= 1) continue;  // constant digit: pass is the identity
    k_block_hist<<<nblk, ST, 0, ctx->stream>>>(src_k, n, 8 * d, bh, nblk);
    LAUNCH_CHECK(ctx);
    int rc = rspcl_exclusive_scan_i32(ctx, bh, bh, (long long)RADIX * nblk, nullptr);
    if (rc) return rc;
    k_radix_scatter<<<nblk, ST, 0, ctx->stream>>>(src_k, src_v, n, 8 * d, bh, nblk, dst_k, dst_v);
    LAUNCH_CHECK(ctx);
    unsigned long long* tk = src_k;
    src_k = dst_k;
    dst_k = tk;
    int* tv = src_v;
    src_v = dst_v;
    dst_v = tv;
  }
  if (src_k != keys) {
    CU(ctx, cudaMemcpyAsync(keys, src_k, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(vals, src_v, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  scratch_free(ctx, bh);
  return RSPCL_OK;
}
