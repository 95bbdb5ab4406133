// voxel.cu -- pcl::ApproximateVoxelGrid<PointXYZRGB>::applyFilter, order- and bit-exact, as one CTA per cloud.
//
// Reference call sites: icp:37,47,59-60,75-76; ndt:34,45,57-58,68-69; incr:36,54-55.
// The PCL filter is a sequential stream over a 512-slot direct-mapped history: a point whose voxel differs from
// the one parked in its slot flushes that slot (emitting the partial centroid) before being accumulated, and the
// slots still occupied at the end are flushed in slot order.  The exact parallel form used here:
//   1. stable counting sort of the points by slot h = (7171 ix + 3079 iy + 4231 iz) & 511 -- warp-private
//      histograms in shared memory, warp w owning the w-th contiguous slice of the input, so input order is kept
//      inside every slot;
//   2. inside a slot, a maximal run of equal (ix,iy,iz) is exactly one flush; each run is summed sequentially in
//      input order by one thread (float, un-fused) so the centroid bits match the stream;
//   3. a run that is evicted by input point j is emitted at position #{evicting points < j}; runs that survive to
//      the end follow in slot order.
// HBM-bound and tiny (~10^4 points per cloud): 16 B read per point for keys, gathers for the run sums.
#include "common.cuh"

namespace {

constexpr int VT = 1024, VW = VT / 32, NSLOT = 512;

struct VoxKey {
  int ix, iy, iz;
  unsigned slot;
};

__device__ __forceinline__ VoxKey vox_key(float4 p, float3 inv) {
  VoxKey k;
  k.ix = floor_to_int_x86(fmul(p.x, inv.x));
  k.iy = floor_to_int_x86(fmul(p.y, inv.y));
  k.iz = floor_to_int_x86(fmul(p.z, inv.z));
  k.slot = ((unsigned)k.ix * 7171u + (unsigned)k.iy * 3079u + (unsigned)k.iz * 4231u) & (NSLOT - 1);
  return k;
}

// block-wide exclusive scan of data[0..n) in place (global or shared), returns the total to every thread.  Four
// consecutive elements per thread and round: a quarter of the rounds (and barriers) of a one-element scan.
__device__ int block_excl_scan(int* data, int n, int* s_warp, int* s_carry) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) *s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 4 * VT) {
    const int i = base + 4 * threadIdx.x;
    int v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (i + u < n) ? data[i + u] : 0;
    const int mine = v[0] + v[1] + v[2] + v[3];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int wv = s_warp[lane], wi = wv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      s_warp[lane] = wi - wv;
    }
    __syncthreads();
    int run = *s_carry + s_warp[wid] + incl - mine;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i + u < n) data[i + u] = run;
      run += v[u];
    }
    __syncthreads();
    if (threadIdx.x == VT - 1) *s_carry = run;
    __syncthreads();
  }
  const int total = *s_carry;
  __syncthreads();  // nobody may reset the carry (next call) before everyone has read it
  return total;
}

__global__ void __launch_bounds__(VT) k_approx_voxel(const float4* __restrict__ pts, const int* count,
                                                     int stride_in, float3 inv, int* __restrict__ sorted_all,
                                                     float4* __restrict__ spt_all, int* __restrict__ ev_all,
                                                     float4* __restrict__ out,
                                                     int* out_count, int stride_out,
                                                     int* __restrict__ overflow) {
  extern __shared__ int smem[];
  int* hist = smem;                         // [VW][NSLOT] warp-private counters
  int* slot_start = hist + VW * NSLOT;      // [NSLOT]
  int* slot_end = slot_start + NSLOT;       // [NSLOT]
  int* slot_rank = slot_end + NSLOT;        // [NSLOT]
  int* s_warp = slot_rank + NSLOT;          // [32]
  int* s_misc = s_warp + 32;                // [4]
  const int seg = blockIdx.x;
  const int n = count[seg];
  const float4* P = pts + (size_t)seg * stride_in;
  int* sorted = sorted_all + (size_t)seg * stride_in;
  float4* spt = spt_all + (size_t)seg * stride_in;  // the points in slot order: later phases stream it, no gathers
  int* ev = ev_all + (size_t)seg * stride_in;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (n == 0) {
    if (tid == 0) out_count[seg] = 0;
    return;
  }
  for (int k = tid; k < VW * NSLOT; k += VT) hist[k] = 0;
  __syncthreads();

  // ---- 1a. warp-private histograms over contiguous slices
  const int L = (n + VW - 1) / VW;
  const int w_lo = min(wid * L, n), w_hi = min(w_lo + L, n);
  for (int i = w_lo + lane; i < w_hi; i += 128) {  // four loads in flight per lane
    float4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + 32 * u < w_hi) q[u] = P[i + 32 * u];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + 32 * u < w_hi) atomicAdd(&hist[wid * NSLOT + vox_key(q[u], inv).slot], 1);
  }
  __syncthreads();
  // ---- 1b. slot totals -> slot offsets -> per-(warp,slot) write cursors
  if (tid < NSLOT) {
    int t = 0;
    for (int w = 0; w < VW; ++w) t += hist[w * NSLOT + tid];
    slot_start[tid] = t;
    slot_rank[tid] = t > 0 ? 1 : 0;
  }
  __syncthreads();
  block_excl_scan(slot_start, NSLOT, s_warp, s_misc);
  const int n_occ = block_excl_scan(slot_rank, NSLOT, s_warp, s_misc);
  if (tid < NSLOT) {
    int run = slot_start[tid];
    for (int w = 0; w < VW; ++w) {
      int t = hist[w * NSLOT + tid];
      hist[w * NSLOT + tid] = run;
      run += t;
    }
    slot_end[tid] = run;
  }
  __syncthreads();
  // ---- 1c. stable scatter: each warp walks its slice in order, 32 points at a time
  float4 nxt = (w_lo + lane < w_hi) ? P[w_lo + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = w_lo; base < w_hi; base += 32) {
    const int i = base + lane;
    const bool act = i < w_hi;
    const unsigned amask = __ballot_sync(0xffffffffu, act);
    const float4 pt = nxt;
    if (i + 32 < w_hi) nxt = P[i + 32];  // the next trip's point is on its way while this one is ranked
    if (act) {
      const unsigned s = vox_key(pt, inv).slot;
      const unsigned peers = __match_any_sync(amask, s);
      const int leader = __ffs(peers) - 1;
      const int rank = __popc(peers & ((1u << lane) - 1u));
      int old = 0;
      if (lane == leader) {
        old = hist[wid * NSLOT + s];
        hist[wid * NSLOT + s] = old + __popc(peers);
      }
      old = __shfl_sync(peers, old, leader);
      sorted[old + rank] = i;
      spt[old + rank] = pt;
    }
    __syncwarp();
  }
  __syncthreads();  // (block-scope barrier also orders the global writes to sorted[] for this CTA)

  // ---- 2. run heads; a head that is not the first point of its slot evicts the previous run (four positions per trip,
  //         every load of the trip issued before the first use)
  for (int p0 = tid; p0 < n; p0 += 4 * VT) {
    float4 a[4], b[4];
    int si[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u * VT;
      if (p < n) {
        a[u] = spt[p];
        b[u] = spt[p > 0 ? p - 1 : 0];
        si[u] = sorted[p];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u * VT;
      if (p >= n) break;
      const VoxKey k = vox_key(a[u], inv);
      int evict = 0;
      if (p != slot_start[k.slot]) {
        const VoxKey kp = vox_key(b[u], inv);
        evict = (kp.ix != k.ix || kp.iy != k.iy || kp.iz != k.iz) ? 1 : 0;
      }
      ev[si[u]] = evict;
    }
  }
  __syncthreads();
  const int n_evict = block_excl_scan(ev, n, s_warp, s_misc);  // ev[i] := #evicting points before input index i
  const int n_out = n_evict + n_occ;
  if (n_out > stride_out) {
    if (tid == 0) {
      atomicExch(overflow, 1);
      out_count[seg] = 0;
    }
    return;
  }
  if (tid == 0) out_count[seg] = n_out;

  // ---- 3. one thread per run: sequential float sums in input order, then the flush arithmetic of PCL
  float4* O = out + (size_t)seg * stride_out;
  for (int p = tid; p < n; p += VT) {
    const VoxKey k = vox_key(spt[p], inv);
    const int s_lo = slot_start[k.slot], s_hi = slot_end[k.slot];
    if (p != s_lo) {
      const VoxKey kp = vox_key(spt[p - 1], inv);
      if (kp.ix == k.ix && kp.iy == k.iy && kp.iz == k.iz) continue;  // not a run head
    }
    float sx = 0.f, sy = 0.f, sz = 0.f, sr = 0.f, sg = 0.f, sb = 0.f;
    int cnt = 0, q = p, evictor = -1;
    float4 pt = spt[q];
    while (true) {
      const unsigned c = __float_as_uint(pt.w);
      sx = fadd(sx, pt.x);
      sy = fadd(sy, pt.y);
      sz = fadd(sz, pt.z);
      sr = fadd(sr, (float)((c >> 16) & 255u));
      sg = fadd(sg, (float)((c >> 8) & 255u));
      sb = fadd(sb, (float)(c & 255u));
      ++cnt;
      ++q;
      if (q >= s_hi) break;
      pt = spt[q];
      const VoxKey kn = vox_key(pt, inv);
      if (kn.ix != k.ix || kn.iy != k.iy || kn.iz != k.iz) {
        evictor = sorted[q];
        break;
      }
    }
    const float fc = (float)cnt;
    const int r = (int)__fdiv_rn(sr, fc), g = (int)__fdiv_rn(sg, fc), b = (int)__fdiv_rn(sb, fc);
    const unsigned rgb = ((unsigned)r << 16) | ((unsigned)g << 8) | (unsigned)b;  // alpha byte 0, as PCL packs it
    const int pos = evictor >= 0 ? ev[evictor] : n_evict + slot_rank[k.slot];
    O[pos] = make_float4(__fdiv_rn(sx, fc), __fdiv_rn(sy, fc), __fdiv_rn(sz, fc), __uint_as_float(rgb));
  }
}

__global__ void k_copy_strided(const float4* __restrict__ src, int stride_src, const int* __restrict__ count,
                               float4* __restrict__ dst, int stride_dst) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    dst[(size_t)seg * stride_dst + i] = src[(size_t)seg * stride_src + i];
}

__global__ void k_voxel_keys(const float4* __restrict__ pts, const int* __restrict__ count, const int* __restrict__ offsets,
                             int stride, float3 inv, int* __restrict__ ijk, int* __restrict__ slot) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  const long long base = offsets[seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    VoxKey k = vox_key(pts[(size_t)seg * stride + i], inv);
    ijk[3 * (base + i) + 0] = k.ix;
    ijk[3 * (base + i) + 1] = k.iy;
    ijk[3 * (base + i) + 2] = k.iz;
    slot[base + i] = (int)k.slot;
  }
}

constexpr size_t VOX_SMEM = (size_t)(VW * NSLOT + 3 * NSLOT + 32 + 4) * sizeof(int);

}  // namespace

int voxel_approx_device(rspcl_ctx* ctx, const rspcl_cloud* in, const float leaf[3], rspcl_cloud* out) {
  if (out->n_seg != in->n_seg) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "voxel_approx: n_seg mismatch");
  const int S = in->n_seg;
  const float3 inv = make_float3(1.0f / leaf[0], 1.0f / leaf[1], 1.0f / leaf[2]);  // PCL: inverse_leaf_size_ in float
  // per device and cheap: set on every call rather than caching a per-process flag (several devices per process)
  CU(ctx, cudaFuncSetAttribute(k_approx_voxel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VOX_SMEM));
  const size_t tot = (size_t)S * (in->stride ? in->stride : 1);
  Scratch scr(ctx);  // released on every exit path
  int *sorted = nullptr, *ev = nullptr, *d_over = nullptr;
  float4* tmp = nullptr;
  float4* spt = nullptr;
  CU(ctx, scr.alloc(&sorted, tot));
  CU(ctx, scr.alloc(&spt, tot));
  CU(ctx, scr.alloc(&ev, tot));
  CU(ctx, scr.alloc(&d_over, 1));
  CU(ctx, cudaMemsetAsync(d_over, 0, sizeof(int), ctx->stream));
  const bool alias = (in == out);
  float4* obuf = out->pts;
  int ostride = out->stride;
  if (alias) {
    CU(ctx, scr.alloc(&tmp, tot));
    obuf = tmp;
    ostride = in->stride;
  }
  ProfScope prof(ctx, "k_approx_voxel", (double)S * in->max_count_hint);
  k_approx_voxel<<<S, VT, VOX_SMEM, ctx->stream>>>(in->pts, in->count, in->stride, inv, sorted, spt, ev, obuf, out->count, ostride,
                                                   d_over);
  LAUNCH_CHECK(ctx);
  if (alias) {
    dim3 grid(blocks_per_seg(ctx, S, in->max_count_hint, 256), S);
    k_copy_strided<<<grid, 256, 0, ctx->stream>>>(tmp, ostride, out->count, out->pts, out->stride);
    LAUNCH_CHECK(ctx);
  }
  int over = 0;
  if (!alias && out->stride < in->max_count_hint) {
    CU(ctx, small_d2h(ctx, &over, d_over, sizeof(int)));
    CU(ctx, ctx_sync(ctx));
  }
  if (over) RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "voxel_approx: output stride %d too small", out->stride);
  if (!alias) {
    out->max_count_hint = in->max_count_hint < out->stride ? in->max_count_hint : out->stride;
  }
  scr.ok();
  out->width = out->height = 0;  // downsampling breaks the organized structure (PCL sets height = 1)
  return RSPCL_OK;
}

extern "C" int rspcl_voxel_approx(rspcl_ctx* ctx, const rspcl_cloud* in, const float leaf[3], rspcl_cloud* out) {
  if (!ctx || !in || !leaf || !out) return RSPCL_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  return voxel_approx_device(ctx, in, leaf, out);
}

extern "C" int rspcl_voxel_keys(rspcl_ctx* ctx, const rspcl_cloud* in, const float leaf[3], int32_t* host_ijk,
                                int32_t* host_slot) {
  if (!ctx || !in || !leaf || !host_ijk || !host_slot) return RSPCL_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  std::vector<int> cnt(in->n_seg), off(in->n_seg + 1);
  int rc = rspcl_cloud_counts(ctx, in, cnt.data());
  if (rc) return rc;
  long long total = 0;
  int maxc = 0;
  for (int s = 0; s < in->n_seg; ++s) {
    off[s] = (int)total;
    total += cnt[s];
    if (cnt[s] > maxc) maxc = cnt[s];
  }
  off[in->n_seg] = (int)total;
  if (!total) return RSPCL_OK;
  int *d_off = nullptr, *d_ijk = nullptr, *d_slot = nullptr;
  CU(ctx, scratch_alloc(ctx, &d_off, (size_t)in->n_seg + 1));
  CU(ctx, scratch_alloc(ctx, &d_ijk, (size_t)total * 3));
  CU(ctx, scratch_alloc(ctx, &d_slot, (size_t)total));
  CU(ctx, small_h2d(ctx, d_off, off.data(), (in->n_seg + 1) * sizeof(int)));
  const float3 inv = make_float3(1.0f / leaf[0], 1.0f / leaf[1], 1.0f / leaf[2]);
  dim3 grid(blocks_per_seg(ctx, in->n_seg, maxc, 256), in->n_seg);
  k_voxel_keys<<<grid, 256, 0, ctx->stream>>>(in->pts, in->count, d_off, in->stride, inv, d_ijk, d_slot);
  LAUNCH_CHECK(ctx);
  CU(ctx, small_d2h(ctx, host_ijk, d_ijk, (size_t)total * 3 * sizeof(int)));
  CU(ctx, small_d2h(ctx, host_slot, d_slot, (size_t)total * sizeof(int)));
  CU(ctx, ctx_sync(ctx));
  scratch_free(ctx, d_off);
  scratch_free(ctx, d_ijk);
  scratch_free(ctx, d_slot);
  return RSPCL_OK;
}
