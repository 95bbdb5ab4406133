// grid.cu -- voxel-hash grid build (K4), brute-force exact NN, getFitnessScore (K9).
//
// Reference call sites: kd-tree construction inside every align() (icp:79,95,109,111; ndt:97-99; incr:58-59) and
// pcl::Registration::getFitnessScore (PCL surface kept by the north-star; never called by the reference).
#include "grid.cuh"
#include <float.h>

namespace {

__global__ void k_grid_init(GridSlot* __restrict__ slots, int* __restrict__ cnt, unsigned cap) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
    GridSlot e;
    e.key = GRID_EMPTY;
    e.start = 0;
    e.cnt = 0;
    slots[i] = e;
    cnt[i] = 0;
  }
}

__global__ void k_grid_finalize(GridSlot* __restrict__ slots, const int* __restrict__ cnt, const int* __restrict__ start,
                                unsigned cap) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
    const int c = cnt[i];
    if (c) {
      slots[i].start = start[i];
      slots[i].cnt = c;
    }
  }
}

// insert each target point's cell; remember (slot, rank-in-cell).  Non-finite points are left out of the grid
// (KdTreeFLANN::convertCloudToArray skips them); points outside the key range raise the range flag.
__global__ void k_grid_insert(const float4* __restrict__ pts, const int* __restrict__ count, int stride, DevGrid g,
                              const int* __restrict__ seg_off, int* __restrict__ range_flag) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  const int base = seg_off[seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[(size_t)seg * stride + i];
    int slot = -1, rank = 0;
    if (finite3(p.x, p.y, p.z)) {
      const int ix = grid_cell(p.x, g.inv_cs), iy = grid_cell(p.y, g.inv_cs), iz = grid_cell(p.z, g.inv_cs);
      if (!grid_in_range(ix, iy, iz)) {
        atomicExch(range_flag, 1);
      } else {
        const unsigned long long key = grid_key(seg, ix, iy, iz);
        unsigned s = grid_hash4(seg, ix, iy, iz) & g.cap_mask;
        while (true) {
          unsigned long long prev = atomicCAS(&g.slots[s].key, GRID_EMPTY, key);
          if (prev == GRID_EMPTY || prev == key) break;
          s = (s + 1) & g.cap_mask;
        }
        slot = (int)s;
        rank = atomicAdd(&g.cnt[s], 1);
      }
    }
    g.slot_of[base + i] = slot;
    g.rank_of[base + i] = rank;
  }
}

__global__ void k_grid_scatter(const float4* __restrict__ pts, const int* __restrict__ count, int stride, DevGrid g,
                               const int* __restrict__ seg_off) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  const int base = seg_off[seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = g.slot_of[base + i];
    if (s < 0) continue;
    float4 p = pts[(size_t)seg * stride + i];
    p.w = __int_as_float(i);
    g.sorted[g.start[s] + g.rank_of[base + i]] = p;
  }
}

__global__ void k_seg_offsets(const int* __restrict__ count, int n_seg, int stride, int* __restrict__ seg_off) {
  // strided positions are enough: scratch indexed by (seg * stride + i)
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += gridDim.x * blockDim.x) seg_off[s] = s * stride;
  (void)count;
}

constexpr int BT = 256;  // brute-force tile / block size

__global__ void __launch_bounds__(BT) k_nn_brute(const float4* __restrict__ q, const int* __restrict__ qcount, int qstride,
                                                 const float4* __restrict__ t, const int* __restrict__ tcount, int tstride,
                                                 int shared_target, int* __restrict__ out_idx, float* __restrict__ out_d2) {
  __shared__ float4 tile[BT];
  const int seg = blockIdx.y;
  const int nq = qcount[seg];
  const int tseg = shared_target ? 0 : seg;
  const int nt = tcount[tseg];
  const float4* T = t + (size_t)tseg * tstride;
  for (int qb = blockIdx.x * BT; qb < nq; qb += gridDim.x * BT) {  // uniform across the block
    const int i = qb + threadIdx.x;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nq) p = q[(size_t)seg * qstride + i];
    int best = -1;
    float bd = INFINITY;
    for (int tb = 0; tb < nt; tb += BT) {
      __syncthreads();
      if (tb + threadIdx.x < nt) tile[threadIdx.x] = T[tb + threadIdx.x];
      __syncthreads();
      const int m = min(BT, nt - tb);
#pragma unroll 4
      for (int k = 0; k < m; ++k) {
        const float4 c = tile[k];
        const float d = dist2_l2simple(p.x, p.y, p.z, c.x, c.y, c.z);
        if (d < bd) {  // ascending index + strict compare = lowest index wins ties; NaN never wins
          bd = d;
          best = tb + k;
        }
      }
    }
    if (i < nq) {
      out_idx[(size_t)seg * qstride + i] = best;
      out_d2[(size_t)seg * qstride + i] = bd;
    }
  }
}

// per-block partial {sum d2, count} over NN distances <= max_range; combined in block order by k_fitness_final
__global__ void __launch_bounds__(256) k_fitness_partial(const int* __restrict__ idx, const float* __restrict__ d2,
                                                         const int* __restrict__ qcount, int qstride, double max_range,
                                                         double* __restrict__ part) {
  __shared__ double s_sum[8], s_cnt[8];
  const int seg = blockIdx.y;
  const int n = qcount[seg];
  double sum = 0.0, cnt = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (idx[(size_t)seg * qstride + i] < 0) continue;
    const double d = (double)d2[(size_t)seg * qstride + i];
    if (d <= max_range) {
      sum += d;
      cnt += 1.0;
    }
  }
  sum = warp_sum(sum);
  cnt = warp_sum(cnt);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    s_sum[wid] = sum;
    s_cnt[wid] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; ++w) {
      a += s_sum[w];
      b += s_cnt[w];
    }
    part[((size_t)seg * gridDim.x + blockIdx.x) * 2 + 0] = a;
    part[((size_t)seg * gridDim.x + blockIdx.x) * 2 + 1] = b;
  }
}

__global__ void k_fitness_final(const double* __restrict__ part, int nblk, int n_seg, double* __restrict__ out) {
  const int seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= n_seg) return;
  double a = 0, b = 0;
  for (int k = 0; k < nblk; ++k) {
    a += part[((size_t)seg * nblk + k) * 2];
    b += part[((size_t)seg * nblk + k) * 2 + 1];
  }
  out[seg] = b > 0 ? a / b : DBL_MAX;
}

__global__ void k_pack_nn(const int* __restrict__ idx, const float* __restrict__ d2, const int* __restrict__ count,
                          const int* __restrict__ off, int stride, int* __restrict__ pidx, float* __restrict__ pd2) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    pidx[off[seg] + i] = idx[(size_t)seg * stride + i];
    pd2[off[seg] + i] = d2[(size_t)seg * stride + i];
  }
}

}  // namespace

int grid_build(rspcl_ctx* ctx, const rspcl_cloud* tgt, float cell_size, DevGrid* g, int* d_range_flag) {
  const int S = tgt->n_seg;
  const long long n_total = (long long)S * (tgt->stride ? tgt->stride : 1);  // scratch is strided like the cloud
  long long occupied_bound = (long long)S * (tgt->max_count_hint > 0 ? tgt->max_count_hint : 1);
  unsigned cap = 1024;
  while ((long long)cap < 2 * occupied_bound && cap < (1u << 30)) cap <<= 1;
  g->cap_mask = cap - 1;
  g->cs = cell_size;
  g->inv_cs = 1.0f / cell_size;
  g->n_total = n_total;
  CU(ctx, scratch_alloc(ctx, &g->slots, (size_t)cap));
  CU(ctx, scratch_alloc(ctx, &g->cnt, (size_t)cap));
  CU(ctx, scratch_alloc(ctx, &g->start, (size_t)cap));
  CU(ctx, scratch_alloc(ctx, &g->sorted, (size_t)n_total));
  CU(ctx, scratch_alloc(ctx, &g->slot_of, (size_t)n_total));
  CU(ctx, scratch_alloc(ctx, &g->rank_of, (size_t)n_total));
  int* seg_off = nullptr;
  CU(ctx, scratch_alloc(ctx, &seg_off, (size_t)S));
  k_grid_init<<<4 * ctx->sm_count, 256, 0, ctx->stream>>>(g->slots, g->cnt, cap);
  LAUNCH_CHECK(ctx);
  k_seg_offsets<<<div_up(S, 256), 256, 0, ctx->stream>>>(tgt->count, S, tgt->stride, seg_off);
  LAUNCH_CHECK(ctx);
  dim3 grid(blocks_per_seg(ctx, S, tgt->max_count_hint, 256), S);
  k_grid_insert<<<grid, 256, 0, ctx->stream>>>(tgt->pts, tgt->count, tgt->stride, *g, seg_off, d_range_flag);
  LAUNCH_CHECK(ctx);
  int rc = rspcl_exclusive_scan_i32(ctx, g->cnt, g->start, cap, nullptr);
  if (rc) return rc;
  k_grid_scatter<<<grid, 256, 0, ctx->stream>>>(tgt->pts, tgt->count, tgt->stride, *g, seg_off);
  LAUNCH_CHECK(ctx);
  k_grid_finalize<<<4 * ctx->sm_count, 256, 0, ctx->stream>>>(g->slots, g->cnt, g->start, cap);
  LAUNCH_CHECK(ctx);
  scratch_free(ctx, seg_off);
  scratch_free(ctx, g->slot_of);
  scratch_free(ctx, g->rank_of);
  scratch_free(ctx, g->cnt);
  scratch_free(ctx, g->start);
  g->slot_of = g->rank_of = g->cnt = g->start = nullptr;
  return RSPCL_OK;
}

void grid_free(rspcl_ctx* ctx, DevGrid* g) {
  scratch_free(ctx, g->slots);
  scratch_free(ctx, g->sorted);
  g->slots = nullptr;
  g->sorted = nullptr;
}

int nn_brute_device(rspcl_ctx* ctx, const rspcl_cloud* query, const rspcl_cloud* tgt, int* d_idx, float* d_d2) {
  const int shared_target = (tgt->n_seg == 1 && query->n_seg > 1) ? 1 : 0;
  if (!shared_target && tgt->n_seg != query->n_seg) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "nearest: n_seg mismatch");
  dim3 grid(blocks_per_seg(ctx, query->n_seg, query->max_count_hint, BT), query->n_seg);
  k_nn_brute<<<grid, BT, 0, ctx->stream>>>(query->pts, query->count, query->stride, tgt->pts, tgt->count, tgt->stride,
                                           shared_target, d_idx, d_d2);
  LAUNCH_CHECK(ctx);
  return RSPCL_OK;
}

extern "C" int rspcl_nearest(rspcl_ctx* ctx, const rspcl_cloud* query, const rspcl_cloud* tgt, int32_t* host_idx,
                             float* host_d2) {
  if (!ctx || !query || !tgt || !host_idx || !host_d2) return RSPCL_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  const size_t tot = (size_t)query->n_seg * (query->stride ? query->stride : 1);
  int* d_idx = nullptr;
  float* d_d2 = nullptr;
  CU(ctx, scratch_alloc(ctx, &d_idx, tot));
  CU(ctx, scratch_alloc(ctx, &d_d2, tot));
  int rc = nn_brute_device(ctx, query, tgt, d_idx, d_d2);
  if (rc) return rc;
  std::vector<int> cnt(query->n_seg), off(query->n_seg + 1);
  rc = rspcl_cloud_counts(ctx, query, cnt.data());
  if (rc) return rc;
  long long total = 0;
  int maxc = 0;
  for (int s = 0; s < query->n_seg; ++s) {
    off[s] = (int)total;
    total += cnt[s];
    if (cnt[s] > maxc) maxc = cnt[s];
  }
  if (total) {
    int *d_off = nullptr, *p_idx = nullptr;
    float* p_d2 = nullptr;
    CU(ctx, scratch_alloc(ctx, &d_off, (size_t)query->n_seg));
    CU(ctx, scratch_alloc(ctx, &p_idx, (size_t)total));
    CU(ctx, scratch_alloc(ctx, &p_d2, (size_t)total));
    CU(ctx, small_h2d(ctx, d_off, off.data(), query->n_seg * sizeof(int)));
    dim3 grid(blocks_per_seg(ctx, query->n_seg, maxc, 256), query->n_seg);
    k_pack_nn<<<grid, 256, 0, ctx->stream>>>(d_idx, d_d2, query->count, d_off, query->stride, p_idx, p_d2);
    LAUNCH_CHECK(ctx);
    CU(ctx, small_d2h(ctx, host_idx, p_idx, (size_t)total * sizeof(int)));
    CU(ctx, small_d2h(ctx, host_d2, p_d2, (size_t)total * sizeof(float)));
    CU(ctx, ctx_sync(ctx));
    scratch_free(ctx, d_off);
    scratch_free(ctx, p_idx);
    scratch_free(ctx, p_d2);
  }
  scratch_free(ctx, d_idx);
  scratch_free(ctx, d_d2);
  return RSPCL_OK;
}

extern "C" int rspcl_fitness(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt, double max_range,
                             double* fitness) {
  if (!ctx || !src || !tgt || !fitness) return RSPCL_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  const int S = src->n_seg;
  const size_t tot = (size_t)S * (src->stride ? src->stride : 1);
  int* d_idx = nullptr;
  float* d_d2 = nullptr;
  double *d_part = nullptr, *d_out = nullptr;
  CU(ctx, scratch_alloc(ctx, &d_idx, tot));
  CU(ctx, scratch_alloc(ctx, &d_d2, tot));
  int rc = nn_brute_device(ctx, src, tgt, d_idx, d_d2);
  if (rc) return rc;
  const int nblk = blocks_per_seg(ctx, S, src->max_count_hint, 256);
  CU(ctx, scratch_alloc(ctx, &d_part, (size_t)S * nblk * 2));
  CU(ctx, scratch_alloc(ctx, &d_out, (size_t)S));
  k_fitness_partial<<<dim3(nblk, S), 256, 0, ctx->stream>>>(d_idx, d_d2, src->count, src->stride, max_range, d_part);
  LAUNCH_CHECK(ctx);
  k_fitness_final<<<div_up(S, 128), 128, 0, ctx->stream>>>(d_part, nblk, S, d_out);
  LAUNCH_CHECK(ctx);
  CU(ctx, small_d2h(ctx, fitness, d_out, S * sizeof(double)));
  CU(ctx, ctx_sync(ctx));
  scratch_free(ctx, d_idx);
  scratch_free(ctx, d_d2);
  scratch_free(ctx, d_part);
  scratch_free(ctx, d_out);
  return RSPCL_OK;
}
