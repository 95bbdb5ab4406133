// pipeline.cu -- the reference's per-frame registration body for a batch of independent frame pairs.
//
// One call performs, for every pair (src frame k, tgt frame k-1): extract_features on each distinct frame
// (types.hpp:34-38 -> edge_extractor.hpp:7-39), ApproximateVoxelGrid on both edge clouds (icp:59-60,75-76),
// the coarse stage from the initial guess (ICP icp:95,104 or NDT ndt:83,92), the fine ICP from identity
// (icp:108-111 / ndt:96-99) and, if asked, transformPointCloud of the full source frame by both transforms
// (icp:116-117 / ndt:104-105).  The target is the previous frame's edge cloud instead of the reference's
// accumulating target (the pairwise formulation is what makes pairs independent and shardable);
// the sequential accumulate-into-global schemes are built on the same primitives by the host facade.
#include "common.cuh"
#include <float.h>

int icp_align_device(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt, const rspcl_icp_params* prm,
                     const float* d_guess, rspcl_icp_result* h_results, rspcl_cloud* aligned, const IcpAlignOpts& o);
int ndt_align_device(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt, const rspcl_ndt_params* prm,
                     const float* d_guess, rspcl_ndt_result* h_results, rspcl_cloud* aligned);
int voxel_approx_device(rspcl_ctx* ctx, const rspcl_cloud* in, const float leaf[3], rspcl_cloud* out);
int refresh_count_hint(rspcl_ctx* ctx, const rspcl_cloud* c, std::vector<int>* counts_out);

namespace {

__global__ void k_gather_segments(const float4* __restrict__ in, const int* __restrict__ in_count, int in_stride,
                                  const int* __restrict__ idx, float4* __restrict__ out, int* __restrict__ out_count,
                                  int out_stride) {
  const int seg = blockIdx.y;
  const int s = idx[seg];
  const int n = min(in_count[s], out_stride);
  if (blockIdx.x == 0 && threadIdx.x == 0) out_count[seg] = n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[(size_t)seg * out_stride + i] = in[(size_t)s * in_stride + i];
}

// K8 fused: full source frame moved by T_coarse then T_fine (two float transforms back to back, bit-identical to
// the reference's two transformPointCloud calls, one read + one write per point)
__global__ void k_transform2_gather(const float4* __restrict__ frames, int w_h, int in_stride, const int* __restrict__ src_idx,
                                    const float* __restrict__ Tc, const float* __restrict__ Tf,
                                    const int* __restrict__ accept, float4* __restrict__ out, int* __restrict__ out_count,
                                    int out_stride) {
  __shared__ float A[16], B[16];
  const int seg = blockIdx.y;
  if (threadIdx.x < 16) {
    A[threadIdx.x] = Tc[seg * 16 + threadIdx.x];
    B[threadIdx.x] = Tf[seg * 16 + threadIdx.x];
  }
  __syncthreads();
  const int ok = accept[seg];
  if (blockIdx.x == 0 && threadIdx.x == 0) out_count[seg] = ok ? w_h : 0;
  if (!ok) return;
  const float4* F = frames + (size_t)src_idx[seg] * in_stride;
  float4* O = out + (size_t)seg * out_stride;
  const int G = gridDim.x * blockDim.x;
  // four points per trip, all four loads issued before the first use (HBM-bound: bytes in flight are what counts)
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < w_h; i0 += 4 * G) {
    float4 p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u * G < w_h) p[u] = __ldcs(&F[i0 + u * G]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + u * G >= w_h) break;
      float4 v = p[u];
      if (finite3(v.x, v.y, v.z)) {
        float3 q = xform_point(A, v.x, v.y, v.z);
        q = xform_point(B, q.x, q.y, q.z);
        v.x = q.x;
        v.y = q.y;
        v.z = q.z;
      }
      __stcs(&O[i0 + u * G], v);
    }
  }
}

// one cloud segment moved by two transforms in a row (exactly two pcl::transformPointCloud calls) into out[0 .. n)
__global__ void k_transform2_seg(const float4* __restrict__ in, const int* __restrict__ in_count, int n_fixed,
                                 const float* __restrict__ Ta, const float* __restrict__ Tb, float4* __restrict__ out) {
  __shared__ float A[16], B[16];
  if (threadIdx.x < 16) {
    A[threadIdx.x] = Ta[threadIdx.x];
    B[threadIdx.x] = Tb[threadIdx.x];
  }
  __syncthreads();
  const int n = in_count ? *in_count : n_fixed;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = in[i];
    if (finite3(p.x, p.y, p.z)) {
      float3 q = xform_point(A, p.x, p.y, p.z);
      q = xform_point(B, q.x, q.y, q.z);
      p.x = q.x;
      p.y = q.y;
      p.z = q.z;
    }
    out[i] = p;
  }
}

__global__ void k_copy_points(const float4* __restrict__ in, int n, float4* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i];
}

}  // namespace

extern "C" int rspcl_register_sequence(rspcl_ctx* ctx, const rspcl_cloud* frames, int coarse_kind, const rspcl_icp_params* icp,
                                       const rspcl_ndt_params* ndt, const float leaf[3], float t_low, float t_high,
                                       const float* guess, rspcl_pair_result* results, rspcl_cloud* out_global,
                                       rspcl_cloud* out_target) {
  if (!ctx || !frames || !icp || !leaf || !guess || !results || !out_global) return RSPCL_ERR_ARG;
  if (coarse_kind == RSPCL_COARSE_NDT && !ndt) return RSPCL_ERR_ARG;
  if (frames->height <= 0) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "register_sequence: frames are not organized");
  const int F = frames->n_seg, npx = frames->width * frames->height;
  if (out_global->n_seg != 1 || (long long)out_global->stride < (long long)F * npx)
    RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "register_sequence: out_global needs 1 segment of stride >= %lld", (long long)F * npx);
  CU(ctx, cudaSetDevice(ctx->device));

  // phase 1 (types.hpp:34-38) + the voxel filter of every edge cloud (icp:59-60,75-76)
  TmpCloud E(ctx);
  int rc = E.init(F, npx);
  if (rc) RSPCL_FAIL(ctx, rc, "register_sequence: scratch allocation failed");
  rc = rspcl_edge_extract(ctx, frames, t_low, t_high, &E.c, nullptr);
  if (rc) return rc;
  rc = voxel_approx_device(ctx, &E.c, leaf, &E.c);
  if (rc) return rc;
  std::vector<int> vcnt;
  rc = refresh_count_hint(ctx, &E.c, &vcnt);
  if (rc) return rc;
  long long cap = 0;
  for (int v : vcnt) cap += v;
  if (cap < 1) cap = 1;
  if (out_target && (out_target->n_seg != 1 || (long long)out_target->stride < cap))
    RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "register_sequence: out_target needs 1 segment of stride >= %lld", cap);

  Scratch scr(ctx);
  float4* tbuf = nullptr;   // the accumulating edge target occupies tbuf[cap - n_t .. cap)
  int* d_cnt = nullptr;     // [0] target count, [1] coarse-aligned source count (NDT path)
  float* d_T = nullptr;     // guess | T_coarse | T_fine of the current frame
  CU(ctx, scr.alloc(&tbuf, (size_t)cap));
  CU(ctx, scr.alloc(&d_cnt, 2));
  CU(ctx, scr.alloc(&d_T, 48));
  auto nblk = [&](long long n) {
    long long b = (n + 255) / 256, m = 8LL * ctx->sm_count;
    return (unsigned)(b < 1 ? 1 : (b > m ? m : b));
  };
  long long n_t = vcnt[0], n_g = npx;
  k_copy_points<<<nblk(n_t), 256, 0, ctx->stream>>>(E.c.pts, (int)n_t, tbuf + (cap - n_t));                 // target = clouds[0].first
  LAUNCH_CHECK(ctx);
  k_copy_points<<<nblk(npx), 256, 0, ctx->stream>>>(frames->pts, npx, out_global->pts);                      // icp:57
  LAUNCH_CHECK(ctx);
  memset(&results[0], 0, sizeof(rspcl_pair_result));
  for (int i = 0; i < 16; ++i) results[0].T_coarse[i] = results[0].T_fine[i] = (i % 5 == 0) ? 1.f : 0.f;
  results[0].converged = 1;
  results[0].n_src = results[0].n_tgt = vcnt[0];

  TmpCloud Ac(ctx);
  if (coarse_kind == RSPCL_COARSE_NDT && Ac.init(1, E.c.max_count_hint > 0 ? E.c.max_count_hint : 1))
    RSPCL_FAIL(ctx, RSPCL_ERR_CUDA, "register_sequence: scratch allocation failed");
  double prev_coarse = DBL_MAX, prev_fine = DBL_MAX;  // one PCL object per stage lives across the frames (icp:35,41)
  for (int k = 1; k < F; ++k) {
    const int m = vcnt[k];
    rspcl_cloud src, tgt;  // single-segment views
    src.pts = E.c.pts + (size_t)k * E.c.stride;
    src.count = E.c.count + k;
    src.n_seg = 1;
    src.stride = E.c.stride;
    src.max_count_hint = m;
    const int nt_i = (int)n_t;
    CU(ctx, small_h2d(ctx, d_cnt, &nt_i, sizeof(int)));
    tgt.pts = tbuf + (cap - n_t);
    tgt.count = d_cnt;
    tgt.n_seg = 1;
    tgt.stride = (int)n_t;
    tgt.max_count_hint = (int)n_t;
    CU(ctx, small_h2d(ctx, d_T, guess + (size_t)k * 16, 16 * sizeof(float)));
    rspcl_pair_result& R = results[k];
    memset(&R, 0, sizeof(R));
    rspcl_icp_result fine;
    fine.prev_mse = prev_fine;
    const float4* fine_in = src.pts;  // what the two transforms below are applied to
    const int* fine_in_count = src.count;
    if (coarse_kind == RSPCL_COARSE_NDT) {
      rspcl_ndt_result nr;
      rc = ndt_align_device(ctx, &src, &tgt, ndt, d_T, &nr, &Ac.c);
      if (rc) return rc;
      memcpy(R.T_coarse, nr.T, 64);
      R.coarse_iterations = nr.iterations;
      IcpAlignOpts o;
      o.h_src_counts = &m;
      rc = icp_align_device(ctx, &Ac.c, &tgt, icp, nullptr, &fine, nullptr, o);
      if (rc) return rc;
    } else {
      rspcl_icp_result cr;
      cr.prev_mse = prev_coarse;
      IcpAlignOpts o;
      o.h_results2 = &fine;
      o.h_src_counts = &m;
      rc = icp_align_device(ctx, &src, &tgt, icp, d_T, &cr, nullptr, o);
      if (rc) return rc;
      memcpy(R.T_coarse, cr.T, 64);
      R.coarse_iterations = cr.iterations;
      prev_coarse = cr.prev_mse;
    }
    prev_fine = fine.prev_mse;
    memcpy(R.T_fine, fine.T, 64);
    R.converged = fine.converged;
    R.fine_iterations = fine.iterations;
    R.n_corr = fine.n_corr;
    R.mse = fine.mse;
    R.n_src = m;
    R.n_tgt = (int)n_t;
    if (!fine.converged) continue;  // failed frames are skipped silently (icp:113-123)
    CU(ctx, small_h2d(ctx, d_T + 16, R.T_coarse, 16 * sizeof(float)));
    CU(ctx, small_h2d(ctx, d_T + 32, R.T_fine, 16 * sizeof(float)));
    // *target = *icp_aligned + *target (icp:119): the fine-aligned edges go in FRONT of the target
    if (m > 0) {
      k_transform2_seg<<<nblk(m), 256, 0, ctx->stream>>>(fine_in, fine_in_count, m, d_T + 16, d_T + 32, tbuf + (cap - n_t - m));
      LAUNCH_CHECK(ctx);
      n_t += m;
    }
    // *global += transformed (icp:116-117,120)
    {
      ProfScope prof(ctx, "k_transform2", (double)npx);
      k_transform2_seg<<<nblk(npx), 256, 0, ctx->stream>>>(frames->pts + (size_t)k * frames->stride, nullptr, npx, d_T + 16, d_T + 32,
                                                          out_global->pts + n_g);
      LAUNCH_CHECK(ctx);
    }
    n_g += npx;
  }
  const int ng_i = (int)n_g, nt_i = (int)n_t;
  CU(ctx, small_h2d(ctx, out_global->count, &ng_i, sizeof(int)));
  out_global->max_count_hint = ng_i;
  out_global->width = out_global->height = 0;
  invalidate_gray(out_global);
  if (out_target) {
    k_copy_points<<<nblk(n_t), 256, 0, ctx->stream>>>(tbuf + (cap - n_t), nt_i, out_target->pts);
    LAUNCH_CHECK(ctx);
    CU(ctx, small_h2d(ctx, out_target->count, &nt_i, sizeof(int)));
    out_target->max_count_hint = nt_i;
    out_target->width = out_target->height = 0;
    invalidate_gray(out_target);
  }
  CU(ctx, ctx_sync(ctx));
  scr.ok();
  return RSPCL_OK;
}

extern "C" int rspcl_register_pairs(rspcl_ctx* ctx, const rspcl_cloud* frames, const int32_t* src_idx, const int32_t* tgt_idx,
                                    int n_pairs, int coarse_kind, const rspcl_icp_params* icp, const rspcl_ndt_params* ndt,
                                    const float leaf[3], float t_low, float t_high, const float* guess,
                                    rspcl_pair_result* results, rspcl_cloud* out_transformed) {
  if (!ctx || !frames || !src_idx || !tgt_idx || n_pairs <= 0 || !icp || !leaf || !results) return RSPCL_ERR_ARG;
  if (coarse_kind == RSPCL_COARSE_NDT && !ndt) return RSPCL_ERR_ARG;
  if (frames->height <= 0) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "register_pairs: frames are not organized");
  const int F = frames->n_seg, npx = frames->width * frames->height;
  for (int i = 0; i < n_pairs; ++i)
    if (src_idx[i] < 0 || src_idx[i] >= F || tgt_idx[i] < 0 || tgt_idx[i] >= F)
      RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "register_pairs: pair %d references a frame outside [0,%d)", i, F);
  if (out_transformed && (out_transformed->n_seg != n_pairs || out_transformed->stride < npx))
    RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "register_pairs: out_transformed needs %d segments of stride >= %d", n_pairs, npx);
  CU(ctx, cudaSetDevice(ctx->device));

  // 1-2. edges of every frame, voxel filter in place
  TmpCloud E(ctx);
  int rc = E.init(F, npx);
  if (rc) RSPCL_FAIL(ctx, rc, "register_pairs: scratch allocation failed");
  rc = rspcl_edge_extract(ctx, frames, t_low, t_high, &E.c, nullptr);
  if (rc) return rc;
  rc = voxel_approx_device(ctx, &E.c, leaf, &E.c);
  if (rc) return rc;
  std::vector<int> vcnt;
  rc = refresh_count_hint(ctx, &E.c, &vcnt);  // the one host round-trip before the registration stages
  if (rc) return rc;
  const int pstride = E.c.max_count_hint > 0 ? E.c.max_count_hint : 1;

  // 3. pair-major batches
  Scratch scr(ctx);
  int *d_src = nullptr, *d_tgt = nullptr;
  CU(ctx, scr.alloc(&d_src, (size_t)n_pairs));
  CU(ctx, scr.alloc(&d_tgt, (size_t)n_pairs));
  CU(ctx, small_h2d(ctx, d_src, src_idx, n_pairs * sizeof(int)));
  CU(ctx, small_h2d(ctx, d_tgt, tgt_idx, n_pairs * sizeof(int)));
  TmpCloud Sc(ctx), Tc(ctx);
  if (Sc.init(n_pairs, pstride) || Tc.init(n_pairs, pstride))
    RSPCL_FAIL(ctx, RSPCL_ERR_CUDA, "register_pairs: scratch allocation failed");
  dim3 gg(blocks_per_seg(ctx, n_pairs, pstride, 256), n_pairs);
  k_gather_segments<<<gg, 256, 0, ctx->stream>>>(E.c.pts, E.c.count, E.c.stride, d_src, Sc.c.pts, Sc.c.count, pstride);
  LAUNCH_CHECK(ctx);
  k_gather_segments<<<gg, 256, 0, ctx->stream>>>(E.c.pts, E.c.count, E.c.stride, d_tgt, Tc.c.pts, Tc.c.count, pstride);
  LAUNCH_CHECK(ctx);
  Sc.c.max_count_hint = Tc.c.max_count_hint = pstride;
  std::vector<int> scnt(n_pairs);
  for (int i = 0; i < n_pairs; ++i) scnt[i] = vcnt[src_idx[i]] < pstride ? vcnt[src_idx[i]] : pstride;

  // 4. coarse stage + 5. fine ICP from identity on the coarse-aligned source
  float* d_guess = nullptr;
  if (guess) {
    CU(ctx, scr.alloc(&d_guess, (size_t)n_pairs * 16));
    CU(ctx, small_h2d(ctx, d_guess, guess, (size_t)n_pairs * 16 * sizeof(float)));
  }
  std::vector<float> hT((size_t)n_pairs * 32);
  std::vector<rspcl_icp_result> fine(n_pairs);
  for (auto& r : fine) r.prev_mse = DBL_MAX;
  if (coarse_kind == RSPCL_COARSE_NDT) {
    TmpCloud Ac(ctx);
    if (Ac.init(n_pairs, pstride)) RSPCL_FAIL(ctx, RSPCL_ERR_CUDA, "register_pairs: scratch allocation failed");
    std::vector<rspcl_ndt_result> nr(n_pairs);
    rc = ndt_align_device(ctx, &Sc.c, &Tc.c, ndt, d_guess, nr.data(), &Ac.c);
    if (rc) return rc;
    for (int i = 0; i < n_pairs; ++i) {
      memcpy(results[i].T_coarse, nr[i].T, 64);
      results[i].coarse_iterations = nr[i].iterations;
    }
    IcpAlignOpts o;
    o.h_src_counts = scnt.data();
    rc = icp_align_device(ctx, &Ac.c, &Tc.c, icp, nullptr, fine.data(), nullptr, o);
    if (rc) return rc;
  } else {
    // coarse and fine ICP of every pair in ONE launch of the persistent kernel: the fine align runs on the same pairs
    // against the same targets, its source being the original source moved by the coarse result, so it keeps the
    // target replica and the certified nearest-neighbour cache of the coarse align
    std::vector<rspcl_icp_result> cr(n_pairs);
    for (auto& r : cr) r.prev_mse = DBL_MAX;
    IcpAlignOpts o;
    o.h_results2 = fine.data();
    o.h_src_counts = scnt.data();
    rc = icp_align_device(ctx, &Sc.c, &Tc.c, icp, d_guess, cr.data(), nullptr, o);
    if (rc) return rc;
    for (int i = 0; i < n_pairs; ++i) {
      memcpy(results[i].T_coarse, cr[i].T, 64);
      results[i].coarse_iterations = cr[i].iterations;
    }
  }
  std::vector<int> accept(n_pairs);
  for (int i = 0; i < n_pairs; ++i) {
    memcpy(results[i].T_fine, fine[i].T, 64);
    results[i].converged = fine[i].converged;
    results[i].fine_iterations = fine[i].iterations;
    results[i].n_corr = fine[i].n_corr;
    results[i].mse = fine[i].mse;
    results[i].n_src = vcnt[src_idx[i]];
    results[i].n_tgt = vcnt[tgt_idx[i]];
    accept[i] = fine[i].converged;
    memcpy(&hT[(size_t)i * 16], results[i].T_coarse, 64);
    memcpy(&hT[(size_t)(n_pairs + i) * 16], results[i].T_fine, 64);
  }
  // 6. full-cloud transform of the accepted frames (stream-ordered: the call returns without waiting for it)
  if (out_transformed) {
    float* d_T = nullptr;
    int* d_acc = nullptr;
    CU(ctx, scr.alloc(&d_T, (size_t)n_pairs * 32));
    CU(ctx, scr.alloc(&d_acc, (size_t)n_pairs));
    CU(ctx, small_h2d(ctx, d_T, hT.data(), hT.size() * sizeof(float)));
    CU(ctx, small_h2d(ctx, d_acc, accept.data(), n_pairs * sizeof(int)));
    dim3 gt(blocks_per_seg(ctx, n_pairs, npx, 256), n_pairs);
    ProfScope prof(ctx, "k_transform2", (double)n_pairs * npx);
    k_transform2_gather<<<gt, 256, 0, ctx->stream>>>(frames->pts, npx, frames->stride, d_src, d_T, d_T + (size_t)n_pairs * 16,
                                                     d_acc, out_transformed->pts, out_transformed->count,
                                                     out_transformed->stride);
    LAUNCH_CHECK(ctx);
    out_transformed->width = frames->width;
    out_transformed->height = frames->height;
    out_transformed->max_count_hint = npx;
    invalidate_gray(out_transformed);
  }
  scr.ok();
  return RSPCL_OK;
}
