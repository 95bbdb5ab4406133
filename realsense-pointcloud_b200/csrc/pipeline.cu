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
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w_h; i += gridDim.x * blockDim.x) {
    float4 p = F[i];
    if (finite3(p.x, p.y, p.z)) {
      float3 q = xform_point(A, p.x, p.y, p.z);
      q = xform_point(B, q.x, q.y, q.z);
      p.x = q.x;
      p.y = q.y;
      p.z = q.z;
    }
    out[(size_t)seg * out_stride + i] = p;
  }
}

}  // namespace

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
