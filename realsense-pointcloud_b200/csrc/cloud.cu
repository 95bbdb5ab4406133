// cloud.cu -- context, device cloud batches, layout conversion, transformPointCloud, concatenation, crop, scan.
//
// Reference call sites replaced: types.hpp:8-10 (cloud type), icp:116-117 / ndt:104-105 / incr:63
// (pcl::transformPointCloud), icp:57,119-120 / ndt:55,107-108 / incr:64 (operator+), blur_filter.hpp:18-36.
// All kernels are streaming, HBM-bound: 128-bit loads/stores, grids sized in multiples of the SM count.
#include "common.cuh"
#include <stdlib.h>

// ------------------------------------------------------------------------------------------------ context
__global__ void k_copy_words(unsigned* __restrict__ dst, const unsigned* __restrict__ src, int n_words) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

__global__ void k_copy_bytes(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

static char* z_take(rspcl_ctx* ctx, size_t bytes) {
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (!ctx->z_host || ctx->z_used + need > ctx->z_cap) return nullptr;
  char* p = ctx->z_host + ctx->z_used;
  ctx->z_used += need;
  return p;
}

cudaError_t small_h2d(rspcl_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  if (bytes == 0) return cudaSuccess;
  char* slot = (bytes <= (64u << 10) && bytes % 4 == 0) ? z_take(ctx, bytes) : nullptr;
  if (!slot) return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream);  // large: plain staged copy
  memcpy(slot, h_src, bytes);
  const int nw = (int)(bytes / 4);
  k_copy_words<<<(nw + 255) / 256 > 8 ? 8 : (nw + 255) / 256, 256, 0, ctx->stream>>>((unsigned*)d_dst, (const unsigned*)slot, nw);
  ctx->launches++;
  return cudaGetLastError();
}

cudaError_t small_d2h(rspcl_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  if (bytes == 0) return cudaSuccess;
  char* slot = (bytes <= (64u << 10) && bytes % 4 == 0 && ((size_t)d_src & 3) == 0) ? z_take(ctx, bytes) : nullptr;
  if (!slot) {
    // large or odd-sized read-back (masks, index dumps, result arrays of big batches): a pinned, mapped staging area
    // filled by a copy kernel and copied out at ctx_sync() -- never an asynchronous copy straight into pageable user
    // memory.  The area is persistent and grow-only (a per-call cudaHostAlloc / cudaFreeHost can stall for seconds behind
    // the driver); only a request that does not fit next to read-backs still pending gets a one-off buffer.
    const size_t need = (bytes + 255) & ~(size_t)255;
    void* bounce = nullptr;
    void* owned = nullptr;
    if (ctx->h_stage_used == 0 && ctx->h_stage_bytes < need) {
      if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
      ctx->h_stage = nullptr;
      ctx->h_stage_bytes = 0;
      const size_t want = need < (8u << 20) ? (8u << 20) : need * 2;
      if (cudaHostAlloc(&ctx->h_stage, want, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) ctx->h_stage_bytes = want;
      else cudaGetLastError();
    }
    if (ctx->h_stage && ctx->h_stage_used + need <= ctx->h_stage_bytes) {
      bounce = (char*)ctx->h_stage + ctx->h_stage_used;
      ctx->h_stage_used += need;
    } else {
      cudaError_t e = cudaHostAlloc(&bounce, bytes, cudaHostAllocMapped | cudaHostAllocPortable);
      if (e != cudaSuccess) return e;
      owned = bounce;
    }
    if (bytes % 4 == 0 && ((size_t)d_src & 3) == 0)
      k_copy_words<<<2 * ctx->sm_count, 256, 0, ctx->stream>>>((unsigned*)bounce, (const unsigned*)d_src, (int)(bytes / 4));
    else
      k_copy_bytes<<<2 * ctx->sm_count, 256, 0, ctx->stream>>>((unsigned char*)bounce, (const unsigned char*)d_src, bytes);
    ctx->launches++;
    ctx->z_pending.push_back({h_dst, bounce, bytes, owned});
    return cudaGetLastError();
  }
  const int nw = (int)(bytes / 4);
  k_copy_words<<<(nw + 255) / 256 > 8 ? 8 : (nw + 255) / 256, 256, 0, ctx->stream>>>((unsigned*)slot, (const unsigned*)d_src, nw);
  ctx->launches++;
  ctx->z_pending.push_back({h_dst, slot, bytes, nullptr});
  return cudaGetLastError();
}

cudaError_t ctx_sync(rspcl_ctx* ctx) {
  // RSPCL_SYNC=spin: poll the stream instead of sleeping in the driver (a host round trip sits on the critical path of
  // every registration call: edge counts, persistent-kernel status)
  static const int spin = [] {
    const char* e = getenv("RSPCL_SYNC");
    return (e && e[0] == 's') ? 1 : 0;
  }();
  cudaError_t e = cudaSuccess;
  if (spin) {
    while ((e = cudaStreamQuery(ctx->stream)) == cudaErrorNotReady) {
    }
  } else {
    e = cudaStreamSynchronize(ctx->stream);
  }
  if (e != cudaSuccess) return e;
  for (auto& p : ctx->z_pending) {
    memcpy(p.dst, p.src, p.n);
    if (p.owned) cudaFreeHost(p.owned);
  }
  ctx->z_pending.clear();
  ctx->z_used = 0;  // everything enqueued so far has executed: the arenas can be recycled
  ctx->h_stage_used = 0;
  return cudaSuccess;
}

extern "C" int rspcl_ctx_create(int device, rspcl_ctx** out) {
  if (!out) return RSPCL_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return RSPCL_ERR_CUDA;
  rspcl_ctx* c = new rspcl_ctx;
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
    delete c;
    return RSPCL_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (cudaHostAlloc((void**)&c->z_host, 4u << 20, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
    c->z_cap = 4u << 20;
  } else {
    c->z_host = nullptr;
    cudaGetLastError();
  }
  // Private stream-ordered pool that keeps freed scratch: after warm-up no call touches the driver allocator, and
  // because the pool belongs to this context's single stream the allocator never inserts a dependency on another
  // context's work (contexts used from different host threads overlap freely).
  cudaMemPoolProps props;
  memset(&props, 0, sizeof(props));
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = device;
  if (cudaMemPoolCreate(&c->pool, &props) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr);
  } else {
    c->pool = nullptr;
    cudaGetLastError();
  }
  *out = c;
  return RSPCL_OK;
}

extern "C" void rspcl_ctx_destroy(rspcl_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  ctx_sync(ctx);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  if (ctx->z_host) cudaFreeHost(ctx->z_host);
  for (int a = 0; a < RSPCL_AUX_STREAMS; ++a) {
    if (ctx->aux[a]) cudaStreamDestroy(ctx->aux[a]);
    if (ctx->ev_join[a]) cudaEventDestroy(ctx->ev_join[a]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* rspcl_last_error(const rspcl_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int rspcl_ctx_sync(rspcl_ctx* ctx) {
  CU(ctx, ctx_sync(ctx));
  return RSPCL_OK;
}
extern "C" int rspcl_timer_start(rspcl_ctx* ctx) {
  CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return RSPCL_OK;
}
extern "C" int rspcl_timer_stop(rspcl_ctx* ctx, float* ms) {
  CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  CU(ctx, cudaEventSynchronize(ctx->ev1));
  CU(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return RSPCL_OK;
}
extern "C" int rspcl_timer_mark(rspcl_ctx* ctx) {
  CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  return RSPCL_OK;
}
extern "C" int rspcl_timer_span(rspcl_ctx* const* ctxs, int n, float* ms) {
  if (!ctxs || n <= 0 || !ms) return RSPCL_ERR_ARG;
  float best = 0.f;
  for (int j = 0; j < n; ++j) CU(ctxs[j], cudaEventSynchronize(ctxs[j]->ev1));
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      float t = 0.f;
      CU(ctxs[i], cudaEventElapsedTime(&t, ctxs[i]->ev0, ctxs[j]->ev1));
      if (t > best) best = t;
    }
  *ms = best;
  return RSPCL_OK;
}
extern "C" long long rspcl_launch_count(const rspcl_ctx* ctx) { return ctx->launches; }

static int prof_drain(rspcl_ctx* ctx) {
  CU(ctx, ctx_sync(ctx));
  for (auto& r : ctx->prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& acc = ctx->prof_acc[ctx->prof_names[r.kernel]];
      acc.ms += ms;
      acc.launches += 1;
      acc.units += r.units;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  ctx->prof_recs.clear();
  return RSPCL_OK;
}
extern "C" int rspcl_profile_enable(rspcl_ctx* ctx, int on) {
  if (!ctx) return RSPCL_ERR_ARG;
  int rc = prof_drain(ctx);
  ctx->prof_on = on != 0;
  return rc;
}
extern "C" int rspcl_profile_reset(rspcl_ctx* ctx) {
  if (!ctx) return RSPCL_ERR_ARG;
  int rc = prof_drain(ctx);
  ctx->prof_acc.clear();
  return rc;
}
extern "C" int rspcl_profile_get(rspcl_ctx* ctx, const char* kernel, double* total_ms, long long* launches, double* units) {
  if (!ctx || !kernel) return RSPCL_ERR_ARG;
  int rc = prof_drain(ctx);
  if (rc) return rc;
  auto it = ctx->prof_acc.find(kernel);
  rspcl_ctx::ProfAcc a;
  if (it != ctx->prof_acc.end()) a = it->second;
  if (total_ms) *total_ms = a.ms;
  if (launches) *launches = a.launches;
  if (units) *units = a.units;
  return RSPCL_OK;
}

extern "C" int rspcl_host_alloc(rspcl_ctx* ctx, size_t bytes, void** out) {
  CU(ctx, cudaMallocHost(out, bytes ? bytes : 1));
  return RSPCL_OK;
}
extern "C" int rspcl_host_free(rspcl_ctx* ctx, void* p) {
  CU(ctx, cudaFreeHost(p));
  return RSPCL_OK;
}

// ------------------------------------------------------------------------------------------------ clouds
extern "C" int rspcl_cloud_create(rspcl_ctx* ctx, int n_seg, int stride, rspcl_cloud** out) {
  if (!ctx || !out || n_seg <= 0 || stride < 0) return RSPCL_ERR_ARG;
  rspcl_cloud* c = new rspcl_cloud;
  c->n_seg = n_seg;
  c->stride = stride;
  c->max_count_hint = stride;
  size_t n = (size_t)n_seg * (size_t)(stride ? stride : 1);
  if (cudaMalloc(&c->pts, n * sizeof(float4)) != cudaSuccess || cudaMalloc(&c->count, n_seg * sizeof(int)) != cudaSuccess) {
    ctx->err = "cudaMalloc failed in rspcl_cloud_create";
    cudaGetLastError();
    if (c->pts) cudaFree(c->pts);
    delete c;
    return RSPCL_ERR_CUDA;
  }
  cudaMemsetAsync(c->count, 0, n_seg * sizeof(int), ctx->stream);
  *out = c;
  return RSPCL_OK;
}

extern "C" void rspcl_cloud_destroy(rspcl_ctx* ctx, rspcl_cloud* c) {
  if (!c) return;
  if (ctx) ctx_sync(ctx);
  cudaFree(c->pts);
  cudaFree(c->count);
  if (c->gray) cudaFree(c->gray);
  delete c;
}

extern "C" int rspcl_cloud_n_seg(const rspcl_cloud* c) { return c->n_seg; }
extern "C" int rspcl_cloud_stride(const rspcl_cloud* c) { return c->stride; }
extern "C" int rspcl_cloud_dims(const rspcl_cloud* c, int* w, int* h) {
  if (w) *w = c->width;
  if (h) *h = c->height;
  return RSPCL_OK;
}

struct Pcl32 {  // pcl::PointXYZRGB in memory
  float x, y, z, w;
  uint32_t rgba;
  uint32_t pad[3];
};

// host-layout staging -> strided float4 (+ gray plane for organized clouds).  One 32 B (or 16 B) load per point.
template <bool PCL32>
__global__ void k_unpack(const void* __restrict__ raw, const int* __restrict__ offsets, const int* __restrict__ count,
                         float4* __restrict__ pts, uint8_t* __restrict__ gray, int stride) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  const long long base = offsets[seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p;
    if (PCL32) {
      const uint4* r = reinterpret_cast<const uint4*>(raw) + 2 * (base + i);
      uint4 a = __ldg(r), b = __ldg(r + 1);
      p = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(b.x));
    } else {
      p = __ldg(reinterpret_cast<const float4*>(raw) + base + i);
    }
    pts[(size_t)seg * stride + i] = p;
    if (gray) {
      uint32_t c = __float_as_uint(p.w);
      // organized_edge_detection.hpp: float((r + g + b) / 3), integer division
      gray[(size_t)seg * stride + i] = (uint8_t)((((c >> 16) & 255u) + ((c >> 8) & 255u) + (c & 255u)) / 3u);
    }
  }
}

template <bool PCL32>
__global__ void k_pack(const float4* __restrict__ pts, const int* __restrict__ offsets, const int* __restrict__ count,
                       void* __restrict__ raw, int stride) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  const long long base = offsets[seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = pts[(size_t)seg * stride + i];
    if (PCL32) {
      uint4* r = reinterpret_cast<uint4*>(raw) + 2 * (base + i);
      r[0] = make_uint4(__float_as_uint(p.x), __float_as_uint(p.y), __float_as_uint(p.z), __float_as_uint(1.0f));
      r[1] = make_uint4(__float_as_uint(p.w), 0u, 0u, 0u);
    } else {
      reinterpret_cast<float4*>(raw)[base + i] = p;
    }
  }
}

// Bulk host<->device copies are issued in large pieces (64 MB: measured 15.4 vs 17.6 ms per e2e step against 4 MB pieces
// on PCIe 5 x16).  Small control traffic of other contexts does not queue behind them: it uses the zero-copy arena.
static cudaError_t chunked_copy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t stream) {
  static const size_t piece = []() {
    const char* e = getenv("RSPCL_COPY_PIECE_MB");  // tuning knob (default 64 MB)
    const long v = e ? atol(e) : 64;
    return (size_t)(v > 0 ? v : 64) << 20;
  }();
  for (size_t off = 0; off < bytes; off += piece) {
    const size_t n = bytes - off < piece ? bytes - off : piece;
    cudaError_t e = cudaMemcpyAsync((char*)dst + off, (const char*)src + off, n, kind, stream);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

int blocks_per_seg(const rspcl_ctx* ctx, int n_seg, int max_count, int threads) {
  // grid-stride kernels: cover max_count once if the chip has room, else cap the batch at ~16 CTAs per SM
  int want = div_up(max_count, threads);
  int cap = (16 * ctx->sm_count) / (n_seg > 0 ? n_seg : 1);
  if (cap < 1) cap = 1;
  int b = want < cap ? want : cap;
  if (b < 1) b = 1;
  if (b > 65535) b = 65535;
  return b;
}

extern "C" int rspcl_cloud_upload(rspcl_ctx* ctx, rspcl_cloud* c, const void* host, int layout, const int32_t* counts,
                                  int n_seg, int width, int height) {
  if (!ctx || !c || !host || !counts || n_seg != c->n_seg) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "upload: bad arguments");
  if (layout != RSPCL_LAYOUT_PCD16 && layout != RSPCL_LAYOUT_PCL32) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "upload: bad layout");
  CU(ctx, cudaSetDevice(ctx->device));
  long long total = 0;
  int maxc = 0;
  std::vector<int> off(n_seg + 1);
  for (int s = 0; s < n_seg; ++s) {
    if (counts[s] < 0 || counts[s] > c->stride) RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "upload: segment %d has %d points, stride %d", s, counts[s], c->stride);
    if (width * height > 0 && counts[s] != width * height) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "upload: organized count mismatch");
    off[s] = (int)total;
    total += counts[s];
    if (counts[s] > maxc) maxc = counts[s];
  }
  off[n_seg] = (int)total;
  if (total >= (1ll << 31)) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "upload: more than 2^31 points in one batch");
  c->width = width;
  c->height = height;
  c->max_count_hint = maxc;
  if (width * height > 0 && !c->gray) CU(ctx, cudaMalloc(&c->gray, (size_t)c->n_seg * c->stride));
  c->gray_valid = false;
  const size_t esz = layout == RSPCL_LAYOUT_PCL32 ? 32 : 16;
  int* d_off = nullptr;
  void* raw = nullptr;
  CU(ctx, scratch_alloc(ctx, &d_off, (size_t)n_seg + 1));
  CU(ctx, scratch_alloc(ctx, (char**)&raw, (size_t)total * esz));
  // small pageable copies: the runtime stages them before returning, so the host vectors may go out of scope
  CU(ctx, small_h2d(ctx, d_off, off.data(), (n_seg + 1) * sizeof(int)));
  CU(ctx, small_h2d(ctx, c->count, counts, n_seg * sizeof(int)));
  if (total > 0) {
    CU(ctx, chunked_copy(raw, host, (size_t)total * esz, cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid(blocks_per_seg(ctx, n_seg, maxc, 256), n_seg);
    ProfScope prof(ctx, "k_unpack", (double)total);
    if (layout == RSPCL_LAYOUT_PCL32)
      k_unpack<true><<<grid, 256, 0, ctx->stream>>>(raw, d_off, c->count, c->pts, c->gray, c->stride);
    else
      k_unpack<false><<<grid, 256, 0, ctx->stream>>>(raw, d_off, c->count, c->pts, c->gray, c->stride);
    LAUNCH_CHECK(ctx);
    c->gray_valid = width * height > 0;  // the unpack kernel wrote the plane
  }
  scratch_free(ctx, d_off);
  scratch_free(ctx, (char*)raw);
  return RSPCL_OK;
}

extern "C" int rspcl_cloud_counts(rspcl_ctx* ctx, const rspcl_cloud* c, int32_t* counts) {
  if (!ctx || !c || !counts) return RSPCL_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, small_d2h(ctx, counts, c->count, c->n_seg * sizeof(int)));
  CU(ctx, ctx_sync(ctx));
  return RSPCL_OK;
}

extern "C" int rspcl_cloud_download(rspcl_ctx* ctx, const rspcl_cloud* c, void* host, int layout, long long capacity_points,
                                    int32_t* counts) {
  if (!ctx || !c) return RSPCL_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  std::vector<int> cnt(c->n_seg);
  int rc = rspcl_cloud_counts(ctx, c, cnt.data());
  if (rc) return rc;
  if (counts) memcpy(counts, cnt.data(), c->n_seg * sizeof(int));
  if (!host) return RSPCL_OK;
  std::vector<int> off(c->n_seg + 1);
  long long total = 0;
  int maxc = 0;
  for (int s = 0; s < c->n_seg; ++s) {
    off[s] = (int)total;
    total += cnt[s];
    if (cnt[s] > maxc) maxc = cnt[s];
  }
  off[c->n_seg] = (int)total;
  if (total > capacity_points) RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "download: %lld points, capacity %lld", total, capacity_points);
  if (total == 0) return RSPCL_OK;
  const size_t esz = layout == RSPCL_LAYOUT_PCL32 ? 32 : 16;
  int* d_off = nullptr;
  void* raw = nullptr;
  CU(ctx, scratch_alloc(ctx, &d_off, (size_t)c->n_seg + 1));
  CU(ctx, scratch_alloc(ctx, (char**)&raw, (size_t)total * esz));
  CU(ctx, small_h2d(ctx, d_off, off.data(), (c->n_seg + 1) * sizeof(int)));
  dim3 grid(blocks_per_seg(ctx, c->n_seg, maxc, 256), c->n_seg);
  if (layout == RSPCL_LAYOUT_PCL32)
    k_pack<true><<<grid, 256, 0, ctx->stream>>>(c->pts, d_off, c->count, raw, c->stride);
  else
    k_pack<false><<<grid, 256, 0, ctx->stream>>>(c->pts, d_off, c->count, raw, c->stride);
  LAUNCH_CHECK(ctx);
  CU(ctx, chunked_copy(host, raw, (size_t)total * esz, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, ctx_sync(ctx));
  scratch_free(ctx, d_off);
  scratch_free(ctx, (char*)raw);
  return RSPCL_OK;
}

extern "C" int rspcl_cloud_invalidate_gray(rspcl_ctx* ctx, rspcl_cloud* c) {
  if (!ctx || !c) return RSPCL_ERR_ARG;
  invalidate_gray(c);
  return RSPCL_OK;
}

// xyz1 of every point, packed (16 B per point) or straight into a 32-byte-pitch host buffer
__global__ void k_pack_xyz1(const float4* __restrict__ pts, const int* __restrict__ offsets, const int* __restrict__ count,
                            float4* __restrict__ out, int stride, int out_pitch_f4) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  const long long base = offsets[seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = pts[(size_t)seg * stride + i];
    p.w = 1.0f;
    out[(size_t)(base + i) * out_pitch_f4] = p;
  }
}

extern "C" int rspcl_cloud_download_xyz_pcl32(rspcl_ctx* ctx, const rspcl_cloud* c, void* host, long long capacity_points,
                                              int mode) {
  if (!ctx || !c || !host || (mode != 0 && mode != 1)) return RSPCL_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  std::vector<int> cnt(c->n_seg);
  int rc = rspcl_cloud_counts(ctx, c, cnt.data());
  if (rc) return rc;
  std::vector<int> off(c->n_seg + 1);
  long long total = 0;
  int maxc = 0;
  for (int s = 0; s < c->n_seg; ++s) {
    off[s] = (int)total;
    total += cnt[s];
    maxc = cnt[s] > maxc ? cnt[s] : maxc;
  }
  off[c->n_seg] = (int)total;
  if (total > capacity_points) RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "download_xyz: %lld points, capacity %lld", total, capacity_points);
  if (total == 0) return RSPCL_OK;
  Scratch scr(ctx);
  int* d_off = nullptr;
  CU(ctx, scr.alloc(&d_off, (size_t)c->n_seg + 1));
  CU(ctx, small_h2d(ctx, d_off, off.data(), (c->n_seg + 1) * sizeof(int)));
  dim3 grid(blocks_per_seg(ctx, c->n_seg, maxc, 256), c->n_seg);
  if (mode == 0) {
    float4* raw = nullptr;
    CU(ctx, scr.alloc(&raw, (size_t)total));
    k_pack_xyz1<<<grid, 256, 0, ctx->stream>>>(c->pts, d_off, c->count, raw, c->stride, 1);
    LAUNCH_CHECK(ctx);
    CU(ctx, cudaMemcpy2DAsync(host, 32, raw, 16, 16, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    void* dev_view = nullptr;  // the pinned buffer as the device sees it (identical under unified addressing)
    CU(ctx, cudaHostGetDevicePointer(&dev_view, host, 0));
    k_pack_xyz1<<<grid, 256, 0, ctx->stream>>>(c->pts, d_off, c->count, (float4*)dev_view, c->stride, 2);
    LAUNCH_CHECK(ctx);
  }
  CU(ctx, ctx_sync(ctx));
  scr.ok();
  return RSPCL_OK;
}

// ------------------------------------------------------------------------------------------------ transform
// K8: p' = R p + t (xyz only, rgba passthrough); 16 B read + 16 B write per point.
__global__ void k_transform(const float4* __restrict__ in, const int* __restrict__ count, const float* __restrict__ T,
                            int broadcast, float4* __restrict__ out, int* __restrict__ out_count, int stride_in,
                            int stride_out) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  __shared__ float M[16];
  if (threadIdx.x < 16) M[threadIdx.x] = T[(broadcast ? 0 : seg * 16) + threadIdx.x];
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0 && out_count) out_count[seg] = n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = in[(size_t)seg * stride_in + i];
    if (finite3(p.x, p.y, p.z)) {
      float3 q = xform_point(M, p.x, p.y, p.z);
      p.x = q.x;
      p.y = q.y;
      p.z = q.z;
    }
    out[(size_t)seg * stride_out + i] = p;
  }
}

int transform_device(rspcl_ctx* ctx, const rspcl_cloud* in, const float* d_T, int broadcast, rspcl_cloud* out) {
  if (out->n_seg != in->n_seg || out->stride < in->max_count_hint)
    RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "transform: output batch (n_seg %d stride %d) too small for input (n_seg %d max %d)",
               out->n_seg, out->stride, in->n_seg, in->max_count_hint);
  dim3 grid(blocks_per_seg(ctx, in->n_seg, in->max_count_hint, 256), in->n_seg);
  k_transform<<<grid, 256, 0, ctx->stream>>>(in->pts, in->count, d_T, broadcast, out->pts, out == in ? nullptr : out->count,
                                             in->stride, out->stride);
  LAUNCH_CHECK(ctx);
  if (out != in) {  // (in place the colours, hence the gray plane, are unchanged)
    out->max_count_hint = in->max_count_hint;
    out->width = in->width;
    out->height = in->height;
    invalidate_gray(out);
  }
  return RSPCL_OK;
}

extern "C" int rspcl_transform(rspcl_ctx* ctx, const rspcl_cloud* in, const float* T, int broadcast, rspcl_cloud* out) {
  if (!ctx || !in || !T || !out) return RSPCL_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  const int nT = broadcast ? 1 : in->n_seg;
  float* d_T = nullptr;
  CU(ctx, scratch_alloc(ctx, &d_T, (size_t)nT * 16));
  CU(ctx, small_h2d(ctx, d_T, T, (size_t)nT * 16 * sizeof(float)));
  int rc = transform_device(ctx, in, d_T, broadcast, out);
  scratch_free(ctx, d_T);
  return rc;
}

// ------------------------------------------------------------------------------------------------ concat / copy
__global__ void k_concat(const float4* __restrict__ a, const int* __restrict__ ca, int sa, const float4* __restrict__ b,
                         const int* __restrict__ cb, int sb, float4* __restrict__ out, int* __restrict__ co, int so,
                         int* __restrict__ overflow) {
  const int seg = blockIdx.y;
  const int na = ca[seg], nb = cb[seg];
  if (na + nb > so) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      atomicExch(overflow, 1);
      co[seg] = 0;
    }
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) co[seg] = na + nb;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < na + nb; i += gridDim.x * blockDim.x)
    out[(size_t)seg * so + i] = i < na ? a[(size_t)seg * sa + i] : b[(size_t)seg * sb + (i - na)];
}

extern "C" int rspcl_concat(rspcl_ctx* ctx, const rspcl_cloud* a, const rspcl_cloud* b, rspcl_cloud* out) {
  if (!ctx || !a || !b || !out || a->n_seg != b->n_seg || out->n_seg != a->n_seg || out == a || out == b)
    RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "concat: bad arguments (out must not alias an input)");
  CU(ctx, cudaSetDevice(ctx->device));
  int* d_over = nullptr;
  CU(ctx, scratch_alloc(ctx, &d_over, 1));
  CU(ctx, cudaMemsetAsync(d_over, 0, sizeof(int), ctx->stream));
  int hint = a->max_count_hint + b->max_count_hint;
  dim3 grid(blocks_per_seg(ctx, a->n_seg, hint, 256), a->n_seg);
  k_concat<<<grid, 256, 0, ctx->stream>>>(a->pts, a->count, a->stride, b->pts, b->count, b->stride, out->pts, out->count,
                                          out->stride, d_over);
  LAUNCH_CHECK(ctx);
  int over = 0;
  CU(ctx, small_d2h(ctx, &over, d_over, sizeof(int)));
  CU(ctx, ctx_sync(ctx));
  scratch_free(ctx, d_over);
  if (over) RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "concat: output stride %d too small", out->stride);
  out->max_count_hint = hint < out->stride ? hint : out->stride;
  out->width = out->height = 0;
  invalidate_gray(out);
  return RSPCL_OK;
}

__global__ void k_copy_seg(const float4* __restrict__ src, const int* __restrict__ cs, float4* __restrict__ dst,
                           int* __restrict__ cd, int cap) {
  const int n = min(*cs, cap);
  if (blockIdx.x == 0 && threadIdx.x == 0) *cd = n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

extern "C" int rspcl_cloud_copy_segment(rspcl_ctx* ctx, const rspcl_cloud* src, int src_seg, rspcl_cloud* dst, int dst_seg) {
  if (!ctx || !src || !dst || src_seg < 0 || src_seg >= src->n_seg || dst_seg < 0 || dst_seg >= dst->n_seg)
    return RSPCL_ERR_ARG;
  if (dst->stride < src->max_count_hint) RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "copy_segment: destination stride too small");
  CU(ctx, cudaSetDevice(ctx->device));
  k_copy_seg<<<2 * ctx->sm_count, 256, 0, ctx->stream>>>(src->pts + (size_t)src_seg * src->stride, src->count + src_seg,
                                                         dst->pts + (size_t)dst_seg * dst->stride, dst->count + dst_seg,
                                                         dst->stride);
  LAUNCH_CHECK(ctx);
  if (src->max_count_hint > dst->max_count_hint) dst->max_count_hint = src->max_count_hint;
  invalidate_gray(dst);  // a reused organized handle must not keep the previous frame's gray plane
  return RSPCL_OK;
}

// ------------------------------------------------------------------------------------------------ crop (K2)
// blur_filter.hpp:27-35: rows [h/5, h/5*4), cols [w/5, w/5*4), written in raster order
__global__ void k_crop35(const float4* __restrict__ in, float4* __restrict__ out, int* __restrict__ out_count, int w, int h,
                         int ow, int oh, int stride_in, int stride_out) {
  const int seg = blockIdx.y;
  const int r0 = h / 5, r1 = h / 5 * 4, c0 = w / 5, c1 = w / 5 * 4;
  const int cw = c1 - c0, n_src = (r1 - r0) * cw, n_out = ow * oh;
  if (blockIdx.x == 0 && threadIdx.x == 0) out_count[seg] = n_out;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += gridDim.x * blockDim.x) {
    // blur_filter.hpp:25 resizes by SHRINKING: entries the copy loop does not reach (w or h not a multiple of 5, e.g.
    // 848x480) keep the input point of the same index
    size_t from = (size_t)i;
    if (i < n_src) from = (size_t)(r0 + i / cw) * w + (c0 + i % cw);
    const float4 p = in[(size_t)seg * stride_in + from];
    out[(size_t)seg * stride_out + i] = p;
  }
}

__global__ void k_gray_from_pts(const float4* __restrict__ pts, const int* __restrict__ count, uint8_t* __restrict__ gray,
                                int stride) {
  const int seg = blockIdx.y;
  const int n = count[seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t c = __float_as_uint(pts[(size_t)seg * stride + i].w);
    gray[(size_t)seg * stride + i] = (uint8_t)((((c >> 16) & 255u) + ((c >> 8) & 255u) + (c & 255u)) / 3u);
  }
}

int ensure_gray(rspcl_ctx* ctx, rspcl_cloud* c) {
  if (c->gray && c->gray_valid) return RSPCL_OK;
  if (!c->gray) CU(ctx, cudaMalloc(&c->gray, (size_t)c->n_seg * (c->stride ? c->stride : 1)));
  dim3 grid(blocks_per_seg(ctx, c->n_seg, c->max_count_hint, 256), c->n_seg);
  k_gray_from_pts<<<grid, 256, 0, ctx->stream>>>(c->pts, c->count, c->gray, c->stride);
  LAUNCH_CHECK(ctx);
  c->gray_valid = true;
  return RSPCL_OK;
}

extern "C" int rspcl_crop35(rspcl_ctx* ctx, const rspcl_cloud* in, rspcl_cloud* out) {
  if (!ctx || !in || !out || in == out) return RSPCL_ERR_ARG;
  if (in->height <= 0) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "crop35: input is not organized");
  const int ow = in->width * 3 / 5, oh = in->height * 3 / 5;
  if (out->n_seg != in->n_seg || out->stride < ow * oh) RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "crop35: output too small");
  CU(ctx, cudaSetDevice(ctx->device));
  dim3 grid(blocks_per_seg(ctx, in->n_seg, ow * oh, 256), in->n_seg);
  k_crop35<<<grid, 256, 0, ctx->stream>>>(in->pts, out->pts, out->count, in->width, in->height, ow, oh, in->stride, out->stride);
  LAUNCH_CHECK(ctx);
  out->width = ow;
  out->height = oh;
  out->max_count_hint = ow * oh;
  invalidate_gray(out);  // rebuilt lazily by the next edge extraction
  return RSPCL_OK;
}

// ------------------------------------------------------------------------------------------------ scan
// Exclusive scan of int32: 1024 threads x 4 items per CTA, block sums scanned recursively.
static constexpr int SCAN_T = 1024, SCAN_ITEMS = 4, SCAN_TILE = SCAN_T * SCAN_ITEMS;

__global__ void k_scan_tiles(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ tile_sums, long long n) {
  __shared__ int warp_tot[32];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    sum += v[k];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = warp_tot[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    warp_tot[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31 && tile_sums) tile_sums[blockIdx.x] = wi;
  }
  __syncthreads();
  int excl = warp_tot[wid] + incl - sum;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = excl;
    excl += v[k];
  }
}

__global__ void k_scan_add(int* __restrict__ out, const int* __restrict__ tile_offsets, long long n) {
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  const int add = tile_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < n) out[base + k] += add;
}

__global__ void k_scan_total(const int* __restrict__ in_last, const int* __restrict__ out_last, int* __restrict__ total) {
  *total = *in_last + *out_last;
}

int rspcl_exclusive_scan_i32(rspcl_ctx* ctx, const int* in, int* out, long long n, int* total_out) {
  if (n <= 0) {
    if (total_out) CU(ctx, cudaMemsetAsync(total_out, 0, sizeof(int), ctx->stream));
    return RSPCL_OK;
  }
  const int tiles = div_up(n, SCAN_TILE);
  int* sums = nullptr;
  if (tiles > 1) CU(ctx, scratch_alloc(ctx, &sums, (size_t)tiles));
  // total = in[n-1] + exclusive[n-1]; read in[n-1] before an in-place scan overwrites it
  int* last_in = nullptr;
  if (total_out) {
    CU(ctx, scratch_alloc(ctx, &last_in, 1));
    CU(ctx, cudaMemcpyAsync(last_in, in + (n - 1), sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  k_scan_tiles<<<tiles, SCAN_T, 0, ctx->stream>>>(in, out, sums, n);
  LAUNCH_CHECK(ctx);
  if (tiles > 1) {
    int rc = rspcl_exclusive_scan_i32(ctx, sums, sums, tiles, nullptr);
    if (rc) return rc;
    k_scan_add<<<tiles, SCAN_T, 0, ctx->stream>>>(out, sums, n);
    LAUNCH_CHECK(ctx);
    scratch_free(ctx, sums);
  }
  if (total_out) {
    k_scan_total<<<1, 1, 0, ctx->stream>>>(last_in, out + (n - 1), total_out);
    LAUNCH_CHECK(ctx);
    scratch_free(ctx, last_in);
  }
  return RSPCL_OK;
}
