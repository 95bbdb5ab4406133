// common.cuh -- shared device/host plumbing of librspcl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/rspcl.h"

struct rspcl_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaMemPool_t pool = nullptr;  // private stream-ordered pool: scratch is never recycled across contexts/streams
  long long launches = 0;
  std::string err;
  int sm_count = 148;
  // grow-only pinned staging buffer for small result read-backs
  void* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  // Zero-copy control channel: small host<->device transfers (counts, indices, 4x4 matrices, convergence states) go
  // through a mapped pinned arena read/written by tiny kernels, so they never queue behind another context's bulk DMA
  // on the copy engines.  Device->host items are copied out of the arena at the next ctx_sync().
  char* z_host = nullptr;
  size_t z_cap = 0, z_used = 0;
  struct ZPending { void* dst; const void* src; size_t n; void* owned; };  // owned: pinned bounce buffer to free
  std::vector<ZPending> z_pending;
  // point-sharded mode (comm.cu)
  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  bool sharded_call = false;  // set for the duration of a *_sharded entry point
  // optional per-kernel event profile
  bool prof_on = false;
  struct ProfRec { cudaEvent_t a, b; int kernel; double units; };
  std::vector<ProfRec> prof_recs;
  std::vector<std::string> prof_names;
  struct ProfAcc { double ms = 0; long long launches = 0; double units = 0; };
  std::map<std::string, ProfAcc> prof_acc;
};

// RAII bracket: records an event pair around the launches issued while it is alive (only when profiling is on)
struct ProfScope {
  rspcl_ctx* ctx;
  int slot = -1;
  ProfScope(rspcl_ctx* c, const char* name, double units) : ctx(c) {
    if (!c->prof_on) return;
    rspcl_ctx::ProfRec r;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    int k = -1;
    for (size_t i = 0; i < c->prof_names.size(); ++i)
      if (c->prof_names[i] == name) k = (int)i;
    if (k < 0) {
      c->prof_names.push_back(name);
      k = (int)c->prof_names.size() - 1;
    }
    r.kernel = k;
    r.units = units;
    cudaEventRecord(r.a, c->stream);
    c->prof_recs.push_back(r);
    slot = (int)c->prof_recs.size() - 1;
  }
  void set_units(double u) {
    if (slot >= 0) ctx->prof_recs[slot].units = u;
  }
  void end() {
    if (slot >= 0 && !ended) cudaEventRecord(ctx->prof_recs[slot].b, ctx->stream);
    ended = true;
  }
  bool ended = false;
  ~ProfScope() { end(); }
};

struct rspcl_cloud {
  float4* pts = nullptr;   // [n_seg * stride] {x,y,z,rgba bits}
  int* count = nullptr;    // [n_seg] device-resident point counts
  uint8_t* gray = nullptr; // organized clouds: (r+g+b)/3 plane written by the upload kernel, [n_seg * stride]
  int n_seg = 0;
  int stride = 0;
  int width = 0, height = 0;  // organized when height > 0
  int max_count_hint = 0;     // host-side upper bound of any segment count (stride if unknown)
};

#define RSPCL_FAIL(ctx, code, ...)                    \
  do {                                                \
    char _b[512];                                     \
    snprintf(_b, sizeof(_b), __VA_ARGS__);            \
    (ctx)->err = _b;                                  \
    return (code);                                    \
  } while (0)

#define CU(ctx, call)                                                                             \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      char _b[512];                                                                               \
      snprintf(_b, sizeof(_b), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      (ctx)->err = _b;                                                                            \
      return RSPCL_ERR_CUDA;                                                                      \
    }                                                                                             \
  } while (0)

#define LAUNCH_CHECK(ctx)            \
  do {                               \
    (ctx)->launches++;               \
    CU(ctx, cudaGetLastError());     \
  } while (0)

// stream-ordered scratch (cudaMallocAsync pool: no synchronisation after warm-up)
template <typename T>
static inline cudaError_t scratch_alloc(rspcl_ctx* ctx, T** p, size_t n) {
  if (ctx->pool) return cudaMallocFromPoolAsync((void**)p, (n ? n : 1) * sizeof(T), ctx->pool, ctx->stream);
  return cudaMallocAsync((void**)p, (n ? n : 1) * sizeof(T), ctx->stream);
}
template <typename T>
static inline void scratch_free(rspcl_ctx* ctx, T* p) {
  if (p) cudaFreeAsync((void*)p, ctx->stream);
}

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

int ensure_stage(rspcl_ctx* ctx, size_t bytes);
cudaError_t small_h2d(rspcl_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
cudaError_t small_d2h(rspcl_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);  // valid after ctx_sync()
cudaError_t ctx_sync(rspcl_ctx* ctx);
int comm_allreduce_f64(rspcl_ctx* ctx, double* buf, size_t n);
// ICP nearest-neighbour cache handed from one align to the next align of the same pairs (icp.cu; owned by the caller)
struct IcpCarry {
  float4* work = nullptr;          // final working cloud of the previous align (.w = cached slot + 1 << 16 | index)
  float* lb = nullptr;             // certified bounds at those positions
  unsigned short* tslot = nullptr; // original target index behind every cached slot
  const float4* tgt_pts = nullptr; // the target the cache refers to
  double max_corr_dist = 0;
  int S = 0, wstride = 0;
  bool valid = false;
};
void icp_carry_free(rspcl_ctx* ctx, IcpCarry* c);
int blocks_per_seg(const rspcl_ctx* ctx, int n_seg, int max_count, int threads);
int transform_device(rspcl_ctx* ctx, const rspcl_cloud* in, const float* d_T, int broadcast, rspcl_cloud* out);
int ensure_gray(rspcl_ctx* ctx, rspcl_cloud* c);

// ---- exclusive scan of int32 on the context stream (scan.cu) ----
int rspcl_exclusive_scan_i32(rspcl_ctx* ctx, const int* in, int* out, long long n, int* total_out /* device, may be null */);

// ---- float helpers that must round exactly like the oracle (no FMA contraction) ----
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }

// x' = ((m00*x + m01*y) + m02*z) + m03, column-major M (pcl/common/impl/transforms.hpp)
__device__ __forceinline__ float3 xform_point(const float* __restrict__ M, float x, float y, float z) {
  float3 o;
  o.x = fadd(fadd(fadd(fmul(M[0], x), fmul(M[4], y)), fmul(M[8], z)), M[12]);
  o.y = fadd(fadd(fadd(fmul(M[1], x), fmul(M[5], y)), fmul(M[9], z)), M[13]);
  o.z = fadd(fadd(fadd(fmul(M[2], x), fmul(M[6], y)), fmul(M[10], z)), M[14]);
  return o;
}

__device__ __forceinline__ bool finite3(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

// static_cast<int>(floor(v)) with x86-64 cvttss2si semantics (NaN / overflow -> INT_MIN)
__device__ __forceinline__ int floor_to_int_x86(float v) {
  float f = floorf(v);
  if (!(f >= -2147483648.0f && f < 2147483648.0f)) return (int)0x80000000;
  return (int)f;
}

// FLANN L2_Simple<float>: ((dx*dx) + dy*dy) + dz*dz
__device__ __forceinline__ float dist2_l2simple(float ax, float ay, float az, float bx, float by, float bz) {
  float d = __fsub_rn(ax, bx);
  float r = fmul(d, d);
  d = __fsub_rn(ay, by);
  r = fadd(r, fmul(d, d));
  d = __fsub_rn(az, bz);
  r = fadd(r, fmul(d, d));
  return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
