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

#define RSPCL_AUX_STREAMS 7

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
  size_t h_stage_bytes = 0, h_stage_used = 0;
  // Zero-copy control channel: small host<->device transfers (counts, indices, 4x4 matrices, convergence states) go
  // through a mapped pinned arena read/written by tiny kernels, so they never queue behind another context's bulk DMA
  // on the copy engines.  Device->host items are copied out of the arena at the next ctx_sync().
  char* z_host = nullptr;
  size_t z_cap = 0, z_used = 0;
  struct ZPending { void* dst; const void* src; size_t n; void* owned; };  // owned: pinned bounce buffer to free
  std::vector<ZPending> z_pending;
  // forked streams for launches that run side by side (persistent ICP: one launch per cluster size)
  cudaStream_t aux[RSPCL_AUX_STREAMS] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[RSPCL_AUX_STREAMS] = {};
  bool persist_ready = false;
  int persist_late = 0;   // consecutive launches with a late cluster start
  int persist_clean = 0;  // consecutive launches without one
  int persist_extra = 2;  // SMs the persistent-ICP planner may plan beyond (+) or below (-) the SM count; lowered when a launch
                          // is seen to need a second wave (icp.cu)
  double cluster_weight[9] = {0, 1, 2, 3, 4, 5, 6, 7, 8};  // SMs a cluster of c CTAs occupies (occupancy query)
  // point-sharded mode (comm.cu)
  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  // one-shot peer-memory exchange of the per-iteration partial sums (comm.cu): every rank's buffer mapped into every
  // other rank through CUDA IPC over NVLink; falls back to ncclAllReduce when the mapping could not be set up
  void* px_local = nullptr;
  void* px_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int* px_err = nullptr;            // device flag: an exchange timed out
  unsigned long long px_seq = 0;    // exchanges issued so far (identical on every rank: the calls are collective)
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
  bool gray_valid = false; // false: pts were rewritten since the plane was computed (ensure_gray rebuilds it lazily)
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

cudaError_t small_h2d(rspcl_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
cudaError_t small_d2h(rspcl_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);  // valid after ctx_sync()
cudaError_t ctx_sync(rspcl_ctx* ctx);
int comm_allreduce_f64(rspcl_ctx* ctx, double* buf, size_t n);

// ---- one-shot all-reduce of a few doubles per pair through peer memory (point-sharded mode)
constexpr int PX_MAXRANKS = 8;
constexpr int PX_WORDS = 8192;   // doubles per (parity, source rank) slot
constexpr int PX_MAXPAIRS = 256; // flag slots per (parity, source rank)
struct PeerX {                   // passed by value to the kernels that exchange
  double* peer[PX_MAXRANKS];     // rank r's buffer as THIS rank sees it (peer[rank] = the local buffer)
  int nranks, rank;
  unsigned long long seq;        // index of this exchange
  int* err;
};
bool comm_peer_ready(const rspcl_ctx* ctx, int n_pairs, int vals_per_pair);
PeerX comm_peer_next(rspcl_ctx* ctx);   // descriptor of the next exchange (advances the sequence number)
int comm_peer_check(rspcl_ctx* ctx);    // after a sync: did any exchange time out?

// Called by one whole warp per pair: lane k < n_vals contributes v; returns the sum over the ranks of value k (in rank
// order, so every rank gets the same bits).  Data first, then a system-scope fence, then one flag per destination; the
// receiver spins on its LOCAL flags (bounded: a missing peer sets *err instead of hanging the GPU).
static __device__ __forceinline__ double peer_allreduce_warp(const PeerX& X, double v, int lane, int n_vals, int pair) {
  const int par = (int)(X.seq & 1ull);
  const size_t flags_off = (size_t)2 * X.nranks * PX_WORDS;  // in 8-byte words
  if (lane < n_vals)
    for (int r = 0; r < X.nranks; ++r) X.peer[r][((size_t)par * X.nranks + X.rank) * PX_WORDS + (size_t)pair * n_vals + lane] = v;
  __threadfence_system();
  __syncwarp();
  if (lane < X.nranks) {
    volatile unsigned long long* f =
        reinterpret_cast<volatile unsigned long long*>(X.peer[lane] + flags_off) + ((size_t)par * X.nranks + X.rank) * PX_MAXPAIRS + pair;
    *f = X.seq + 1;  // (st.volatile to peer memory after the fence: the data is visible before the flag)
    volatile unsigned long long* mine =
        reinterpret_cast<volatile unsigned long long*>(X.peer[X.rank] + flags_off) + ((size_t)par * X.nranks + lane) * PX_MAXPAIRS + pair;
    const long long t0 = clock64();
    while (*mine < X.seq + 1) {
      if (clock64() - t0 > 4000000000ll) {  // ~2 s: a peer never arrived
        *X.err = 1;
        break;
      }
    }
  }
  __syncwarp();
  __threadfence_system();
  double out = 0.0;
  if (lane < n_vals) {
    const volatile double* L = X.peer[X.rank];
    for (int r = 0; r < X.nranks; ++r) out += L[((size_t)par * X.nranks + r) * PX_WORDS + (size_t)pair * n_vals + lane];
  }
  return out;
}
// Options of the device-level ICP align (icp.cu)
struct IcpAlignOpts {
  int* d_corr_out = nullptr;                 // device correspondence dump [iteration][pair][stride] (match index or -1)
  int corr_iters = 0;                        // iterations to dump (1 = PCL-style first correspondences)
  rspcl_icp_result* h_results2 = nullptr;    // run a SECOND align from identity on the first one's output (icp:108-111)
  const int* h_src_counts = nullptr;         // host copy of the source counts if the caller has one (saves a read-back)
  const unsigned char* h_skip = nullptr;     // pairs to leave alone (finished elsewhere)
  bool no_persist = false;                   // go straight to the global-memory path
};

// Every stream-ordered scratch buffer of one call, released on EVERY exit path (error returns included).  A call that
// returns an error without having reached its ctx_sync() may leave read-backs pending whose destinations are about to
// go out of scope: unless ok() was called, the destructor drains the stream and drops them.
struct Scratch {
  rspcl_ctx* ctx;
  std::vector<void*> ptrs;
  bool fine = false;
  explicit Scratch(rspcl_ctx* c) : ctx(c) {}
  template <typename T>
  cudaError_t alloc(T** p, size_t n) {
    const cudaError_t e = scratch_alloc(ctx, p, n);
    if (e == cudaSuccess) ptrs.push_back((void*)*p);
    return e;
  }
  void ok() { fine = true; }
  ~Scratch() {
    if (!fine) {
      cudaStreamSynchronize(ctx->stream);
      for (auto& q : ctx->z_pending)
        if (q.owned) cudaFreeHost(q.owned);
      ctx->z_pending.clear();
      ctx->z_used = 0;
      ctx->h_stage_used = 0;
      cudaGetLastError();
    }
    for (void* q : ptrs) cudaFreeAsync(q, ctx->stream);
  }
};

// scratch cloud batch (stream-ordered allocation)
struct TmpCloud {
  rspcl_cloud c;
  rspcl_ctx* ctx;
  explicit TmpCloud(rspcl_ctx* x) : ctx(x) {}
  int init(int n_seg, int stride) {
    c.n_seg = n_seg;
    c.stride = stride;
    c.max_count_hint = stride;
    if (scratch_alloc(ctx, &c.pts, (size_t)n_seg * (stride ? stride : 1)) != cudaSuccess) return RSPCL_ERR_CUDA;
    if (scratch_alloc(ctx, &c.count, (size_t)n_seg) != cudaSuccess) return RSPCL_ERR_CUDA;
    return RSPCL_OK;
  }
  ~TmpCloud() {
    scratch_free(ctx, c.pts);
    scratch_free(ctx, c.count);
    if (c.gray) cudaFree(c.gray);
  }
};
int blocks_per_seg(const rspcl_ctx* ctx, int n_seg, int max_count, int threads);
int transform_device(rspcl_ctx* ctx, const rspcl_cloud* in, const float* d_T, int broadcast, rspcl_cloud* out);
int ensure_gray(rspcl_ctx* ctx, rspcl_cloud* c);
// an operation rewrote the colours of c->pts: the cached gray plane (if any) is stale
static inline void invalidate_gray(rspcl_cloud* c) { c->gray_valid = false; }

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

// ---- shared-memory / cluster PTX helpers
static __device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// 8-byte store into a (possibly remote) CTA of the cluster that completes 8 transaction bytes on that CTA's mbarrier:
// data and signal travel together, no fence and no cluster-wide barrier
static __device__ __forceinline__ void st_async_f64(unsigned remote_addr, double v, unsigned remote_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
               "l"(__double_as_longlong(v)), "r"(remote_mbar)
               : "memory");
}
static __device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_arrive_expect_tx(unsigned addr, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_wait_cluster(unsigned addr, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}


__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
