// comm.cu -- point-sharded multi-GPU mode (BASELINE configs[4]): one process per GPU, every rank holds a
// shard of the SOURCE points and a replica of the target; the per-iteration partial sums (17 fp64 per ICP pair,
// 28 fp64 per NDT derivative evaluation) are combined with an NCCL all-reduce over NVLink on the context's stream,
// after which every rank runs the identical solve (no broadcast).
//
// NCCL is bound lazily (dlopen of libnccl.so.2 -- the copy already loaded by torch when the caller is bench.py, else
// the system library), so the library has no hard dependency on it: single-GPU users never touch it.
#include "common.cuh"
#include <dlfcn.h>
#include <stdlib.h>

namespace {
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(NcclComm);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef const char* (*fn_get_error_string)(int);

struct NcclApi {
  void* lib = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_get_error_string get_error_string = nullptr;
};

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) {
      api.get_unique_id = (fn_get_unique_id)dlsym(api.lib, "ncclGetUniqueId");
      api.comm_init_rank = (fn_comm_init_rank)dlsym(api.lib, "ncclCommInitRank");
      api.comm_destroy = (fn_comm_destroy)dlsym(api.lib, "ncclCommDestroy");
      api.all_reduce = (fn_all_reduce)dlsym(api.lib, "ncclAllReduce");
      api.all_gather = (fn_all_gather)dlsym(api.lib, "ncclAllGather");
      api.get_error_string = (fn_get_error_string)dlsym(api.lib, "ncclGetErrorString");
    }
  }
  if (!api.lib || !api.get_unique_id || !api.comm_init_rank || !api.comm_destroy || !api.all_reduce) return nullptr;
  return &api;
}
}  // namespace

// Peer-memory exchange buffers: every rank allocates [2 parities][nranks][PX_WORDS] doubles + flags, exports the
// allocation through CUDA IPC, the handles travel through one ncclAllGather, and every rank maps every peer's buffer
// (NVLink P2P on an NVSwitch box).  RSPCL_PEER_XCHG=0 keeps the ncclAllReduce path.
static size_t px_bytes(int nranks) {
  return ((size_t)2 * nranks * PX_WORDS + (size_t)2 * nranks * PX_MAXPAIRS) * 8;
}

static void peer_teardown(rspcl_ctx* ctx) {
  for (int r = 0; r < PX_MAXRANKS; ++r) {
    if (ctx->px_peer[r] && r != ctx->rank) cudaIpcCloseMemHandle(ctx->px_peer[r]);
    ctx->px_peer[r] = nullptr;
  }
  if (ctx->px_local) cudaFree(ctx->px_local);
  if (ctx->px_err) cudaFree(ctx->px_err);
  ctx->px_local = nullptr;
  ctx->px_err = nullptr;
  ctx->px_seq = 0;
  cudaGetLastError();
}

static void peer_setup(rspcl_ctx* ctx, NcclApi* a) {
  const char* env = getenv("RSPCL_PEER_XCHG");
  if ((env && env[0] == '0') || ctx->nranks < 2 || ctx->nranks > PX_MAXRANKS || !a->all_gather) return;
  const size_t bytes = px_bytes(ctx->nranks);
  cudaIpcMemHandle_t mine;
  char* d_handles = nullptr;
  bool ok = cudaMalloc(&ctx->px_local, bytes) == cudaSuccess && cudaMemset(ctx->px_local, 0, bytes) == cudaSuccess &&
            cudaMalloc((void**)&ctx->px_err, sizeof(int)) == cudaSuccess && cudaMemset(ctx->px_err, 0, sizeof(int)) == cudaSuccess &&
            cudaIpcGetMemHandle(&mine, ctx->px_local) == cudaSuccess &&
            cudaMalloc((void**)&d_handles, (size_t)ctx->nranks * sizeof(mine)) == cudaSuccess;
  std::vector<cudaIpcMemHandle_t> all(ctx->nranks);
  if (ok) {
    ok = cudaMemcpy(d_handles + (size_t)ctx->rank * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice) == cudaSuccess &&
         a->all_gather(d_handles + (size_t)ctx->rank * sizeof(mine), d_handles, sizeof(mine), /*ncclInt8*/ 0, (NcclComm)ctx->nccl_comm,
                       ctx->stream) == 0 &&
         cudaStreamSynchronize(ctx->stream) == cudaSuccess &&
         cudaMemcpy(all.data(), d_handles, (size_t)ctx->nranks * sizeof(mine), cudaMemcpyDeviceToHost) == cudaSuccess;
  }
  // every rank must take the same decision: a rank whose setup failed still went through the all-gather above when it
  // could; the final agreement is one more tiny all-reduce of the ok flags
  for (int r = 0; ok && r < ctx->nranks; ++r) {
    if (r == ctx->rank) {
      ctx->px_peer[r] = ctx->px_local;
    } else if (cudaIpcOpenMemHandle(&ctx->px_peer[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      ctx->px_peer[r] = nullptr;
      ok = false;
    }
  }
  if (d_handles) cudaFree(d_handles);
  double flag = ok ? 0.0 : 1.0, *d_flag = nullptr;
  if (cudaMalloc((void**)&d_flag, sizeof(double)) == cudaSuccess) {
    cudaMemcpy(d_flag, &flag, sizeof(double), cudaMemcpyHostToDevice);
    if (a->all_reduce(d_flag, d_flag, 1, /*ncclFloat64*/ 8, /*ncclSum*/ 0, (NcclComm)ctx->nccl_comm, ctx->stream) == 0 &&
        cudaStreamSynchronize(ctx->stream) == cudaSuccess)
      cudaMemcpy(&flag, d_flag, sizeof(double), cudaMemcpyDeviceToHost);
    else
      flag = 1.0;
    cudaFree(d_flag);
  } else {
    flag = 1.0;
  }
  cudaGetLastError();
  if (flag != 0.0) peer_teardown(ctx);  // somebody could not map a peer: everybody stays on NCCL
}

bool comm_peer_ready(const rspcl_ctx* ctx, int n_pairs, int vals_per_pair) {
  const char* env = getenv("RSPCL_PEER_XCHG");  // "0": keep ncclAllReduce for this call (must be set on every rank alike)
  if (env && env[0] == '0') return false;
  return ctx->px_local != nullptr && ctx->nranks > 1 && n_pairs <= PX_MAXPAIRS && (long long)n_pairs * vals_per_pair <= PX_WORDS;
}

PeerX comm_peer_next(rspcl_ctx* ctx) {
  PeerX X;
  for (int r = 0; r < PX_MAXRANKS; ++r) X.peer[r] = (double*)ctx->px_peer[r];
  X.nranks = ctx->nranks;
  X.rank = ctx->rank;
  X.seq = ctx->px_seq++;
  X.err = ctx->px_err;
  return X;
}

int comm_peer_check(rspcl_ctx* ctx) {
  if (!ctx->px_err) return RSPCL_OK;
  int e = 0;
  CU(ctx, cudaMemcpy(&e, ctx->px_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) RSPCL_FAIL(ctx, RSPCL_ERR_CUDA, "point-sharded exchange timed out: a rank did not reach the collective");
  return RSPCL_OK;
}

extern "C" int rspcl_comm_unique_id(void* id128) {
  NcclApi* a = nccl();
  if (!a || !id128) return RSPCL_ERR_ARG;
  NcclUniqueId id;
  if (a->get_unique_id(&id) != 0) return RSPCL_ERR_CUDA;
  memcpy(id128, &id, 128);
  return RSPCL_OK;
}

extern "C" int rspcl_comm_init(rspcl_ctx* ctx, int nranks, int rank, const void* id128) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return RSPCL_ERR_ARG;
  NcclApi* a = nccl();
  if (!a) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "comm_init: libnccl.so.2 could not be loaded");
  CU(ctx, cudaSetDevice(ctx->device));
  NcclUniqueId id;
  memcpy(&id, id128, 128);
  NcclComm comm = nullptr;
  const int rc = a->comm_init_rank(&comm, nranks, id, rank);
  if (rc != 0) RSPCL_FAIL(ctx, RSPCL_ERR_CUDA, "ncclCommInitRank failed: %s", a->get_error_string ? a->get_error_string(rc) : "?");
  ctx->nccl_comm = comm;
  ctx->nranks = nranks;
  ctx->rank = rank;
  peer_setup(ctx, a);  // best effort: without it the partial sums go through ncclAllReduce
  return RSPCL_OK;
}

extern "C" int rspcl_comm_destroy(rspcl_ctx* ctx) {
  if (!ctx) return RSPCL_ERR_ARG;
  NcclApi* a = nccl();
  if (a && ctx->nccl_comm) {
    cudaStreamSynchronize(ctx->stream);
    peer_teardown(ctx);
    a->comm_destroy((NcclComm)ctx->nccl_comm);
  }
  ctx->nccl_comm = nullptr;
  ctx->nranks = 1;
  ctx->rank = 0;
  return RSPCL_OK;
}

// sum-all-reduce of n doubles in place on the context stream (no-op for a single rank)
int comm_allreduce_f64(rspcl_ctx* ctx, double* buf, size_t n) {
  if (!ctx->nccl_comm || ctx->nranks <= 1) return RSPCL_OK;
  NcclApi* a = nccl();
  if (!a) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "allreduce: NCCL not available");
  const int rc = a->all_reduce(buf, buf, n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, (NcclComm)ctx->nccl_comm, ctx->stream);
  if (rc != 0) RSPCL_FAIL(ctx, RSPCL_ERR_CUDA, "ncclAllReduce failed: %s", a->get_error_string ? a->get_error_string(rc) : "?");
  ctx->launches++;
  return RSPCL_OK;
}
