// comm.cu -- point-sharded multi-GPU mode (BASELINE configs[4]): one process per GPU, every rank holds a
// shard of the SOURCE points and a replica of the target; the per-iteration partial sums (17 fp64 per ICP pair,
// 28 fp64 per NDT derivative evaluation) are combined with an NCCL all-reduce over NVLink on the context's stream,
// after which every rank runs the identical solve (no broadcast).
//
// NCCL is bound lazily (dlopen of libnccl.so.2 -- the copy already loaded by torch when the caller is bench.py, else
// the system library), so the library has no hard dependency on it: single-GPU users never touch it.
#include "common.cuh"
#include <dlfcn.h>

namespace {
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(NcclComm);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef const char* (*fn_get_error_string)(int);

struct NcclApi {
  void* lib = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_reduce all_reduce = nullptr;
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
      api.get_error_string = (fn_get_error_string)dlsym(api.lib, "ncclGetErrorString");
    }
  }
  if (!api.lib || !api.get_unique_id || !api.comm_init_rank || !api.comm_destroy || !api.all_reduce) return nullptr;
  return &api;
}
}  // namespace

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
  return RSPCL_OK;
}

extern "C" int rspcl_comm_destroy(rspcl_ctx* ctx) {
  if (!ctx) return RSPCL_ERR_ARG;
  NcclApi* a = nccl();
  if (a && ctx->nccl_comm) {
    cudaStreamSynchronize(ctx->stream);
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
