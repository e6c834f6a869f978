// Multi-GPU plumbing for the configs that shard naturally: row blocks of a tall A are
// sketched locally and the l x n sketch is summed over NVLink with ONE ncclAllReduce per adaptive round; the
// CholeskyQR2 tail needs one k x k all-reduce per pass.  Everything else (QRCP, T-solve, small factors) runs
// replicated on identical inputs, so k and p agree on every rank without a broadcast.
// The reference has no distributed code at all; this is additive.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2") (the torch-bundled 2.28 when torch is in the
// process, else the system 2.27), so libbrapprox.so itself has no link-time NCCL dependency.
#include "common.cuh"
#include <dlfcn.h>

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
typedef int nccl_result;
constexpr int NCCL_FLOAT64 = 8;   // ncclFloat64 / ncclDouble
constexpr int NCCL_SUM = 0;       // ncclSum

struct NcclApi {
  void* h = nullptr;
  nccl_result (*GetUniqueId)(nccl_uid*) = nullptr;
  nccl_result (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
  nccl_result (*CommDestroy)(nccl_comm) = nullptr;
  nccl_result (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(nccl_result) = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  if (a.h) return a;
  a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!a.h) return a;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
  a.AllReduce = (decltype(a.AllReduce))dlsym(a.h, "ncclAllReduce");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString;
  return a;
}

}  // namespace

// in-place sum over the ranks of ctx's communicator, on ctx's stream; no-op for a single rank
int bra_allreduce_sum_f64(bra_ctx* ctx, double* buf, int64_t count) {
  if (ctx->world <= 1 || count <= 0) return BRA_OK;
  ProfScope ps(ctx, BRA_PROF_COMM);
  nccl_result r = api().AllReduce(buf, buf, (size_t)count, NCCL_FLOAT64, NCCL_SUM, (nccl_comm)ctx->nccl_comm, ctx->stream);
  if (r != 0) {
    ctx->set_error(std::string("ncclAllReduce: ") + api().GetErrorString(r));
    return BRA_ERR_COMM;
  }
  ctx->collectives++;
  return BRA_OK;
}

extern "C" {

int bra_comm_unique_id(void* id128) {
  if (!id128) return -1;
  if (!api().ok) return BRA_ERR_COMM;
  nccl_uid id;
  if (api().GetUniqueId(&id) != 0) return BRA_ERR_COMM;
  std::memcpy(id128, &id, 128);
  return BRA_OK;
}

int bra_comm_init(bra_ctx* ctx, const void* id128, int rank, int world) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(id128 != nullptr, 2, "id128");
  BRA_CHECK_ARG(world >= 1, 4, "world");
  BRA_CHECK_ARG(rank >= 0 && rank < world, 3, "rank");
  if (!api().ok) {
    ctx->set_error("libnccl.so.2 could not be loaded");
    return BRA_ERR_COMM;
  }
  BRA_CUDA(cudaSetDevice(ctx->device));
  if (ctx->nccl_comm) {
    api().CommDestroy((nccl_comm)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  nccl_uid id;
  std::memcpy(&id, id128, 128);
  nccl_comm comm = nullptr;
  nccl_result r = api().CommInitRank(&comm, world, id, rank);
  if (r != 0) {
    ctx->set_error(std::string("ncclCommInitRank: ") + api().GetErrorString(r));
    return BRA_ERR_COMM;
  }
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->world = world;
  return BRA_OK;
}

int bra_comm_destroy(bra_ctx* ctx) {
  if (!ctx) return -1;
  if (ctx->nccl_comm && api().ok) {
    cudaStreamSynchronize(ctx->stream);
    api().CommDestroy((nccl_comm)ctx->nccl_comm);
  }
  ctx->nccl_comm = nullptr;
  ctx->rank = 0;
  ctx->world = 1;
  ctx->shard_row0 = 0;
  ctx->shard_m_global = 0;
  return BRA_OK;
}

int bra_set_row_shard(bra_ctx* ctx, int64_t row0, int64_t m_global) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(row0 >= 0, 2, "row0");
  BRA_CHECK_ARG(m_global >= 0, 3, "m_global");
  ctx->shard_row0 = row0;
  ctx->shard_m_global = m_global;
  return BRA_OK;
}

uint64_t bra_collective_count(bra_ctx* ctx) { return ctx ? ctx->collectives : 0; }

}  // extern "C"
