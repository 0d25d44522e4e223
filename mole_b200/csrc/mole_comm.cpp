// NCCL plumbing for walker sharding across the GPUs of one box.
// The reference's only "collective" is the in-process gather of per-worker sample vectors
// (concatenate_worker_data, src/vmc/src/vmc.rs:108-130); here it is one ncclAllReduce(fp64, sum) of
// the 62-double accumulator vector per optimisation iteration, plus a few scalars per DMC step.
// NCCL is resolved with dlopen at run time so that the library has no link-time dependency and
// shares the NCCL instance already loaded by the host process (e.g. torch's bundled copy).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstring>
#include <string>
#include <vector>
#include "mole_internal.h"

namespace {
typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm_t;
enum { NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_MAX = 2 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(nccl_unique_id*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi load_api() {
  NcclApi a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.handle) break;
  }
  if (!a.handle) return a;
  a.GetUniqueId = (int (*)(nccl_unique_id*))dlsym(a.handle, "ncclGetUniqueId");
  a.CommInitRank = (int (*)(nccl_comm_t*, int, nccl_unique_id, int))dlsym(a.handle, "ncclCommInitRank");
  a.CommDestroy = (int (*)(nccl_comm_t))dlsym(a.handle, "ncclCommDestroy");
  a.AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(a.handle, "ncclAllReduce");
  a.AllGather = (int (*)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t))dlsym(a.handle, "ncclAllGather");
  a.GetErrorString = (const char* (*)(int))dlsym(a.handle, "ncclGetErrorString");
  a.Send = (int (*)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(a.handle, "ncclSend");
  a.Recv = (int (*)(void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(a.handle, "ncclRecv");
  a.GroupStart = (int (*)())dlsym(a.handle, "ncclGroupStart");
  a.GroupEnd = (int (*)())dlsym(a.handle, "ncclGroupEnd");
  a.ok = a.Send && a.Recv && a.GroupStart && a.GroupEnd && a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather && a.GetErrorString;
  return a;
}

// resolved once per process (the initialisation of a function-local static is thread-safe: contexts may be driven
// from different threads)
NcclApi& api() {
  static NcclApi a = load_api();
  return a;
}

int nccl_fail(mole_ctx_s* ctx, const char* what, int rc) {
  return mole_set_error(ctx, MOLE_ERR_NCCL, std::string(what) + ": " + (api().GetErrorString ? api().GetErrorString(rc) : "?"));
}
}  // namespace

// sum / max allreduce of a few host scalars (DMC per-step energies, branching normalisation)
int32_t mole_comm_allreduce_host(mole_ctx_s* ctx, double* sum_vals, int n_sum, double* max_vals, int n_max) {
  if (!ctx || ctx->nranks <= 1 || !ctx->nccl_comm) return MOLE_OK;
  if (n_sum + n_max > 16) return MOLE_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)ctx->stream;
  cudaSetDevice(ctx->device);
  double* d = ctx->comm_scratch;
  if (n_sum) cudaMemcpyAsync(d, sum_vals, n_sum * sizeof(double), cudaMemcpyHostToDevice, st);
  if (n_max) cudaMemcpyAsync(d + n_sum, max_vals, n_max * sizeof(double), cudaMemcpyHostToDevice, st);
  int rc;
  if (n_sum && (rc = api().AllReduce(d, d, n_sum, NCCL_FLOAT64, NCCL_SUM, ctx->nccl_comm, st)) != 0) return nccl_fail(ctx, "ncclAllReduce", rc);
  if (n_max && (rc = api().AllReduce(d + n_sum, d + n_sum, n_max, NCCL_FLOAT64, NCCL_MAX, ctx->nccl_comm, st)) != 0) return nccl_fail(ctx, "ncclAllReduce", rc);
  if (n_sum) cudaMemcpyAsync(sum_vals, d, n_sum * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (n_max) cudaMemcpyAsync(max_vals, d + n_sum, n_max * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) return mole_set_error(ctx, MOLE_ERR_CUDA, "allreduce_host: stream sync failed");
  return MOLE_OK;
}

// all-gather of n DEVICE doubles per rank on the context stream, no host synchronisation: the DMC block
// loop exchanges {sum w E, sum w, sum w', max w'} with ONE small collective per time step and the
// consuming kernels fold the nranks rows themselves (sums in rank order, max)
int32_t mole_comm_allgather_device(mole_ctx_s* ctx, const double* send_dev, double* recv_dev, int n) {
  if (!ctx || ctx->nranks <= 1 || !ctx->nccl_comm) return MOLE_OK;
  MOLE_RANGE("mole_comm_allgather");
  const int rc = api().AllGather(send_dev, recv_dev, (size_t)n, NCCL_FLOAT64, ctx->nccl_comm, (cudaStream_t)ctx->stream);
  if (rc != 0) return nccl_fail(ctx, "ncclAllGather", rc);
  return MOLE_OK;
}

// all-gather of a few HOST doubles per rank (population rebalancing: total weight and walker count of every rank)
int32_t mole_comm_allgather_host(mole_ctx_s* ctx, const double* mine, int n, double* all) {
  if (!ctx || ctx->nranks <= 1 || !ctx->nccl_comm) {
    for (int i = 0; i < n; ++i) all[i] = mine[i];
    return MOLE_OK;
  }
  if (n * ctx->nranks + n > 64) return MOLE_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)ctx->stream;
  cudaSetDevice(ctx->device);
  double* d = nullptr;
  if (cudaMalloc(&d, 128 * sizeof(double)) != cudaSuccess) return mole_set_error(ctx, MOLE_ERR_CUDA, "cudaMalloc failed");
  cudaMemcpyAsync(d, mine, n * sizeof(double), cudaMemcpyHostToDevice, st);
  const int rc = api().AllGather(d, d + 64, (size_t)n, NCCL_FLOAT64, ctx->nccl_comm, st);
  if (rc != 0) { cudaFree(d); return nccl_fail(ctx, "ncclAllGather", rc); }
  cudaMemcpyAsync(all, d + 64, (size_t)n * ctx->nranks * sizeof(double), cudaMemcpyDeviceToHost, st);
  const cudaError_t ce = cudaStreamSynchronize(st);
  cudaFree(d);
  if (ce != cudaSuccess) return mole_set_error(ctx, MOLE_ERR_CUDA, "allgather_host: stream sync failed");
  return MOLE_OK;
}

// one grouped exchange of walker rows between ranks: moves[i] = {src, dst, first row in the sender's / receiver's
// buffer, count}; row_len doubles per walker
int32_t mole_comm_exchange_rows(mole_ctx_s* ctx, const double* send_dev, double* recv_dev, int row_len,
                                const std::vector<MoleMove>& moves) {
  if (!ctx || ctx->nranks <= 1 || !ctx->nccl_comm) return MOLE_OK;
  cudaStream_t st = (cudaStream_t)ctx->stream;
  int rc = api().GroupStart();
  if (rc != 0) return nccl_fail(ctx, "ncclGroupStart", rc);
  for (const MoleMove& m : moves) {
    if (m.count <= 0) continue;
    if (m.src == ctx->rank && (rc = api().Send(send_dev + (size_t)m.send_first * row_len, (size_t)m.count * row_len, NCCL_FLOAT64, m.dst,
                                                ctx->nccl_comm, st)) != 0) break;
    if (m.dst == ctx->rank && (rc = api().Recv(recv_dev + (size_t)m.recv_first * row_len, (size_t)m.count * row_len, NCCL_FLOAT64, m.src,
                                                ctx->nccl_comm, st)) != 0) break;
  }
  const int rc2 = api().GroupEnd();
  if (rc != 0) return nccl_fail(ctx, "ncclSend/ncclRecv", rc);
  if (rc2 != 0) return nccl_fail(ctx, "ncclGroupEnd", rc2);
  return MOLE_OK;
}

int32_t mole_comm_warm_p2p(mole_ctx_s* ctx) {
  if (!ctx || ctx->nranks <= 1 || !ctx->nccl_comm || ctx->p2p_warm) return MOLE_OK;
  const int n = ctx->nranks;
  double* buf = nullptr;                                       // [0] what this rank sends, [1 + src] what it receives
  if (cudaMalloc(&buf, (size_t)(n + 1) * sizeof(double)) != cudaSuccess) return mole_set_error(ctx, MOLE_ERR_CUDA, "cudaMalloc failed");
  cudaMemsetAsync(buf, 0, (size_t)(n + 1) * sizeof(double), (cudaStream_t)ctx->stream);
  std::vector<MoleMove> moves;
  for (int src = 0; src < n; ++src)
    for (int dst = 0; dst < n; ++dst)
      if (src != dst) moves.push_back(MoleMove{src, dst, 0, 1 + src, 1});
  const int32_t rc = mole_comm_exchange_rows(ctx, buf, buf, 1, moves);
  cudaStreamSynchronize((cudaStream_t)ctx->stream);
  cudaFree(buf);
  if (rc == MOLE_OK) ctx->p2p_warm = 1;
  return rc;
}

extern "C" {

int32_t mole_comm_get_unique_id(uint8_t id[MOLE_NCCL_UNIQUE_ID_BYTES]) {
  if (!id) return MOLE_ERR_INVALID_ARG;
  if (!api().ok) return mole_set_error(nullptr, MOLE_ERR_NCCL, "libnccl.so.2 could not be loaded");
  nccl_unique_id u;
  const int rc = api().GetUniqueId(&u);
  if (rc != 0) return nccl_fail(nullptr, "ncclGetUniqueId", rc);
  static_assert(sizeof(u) == MOLE_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
  memcpy(id, &u, sizeof(u));
  return MOLE_OK;
}

int32_t mole_comm_init(mole_ctx_t ctx, int32_t nranks, int32_t rank, const uint8_t id[MOLE_NCCL_UNIQUE_ID_BYTES]) {
  if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return MOLE_ERR_INVALID_ARG;
  if (!api().ok) return mole_set_error(ctx, MOLE_ERR_NCCL, "libnccl.so.2 could not be loaded");
  if (cudaSetDevice(ctx->device) != cudaSuccess) return mole_set_error(ctx, MOLE_ERR_CUDA, "cudaSetDevice failed");
  nccl_unique_id u;
  memcpy(&u, id, sizeof(u));
  nccl_comm_t comm = nullptr;
  const int rc = api().CommInitRank(&comm, nranks, u, rank);
  if (rc != 0) return nccl_fail(ctx, "ncclCommInitRank", rc);
  ctx->nccl_comm = comm;
  ctx->nranks = nranks;
  ctx->rank = rank;
  if (!ctx->comm_scratch && cudaMalloc(&ctx->comm_scratch, 16 * sizeof(double)) != cudaSuccess)
    return mole_set_error(ctx, MOLE_ERR_CUDA, "cudaMalloc(comm_scratch) failed");
  return MOLE_OK;
}

int32_t mole_comm_destroy(mole_ctx_t ctx) {
  if (!ctx) return MOLE_ERR_INVALID_ARG;
  if (ctx->nccl_comm) api().CommDestroy(ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->nranks = 1;
  ctx->rank = 0;
  if (ctx->comm_scratch) { cudaFree(ctx->comm_scratch); ctx->comm_scratch = nullptr; }
  return MOLE_OK;
}

int32_t mole_acc_allreduce(mole_ens_t e) {
  if (!e) return MOLE_ERR_INVALID_ARG;
  mole_ctx_s* ctx = e->ctx;
  if (ctx->nranks <= 1) return MOLE_OK;
  if (!ctx->nccl_comm) return mole_set_error(ctx, MOLE_ERR_NCCL, "mole_acc_allreduce before mole_comm_init");
  MOLE_RANGE("mole_acc_allreduce");
  cudaSetDevice(ctx->device);
  const int rc = api().AllReduce(e->acc, e->acc, ACC_DEV_LEN, NCCL_FLOAT64, NCCL_SUM, ctx->nccl_comm, (cudaStream_t)ctx->stream);
  if (rc != 0) return nccl_fail(ctx, "ncclAllReduce", rc);
  return MOLE_OK;
}

int32_t mole_gram_allreduce(mole_ens_t e) {
  if (!e) return MOLE_ERR_INVALID_ARG;
  mole_ctx_s* ctx = e->ctx;
  if (ctx->nranks <= 1 || !e->gram) return MOLE_OK;
  if (!ctx->nccl_comm) return mole_set_error(ctx, MOLE_ERR_NCCL, "mole_gram_allreduce before mole_comm_init");
  MOLE_RANGE("mole_gram_allreduce");
  cudaSetDevice(ctx->device);
  const int rc = api().AllReduce(e->gram, e->gram, 48 * 48, NCCL_FLOAT64, NCCL_SUM, ctx->nccl_comm, (cudaStream_t)ctx->stream);
  if (rc != 0) return nccl_fail(ctx, "ncclAllReduce", rc);
  return MOLE_OK;
}

}  // extern "C"
