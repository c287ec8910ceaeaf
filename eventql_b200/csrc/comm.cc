// comm.cc - multi-GPU: NCCL communicator (one process per GPU) and the merge of partial aggregate tables.
// NCCL is loaded with dlopen so that the library also loads on hosts without it; a missing NCCL is a hard
// error at evqgpu_comm_init, never a silent single-GPU fallback.
#include <dlfcn.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "context.h"
#include "query.h"

using namespace evq;

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum ncclDataType { ncclUint8 = 1, ncclUint64 = 5, ncclFloat64 = 8 };
enum ncclRedOp { ncclSum = 0, ncclMax = 2, ncclMin = 3 };

struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

Nccl& nccl() {
  static Nccl n;
  if (n.lib) return n;
  const char* names[] = {getenv("EVQGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* name : names) {
    if (!name) continue;
    n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (n.lib) break;
  }
  if (!n.lib) fail(EVQGPU_ERR_CUDA, "cannot load NCCL (libnccl.so.2): %s", dlerror());
#define EVQ_SYM(field, sym)                                                     \
  *(void**) (&n.field) = dlsym(n.lib, sym);                                      \
  if (!n.field) fail(EVQGPU_ERR_CUDA, "NCCL symbol %s not found", sym);
  EVQ_SYM(GetUniqueId, "ncclGetUniqueId")
  EVQ_SYM(CommInitRank, "ncclCommInitRank")
  EVQ_SYM(CommDestroy, "ncclCommDestroy")
  EVQ_SYM(AllGather, "ncclAllGather")
  EVQ_SYM(AllReduce, "ncclAllReduce")
  EVQ_SYM(Send, "ncclSend")
  EVQ_SYM(Recv, "ncclRecv")
  EVQ_SYM(GroupStart, "ncclGroupStart")
  EVQ_SYM(GroupEnd, "ncclGroupEnd")
  EVQ_SYM(GetErrorString, "ncclGetErrorString")
#undef EVQ_SYM
  return n;
}

#define EVQ_NCCL(expr)                                                                         \
  do {                                                                                         \
    int _r = (expr);                                                                           \
    if (_r != ncclSuccess) fail(EVQGPU_ERR_CUDA, "%s failed: %s", #expr, nccl().GetErrorString(_r)); \
  } while (0)

}  // namespace

extern "C" {

int evqgpu_comm_unique_id(void* id_out) {
  return guarded([&] {
    if (!id_out) fail(EVQGPU_ERR_ARG, "evqgpu_comm_unique_id: null argument");
    ncclUniqueId id;
    EVQ_NCCL(nccl().GetUniqueId(&id));
    static_assert(sizeof(id) == EVQGPU_COMM_ID_BYTES, "id size");
    memcpy(id_out, &id, sizeof(id));
  });
}

int evqgpu_comm_init(evqgpu_ctx* ctx, const void* id, int rank, int nranks) {
  return guarded([&] {
    if (!ctx || !id) fail(EVQGPU_ERR_ARG, "evqgpu_comm_init: null argument");
    if (ctx->nccl_comm) fail(EVQGPU_ERR_ARG, "communicator already initialised");
    use_device(ctx);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm;
    EVQ_NCCL(nccl().CommInitRank(&comm, nranks, uid, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
  });
}

int evqgpu_comm_destroy(evqgpu_ctx* ctx) {
  return guarded([&] {
    if (!ctx || !ctx->nccl_comm) return;
    nccl().CommDestroy((ncclComm_t) ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
  });
}

int evqgpu_query_merge(evqgpu_query* q) {
  return guarded([&] {
    if (!q) fail(EVQGPU_ERR_ARG, "evqgpu_query_merge: null query");
    evq::merge_query(*q);
  });
}

}  // extern "C"

namespace evq {

// Used by merge.cu
void comm_all_gather(evqgpu_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
  EVQ_NCCL(nccl().AllGather(send, recv, bytes_per_rank, ncclUint8, (ncclComm_t) ctx->nccl_comm, ctx->stream));
}

// small host-side all-gather (bounds, counts): staged through device memory because NCCL moves device buffers
std::vector<uint64_t> comm_all_gather_host(evqgpu_ctx* ctx, const std::vector<uint64_t>& mine) {
  use_device(ctx);
  const size_t n = mine.size();
  DevBuf send, recv;
  send.alloc(std::max<size_t>(n, 1) * 8);
  recv.alloc(std::max<size_t>(n, 1) * 8 * ctx->nranks);
  std::vector<uint64_t> all(n * ctx->nranks);
  if (n == 0) return all;
  EVQ_CUDA(cudaMemcpyAsync(send.p, mine.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream));
  comm_all_gather(ctx, send.p, recv.p, n * 8);
  EVQ_CUDA(cudaMemcpyAsync(all.data(), recv.p, all.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
  return all;
}

void comm_all_to_all(evqgpu_ctx* ctx, const void* send, const uint64_t* send_off, const uint64_t* send_bytes, void* recv,
                     const uint64_t* recv_off, const uint64_t* recv_bytes) {
  EVQ_NCCL(nccl().GroupStart());
  for (int r = 0; r < ctx->nranks; ++r) {
    if (send_bytes[r]) EVQ_NCCL(nccl().Send((const uint8_t*) send + send_off[r], send_bytes[r], ncclUint8, r, (ncclComm_t) ctx->nccl_comm, ctx->stream));
    if (recv_bytes[r]) EVQ_NCCL(nccl().Recv((uint8_t*) recv + recv_off[r], recv_bytes[r], ncclUint8, r, (ncclComm_t) ctx->nccl_comm, ctx->stream));
  }
  EVQ_NCCL(nccl().GroupEnd());
}

}  // namespace evq
