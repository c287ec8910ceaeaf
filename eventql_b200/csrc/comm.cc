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

namespace {

// Peer-mapped exchange buffers for the fused merge tail (query.cu launch_tail, codegen.cc evq_tail): every rank exports
// one small cudaMalloc'ed buffer with CUDA IPC and maps all peers' buffers (NVLink P2P inside the box).  Whether the
// fused path is used is decided HERE, collectively: every rank reports whether its part worked, and the path is on only
// if it worked everywhere - otherwise all ranks keep the NCCL all-gather merge.
void p2p_setup(evqgpu_ctx* ctx) {
  ctx->p2p_ok = false;
  uint64_t ok = getenv("EVQGPU_NO_P2P") ? 0 : 1;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (ok && ctx->nranks > 16) ok = 0;
  if (ok) {
    if (cudaMalloc(&ctx->p2p_local, EVQ_P2P_BYTES) != cudaSuccess || cudaMemset(ctx->p2p_local, 0, EVQ_P2P_BYTES) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess || cudaIpcGetMemHandle(&mine, ctx->p2p_local) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
    }
  }
  std::vector<uint64_t> msg(9, 0);
  msg[0] = ok;
  memcpy(&msg[1], &mine, 64);
  const std::vector<uint64_t> all = comm_all_gather_host(ctx, msg);   // (also orders every rank's memset before any peer's first write)
  bool all_ok = true;
  for (int r = 0; r < ctx->nranks; ++r) all_ok = all_ok && all[(size_t) r * 9] == 1;
  uint64_t opened = all_ok ? 1 : 0;
  if (all_ok) {
    for (int r = 0; r < ctx->nranks && opened; ++r) {
      if (r == ctx->rank) { ctx->p2p_peer[r] = ctx->p2p_local; continue; }
      cudaIpcMemHandle_t h;
      memcpy(&h, &all[(size_t) r * 9 + 1], 64);
      if (cudaIpcOpenMemHandle(&ctx->p2p_peer[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ctx->p2p_peer[r] = nullptr;
        opened = 0;
      }
    }
  }
  const std::vector<uint64_t> all2 = comm_all_gather_host(ctx, std::vector<uint64_t>{opened});
  bool everywhere = true;
  for (uint64_t o : all2) everywhere = everywhere && o == 1;
  if (!everywhere) {
    for (int r = 0; r < ctx->nranks; ++r)
      if (r != ctx->rank && ctx->p2p_peer[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
    memset(ctx->p2p_peer, 0, sizeof(ctx->p2p_peer));
    if (ctx->p2p_local) cudaFree(ctx->p2p_local);
    ctx->p2p_local = nullptr;
    cudaGetLastError();
    return;
  }
  ctx->p2p_ok = true;
  ctx->p2p_epoch = 0;
}

void p2p_teardown(evqgpu_ctx* ctx) {
  if (!ctx->p2p_local) return;
  cudaDeviceSynchronize();
  for (int r = 0; r < ctx->nranks; ++r)
    if (r != ctx->rank && ctx->p2p_peer[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
  memset(ctx->p2p_peer, 0, sizeof(ctx->p2p_peer));
  cudaFree(ctx->p2p_local);
  ctx->p2p_local = nullptr;
  ctx->p2p_ok = false;
  cudaGetLastError();
}

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
    if (nranks > 1) p2p_setup(ctx);
  });
}

int evqgpu_comm_destroy(evqgpu_ctx* ctx) {
  return guarded([&] {
    if (!ctx || !ctx->nccl_comm) return;
    p2p_teardown(ctx);
    nccl().CommDestroy((ncclComm_t) ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
  });
}

int evqgpu_query_merge_rows(evqgpu_query* q, const void* base, const uint64_t* row_starts, const uint64_t* row_ends, uint64_t nrows) {
  return guarded([&] {
    if (!q || (nrows && (!base || !row_starts || !row_ends))) fail(EVQGPU_ERR_ARG, "evqgpu_query_merge_rows: null argument");
    if (!q->coordinator) fail(EVQGPU_ERR_ARG, "evqgpu_query_merge_rows: the plan was not created with EVQGPU_QUERY_COORDINATOR");
    if (q->state_ops.empty()) {
      evq::KernelShape none;
      evq::layout_states(*q, none);
    }
    evq::coordinator_parse_rows(*q, (const uint8_t*) base, row_starts, row_ends, nrows);
  });
}

int evqgpu_query_merge_finish(evqgpu_query* q) {
  return guarded([&] {
    if (!q) fail(EVQGPU_ERR_ARG, "evqgpu_query_merge_finish: null query");
    if (!q->coordinator) fail(EVQGPU_ERR_ARG, "evqgpu_query_merge_finish: the plan was not created with EVQGPU_QUERY_COORDINATOR");
    evq::coordinator_finish(*q);
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

void comm_all_reduce_sum_u64(evqgpu_ctx* ctx, void* buf, size_t nwords) {
  EVQ_NCCL(nccl().AllReduce(buf, buf, nwords, ncclUint64, ncclSum, (ncclComm_t) ctx->nccl_comm, ctx->stream));
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
