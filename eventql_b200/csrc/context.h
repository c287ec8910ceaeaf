// context.h - evqgpu_ctx: one CUDA device, one stream, the kernel cache and (optionally) an NCCL communicator.
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
#include "util.h"

namespace evq {
struct JitModule;
struct Pool;   // device memory blocks cached for reuse (util.cc)
}

struct evqgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 0;
  int smem_optin = 0;        // max dynamic shared memory per CTA (opt-in)
  int cc_major = 0, cc_minor = 0;
  // JIT cache: generated source -> loaded module
  std::map<std::string, std::shared_ptr<evq::JitModule>> jit_cache;
  // NCCL (loaded lazily with dlopen, see comm.cc)
  void* nccl_comm = nullptr;
  int rank = 0, nranks = 1;
  // peer-to-peer exchange buffers for the fused merge tail of the dense tier (comm.cc: p2p_setup): every rank owns
  // [2 parities][16 ranks][64 KiB] of state slots + one flag word per source rank, mapped into all peers with CUDA IPC
  // over NVLink.  p2p_ok is agreed on by all ranks (all or none); without it the merge takes the NCCL all-gather path.
  bool p2p_ok = false;
  void* p2p_local = nullptr;          // this rank's buffer (cudaMalloc, exported)
  void* p2p_peer[16] = {nullptr};     // the peers' buffers as mapped here (p2p_peer[rank] == p2p_local)
  uint64_t p2p_epoch = 0;             // merges done through the buffers; the same on all ranks (merges are collective)
  // scratch for small device->host reads
  void* pinned_scratch = nullptr;   // 4 KiB pinned
  uint64_t kernel_launches = 0;     // total kernels launched through this context
  bool profiling = false;           // bracket scan kernel launches with events (evqgpu_ctx_set_profiling)
  // string dictionary of the context (strings.cu): every distinct value of a string column a query touches, and every
  // string literal, gets a dense code; code 0 is the empty string (what a NULL string compares as, boolean.cc:235-257)
  std::unordered_map<std::string, uint32_t> string_codes;
  std::vector<std::string> code_strings;
  // multi-rank jobs: the first dict_agreed entries are identical on all ranks (strings.cu sync_dictionary); codes behind
  // them are provisional until the next synchronisation, which renumbers them - in the code columns registered here too
  uint32_t dict_agreed = 0;
  std::vector<std::pair<void*, void*>> code_columns;   // (evqgpu_table*, its string Column*) of every code column built
  // device memory pool of THIS context.  All work of a context is ordered on its one stream, so a block freed by the
  // context and handed out again to the same context is reused in stream order; blocks never move between contexts
  // (two contexts on one device have two streams: a shared pool would hand a block to the other stream while kernels of
  // the first still use it).
  std::shared_ptr<evq::Pool> pool;
};

namespace evq {

// the context whose entry point runs on this thread: device allocations are served from its pool
void set_current_ctx(const evqgpu_ctx* ctx);
inline void use_device(const evqgpu_ctx* ctx) {
  EVQ_CUDA(cudaSetDevice(ctx->device));
  set_current_ctx(ctx);
}

// Device memory goes through a small per-CONTEXT pool: loading a column allocates ~8 buffers (stream, indexes, scan
// scratch) and cudaMalloc / cudaFree of hundreds of MB cost about a millisecond each (cudaFree also synchronises the
// device), which dominated the per-column load time.  Blocks are reused by best fit (within 25 % slack); a pool is
// trimmed when it holds more than 24 GiB and released with its context.  A block returns to the pool it came from.
std::shared_ptr<Pool> pool_create();
void* pool_alloc(uint64_t bytes, uint64_t* granted, std::shared_ptr<Pool>* owner);
void pool_free(const std::shared_ptr<Pool>& owner, void* p, uint64_t granted);
void pool_trim(const std::shared_ptr<Pool>& pool);

// device allocation that frees itself
struct DevBuf {
  void* p = nullptr;
  uint64_t bytes = 0;
  std::shared_ptr<Pool> owner;   // the pool the block came from (null: plain cudaMalloc, no context was current)
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes), owner(std::move(o.owner)) { o.p = nullptr; o.bytes = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; bytes = o.bytes; owner = std::move(o.owner); o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(uint64_t n) {
    release();
    if (n == 0) n = 16;
    uint64_t granted = 0;
    p = pool_alloc(n, &granted, &owner);
    bytes = granted;
  }
  void release() {
    if (p) pool_free(owner, p, bytes);
    p = nullptr;
    bytes = 0;
    owner.reset();
  }
  template <typename T> T* as() const { return (T*) p; }
};

}  // namespace evq
