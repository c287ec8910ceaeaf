// context.h - evqgpu_ctx: one CUDA device, one stream, the kernel cache and (optionally) an NCCL communicator.
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "util.h"

namespace evq {
struct JitModule;
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
  // scratch for small device->host reads
  void* pinned_scratch = nullptr;   // 4 KiB pinned
  uint64_t kernel_launches = 0;     // total kernels launched through this context
  bool profiling = false;           // bracket scan kernel launches with events (evqgpu_ctx_set_profiling)
};

namespace evq {

inline void use_device(const evqgpu_ctx* ctx) { EVQ_CUDA(cudaSetDevice(ctx->device)); }

// device allocation that frees itself
struct DevBuf {
  void* p = nullptr;
  uint64_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(uint64_t n) {
    release();
    if (n == 0) n = 16;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) { p = nullptr; fail(EVQGPU_ERR_NOMEM, "cudaMalloc(%llu) failed: %s", (unsigned long long) n, cudaGetErrorString(e)); }
    bytes = n;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename T> T* as() const { return (T*) p; }
};

}  // namespace evq
