// jit.h - NVRTC compilation of the per-query kernel text for sm_100a and loading through the CUDA runtime's
// library API.  No tracing compiler, no Triton: the text is hand-written CUDA (kernels/*.cuh) with the
// query's expression programs spelled out as C by csrc/codegen.
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "context.h"

namespace evq {

struct JitModule {
  cudaLibrary_t lib = nullptr;
  std::map<std::string, cudaKernel_t> kernels;
  std::string source;
  float compile_ms = 0;
  bool from_disk = false;   // the cubin came from the on-disk cache (jit.cc), NVRTC did not run
  ~JitModule();
};

// Compile `source` (cached per context by source text) and resolve the named extern "C" kernels.
std::shared_ptr<JitModule> jit_compile(evqgpu_ctx* ctx, const std::string& source, const std::vector<std::string>& kernels,
                                       float* compile_ms_out);

// compile only (no device needed): returns the cubin; used by the build check
std::vector<char> jit_compile_to_cubin(const std::string& source, std::string* log);

// the fixed device text (kernels/evq_abi.h, evq_prelude.cuh, evq_scan_kernel.cuh), embedded at build time
extern const char* const kSrcAbi;
extern const char* const kSrcPrelude;
extern const char* const kSrcScanKernel;
extern const char* const kSrcScanFast;

}  // namespace evq
