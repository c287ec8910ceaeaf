#include "jit.h"
#include <nvrtc.h>
#include <chrono>
#include <functional>
#include <stdio.h>
#include <stdlib.h>

namespace evq {

JitModule::~JitModule() {
  if (lib) cudaLibraryUnload(lib);
}

std::vector<char> jit_compile_to_cubin(const std::string& source, std::string* log) {
  nvrtcProgram prog;
  // EVQGPU_JIT_DUMP_DIR=<dir>: keep the specialised kernel text as a file and compile it under that name, so that
  // profilers (ncu --import-source on) and cuda-gdb can show the source lines behind the SASS
  std::string name = "evq_query.cu";
  if (const char* dir = getenv("EVQGPU_JIT_DUMP_DIR")) {
    char buf[64];
    snprintf(buf, sizeof(buf), "/evq_query_%016zx.cu", std::hash<std::string>()(source));
    name = std::string(dir) + buf;
    if (FILE* f = fopen(name.c_str(), "w")) {
      fwrite(source.data(), 1, source.size(), f);
      fclose(f);
    }
  }
  if (nvrtcCreateProgram(&prog, source.c_str(), name.c_str(), 0, nullptr, nullptr) != NVRTC_SUCCESS)
    fail(EVQGPU_ERR_CUDA, "nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--extra-device-vectorization"};
  nvrtcResult rc = nvrtcCompileProgram(prog, 4, opts);
  size_t log_size = 0;
  nvrtcGetProgramLogSize(prog, &log_size);
  std::string l(log_size, '\0');
  if (log_size > 1) nvrtcGetProgramLog(prog, &l[0]);
  if (log) *log = l;
  if (rc != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    fail(EVQGPU_ERR_CUDA, "NVRTC compilation failed: %s\n%s", nvrtcGetErrorString(rc), l.c_str());
  }
  size_t n = 0;
  if (nvrtcGetCUBINSize(prog, &n) != NVRTC_SUCCESS || n == 0) {
    nvrtcDestroyProgram(&prog);
    fail(EVQGPU_ERR_CUDA, "NVRTC produced no cubin");
  }
  std::vector<char> cubin(n);
  nvrtcGetCUBIN(prog, cubin.data());
  nvrtcDestroyProgram(&prog);
  return cubin;
}

std::shared_ptr<JitModule> jit_compile(evqgpu_ctx* ctx, const std::string& source, const std::vector<std::string>& kernels,
                                       float* compile_ms_out) {
  if (compile_ms_out) *compile_ms_out = 0;
  auto it = ctx->jit_cache.find(source);
  if (it != ctx->jit_cache.end()) return it->second;
  auto t0 = std::chrono::steady_clock::now();
  std::string log;
  std::vector<char> cubin = jit_compile_to_cubin(source, &log);
  auto m = std::make_shared<JitModule>();
  m->source = source;
  use_device(ctx);
  EVQ_CUDA(cudaLibraryLoadData(&m->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  for (const auto& k : kernels) {
    cudaKernel_t kern;
    EVQ_CUDA(cudaLibraryGetKernel(&kern, m->lib, k.c_str()));
    m->kernels[k] = kern;
  }
  m->compile_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (compile_ms_out) *compile_ms_out = m->compile_ms;
  ctx->jit_cache[source] = m;
  return m;
}

}  // namespace evq
