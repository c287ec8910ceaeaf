#include "jit.h"
#include <nvrtc.h>
#include <chrono>

namespace evq {

JitModule::~JitModule() {
  if (lib) cudaLibraryUnload(lib);
}

std::vector<char> jit_compile_to_cubin(const std::string& source, std::string* log) {
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, source.c_str(), "evq_query.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
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
