#include "jit.h"
#include <nvrtc.h>
#include <chrono>
#include <functional>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

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

// ---- on-disk cubin cache --------------------------------------------------------------------------------------------------
// A specialised kernel costs NVRTC ~0.3 - 0.5 s; the text is a pure function of (plan, column layout, statistics), so a
// restarted process - or the next one on the same host - finds the cubin under <dir>/<key>.cubin, key = two independent
// 64-bit hashes of the kernel text, its length and the NVRTC version.  The file starts with the text length and both
// hashes again (a truncated or foreign file is ignored and rewritten); it is written to a temporary name and renamed.
// Directory: $EVQGPU_CACHE_DIR, else $XDG_CACHE_HOME/evqgpu, else ~/.cache/evqgpu; EVQGPU_CACHE_DIR="" disables the cache.
namespace {

struct CacheKey { uint64_t h1, h2, len; };

CacheKey cache_key(const std::string& s) {
  int major = 0, minor = 0;
  nvrtcVersion(&major, &minor);
  uint64_t h1 = 1469598103934665603ull ^ (uint64_t) (major * 100 + minor), h2 = 0x9e3779b97f4a7c15ull + (uint64_t) (major * 100 + minor);
  for (unsigned char c : s) {
    h1 = (h1 ^ c) * 1099511628211ull;
    h2 = (h2 + c) * 0xff51afd7ed558ccdull;
    h2 ^= h2 >> 29;
  }
  return CacheKey{h1, h2, (uint64_t) s.size()};
}

std::string cache_dir() {
  if (const char* d = getenv("EVQGPU_CACHE_DIR")) return d;
  if (const char* x = getenv("XDG_CACHE_HOME")) return std::string(x) + "/evqgpu";
  if (const char* h = getenv("HOME")) return std::string(h) + "/.cache/evqgpu";
  return "";
}

std::string cache_path(const CacheKey& k) {
  const std::string dir = cache_dir();
  if (dir.empty()) return "";
  char name[80];
  snprintf(name, sizeof(name), "/%016llx%016llx.cubin", (unsigned long long) k.h1, (unsigned long long) k.h2);
  return dir + name;
}

bool cache_load(const CacheKey& k, std::vector<char>* cubin) {
  const std::string path = cache_path(k);
  if (path.empty()) return false;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  uint64_t hdr[4] = {0, 0, 0, 0};
  bool ok = fread(hdr, 8, 4, f) == 4 && hdr[0] == k.len && hdr[1] == k.h1 && hdr[2] == k.h2 && hdr[3] > 0 && hdr[3] < (1ull << 30);
  if (ok) {
    cubin->resize(hdr[3]);
    ok = fread(cubin->data(), 1, hdr[3], f) == hdr[3] && fgetc(f) == EOF;
  }
  fclose(f);
  return ok;
}

void cache_store(const CacheKey& k, const std::vector<char>& cubin) {
  const std::string path = cache_path(k);
  if (path.empty()) return;
  const std::string dir = cache_dir();
  for (size_t i = 1; i <= dir.size(); ++i)   // mkdir -p
    if (i == dir.size() || dir[i] == '/') mkdir(dir.substr(0, i).c_str(), 0755);
  char tmp[32];
  snprintf(tmp, sizeof(tmp), ".tmp%d", (int) getpid());
  const std::string tpath = path + tmp;
  FILE* f = fopen(tpath.c_str(), "wb");
  if (!f) return;   // a cache that cannot be written is not an error
  const uint64_t hdr[4] = {k.len, k.h1, k.h2, (uint64_t) cubin.size()};
  const bool ok = fwrite(hdr, 8, 4, f) == 4 && fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
  fclose(f);
  if (!ok || rename(tpath.c_str(), path.c_str()) != 0) unlink(tpath.c_str());
}

}  // namespace

std::shared_ptr<JitModule> jit_compile(evqgpu_ctx* ctx, const std::string& source, const std::vector<std::string>& kernels,
                                       float* compile_ms_out) {
  if (compile_ms_out) *compile_ms_out = 0;
  auto it = ctx->jit_cache.find(source);
  if (it != ctx->jit_cache.end()) return it->second;
  auto t0 = std::chrono::steady_clock::now();
  std::string log;
  std::vector<char> cubin;
  const CacheKey key = cache_key(source);
  const bool from_disk = !getenv("EVQGPU_JIT_DUMP_DIR") && cache_load(key, &cubin);
  if (!from_disk) {
    cubin = jit_compile_to_cubin(source, &log);
    cache_store(key, cubin);
  }
  auto m = std::make_shared<JitModule>();
  m->source = source;
  m->from_disk = from_disk;
  use_device(ctx);
  if (from_disk && cudaLibraryLoadData(&m->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) {
    cudaGetLastError();   // a cubin the driver refuses (another toolkit wrote it): compile it here
    m->lib = nullptr;
    m->from_disk = false;
    cubin = jit_compile_to_cubin(source, &log);
    cache_store(key, cubin);
  }
  if (!m->lib) EVQ_CUDA(cudaLibraryLoadData(&m->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
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
