#include "util.h"
#include <stdexcept>
#include <vector>

namespace evq {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

const char* last_error() { return g_last_error.c_str(); }

void fail(int status, const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Error{status, buf};
}

}  // namespace evq

// ---- device memory pool (context.h) ----------------------------------------------------------------------------------
#include <map>
#include <mutex>
#include "context.h"

namespace evq {

namespace {
struct Pool {
  std::multimap<uint64_t, void*> free_blocks;   // size -> block
  uint64_t held = 0;
};
std::mutex g_pool_mu;
std::map<int, Pool> g_pools;
const uint64_t kPoolLimit = 24ull << 30;

uint64_t pool_round(uint64_t n) {
  if (n <= (1u << 20)) return (n + 511) & ~511ull;
  return (n + (2u << 20) - 1) & ~((2ull << 20) - 1);   // 2 MiB granules for large blocks
}
}  // namespace

void* pool_alloc(uint64_t bytes, uint64_t* granted) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t want = pool_round(bytes);
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    Pool& P = g_pools[dev];
    auto it = P.free_blocks.lower_bound(want);
    if (it != P.free_blocks.end() && it->first <= want + want / 4) {
      void* p = it->second;
      *granted = it->first;
      P.held -= it->first;
      P.free_blocks.erase(it);
      return p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {   // give the cached blocks back and try once more
    cudaGetLastError();
    pool_trim(dev);
    e = cudaMalloc(&p, want);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    fail(EVQGPU_ERR_NOMEM, "cudaMalloc(%llu) failed: %s", (unsigned long long) want, cudaGetErrorString(e));
  }
  *granted = want;
  return p;
}

void pool_free(void* p, uint64_t granted) {
  int dev = 0;
  cudaGetDevice(&dev);
  bool trim = false;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    Pool& P = g_pools[dev];
    P.free_blocks.emplace(granted, p);
    P.held += granted;
    trim = P.held > kPoolLimit;
  }
  if (trim) pool_trim(dev);
}

void pool_trim(int device) {
  std::multimap<uint64_t, void*> blocks;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    Pool& P = g_pools[device];
    blocks.swap(P.free_blocks);
    P.held = 0;
  }
  for (auto& b : blocks) cudaFree(b.second);
}

}  // namespace evq
