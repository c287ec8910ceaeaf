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

struct Pool {
  std::mutex mu;
  std::multimap<uint64_t, void*> free_blocks;   // size -> block
  uint64_t held = 0;
  int device = 0;
  ~Pool() {
    for (auto& b : free_blocks) cudaFree(b.second);
  }
};

namespace {
const uint64_t kPoolLimit = 24ull << 30;
thread_local const evqgpu_ctx* g_current_ctx = nullptr;

uint64_t pool_round(uint64_t n) {
  if (n <= (1u << 20)) return (n + 511) & ~511ull;
  return (n + (2u << 20) - 1) & ~((2ull << 20) - 1);   // 2 MiB granules for large blocks
}
}  // namespace

void set_current_ctx(const evqgpu_ctx* ctx) { g_current_ctx = ctx; }

std::shared_ptr<Pool> pool_create() {
  auto p = std::make_shared<Pool>();
  cudaGetDevice(&p->device);
  return p;
}

void* pool_alloc(uint64_t bytes, uint64_t* granted, std::shared_ptr<Pool>* owner) {
  const uint64_t want = pool_round(bytes);
  std::shared_ptr<Pool> P = g_current_ctx ? g_current_ctx->pool : nullptr;
  *owner = P;
  if (P) {
    std::lock_guard<std::mutex> lk(P->mu);
    auto it = P->free_blocks.lower_bound(want);
    if (it != P->free_blocks.end() && it->first <= want + want / 4) {
      void* p = it->second;
      *granted = it->first;
      P->held -= it->first;
      P->free_blocks.erase(it);
      return p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess && P) {   // give the cached blocks back and try once more
    cudaGetLastError();
    pool_trim(P);
    e = cudaMalloc(&p, want);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    fail(EVQGPU_ERR_NOMEM, "cudaMalloc(%llu) failed: %s", (unsigned long long) want, cudaGetErrorString(e));
  }
  *granted = want;
  return p;
}

void pool_free(const std::shared_ptr<Pool>& owner, void* p, uint64_t granted) {
  if (!owner) {
    cudaFree(p);
    return;
  }
  bool trim = false;
  {
    std::lock_guard<std::mutex> lk(owner->mu);
    owner->free_blocks.emplace(granted, p);
    owner->held += granted;
    trim = owner->held > kPoolLimit;
  }
  if (trim) pool_trim(owner);
}

void pool_trim(const std::shared_ptr<Pool>& pool) {
  if (!pool) return;
  std::multimap<uint64_t, void*> blocks;
  {
    std::lock_guard<std::mutex> lk(pool->mu);
    blocks.swap(pool->free_blocks);
    pool->held = 0;
  }
  for (auto& b : blocks) cudaFree(b.second);   // (cudaFree synchronises the device: nothing still reads the blocks)
}

}  // namespace evq
