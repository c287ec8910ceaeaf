// util.h - error plumbing shared by the library's translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/evqgpu.h"

namespace evq {

// thread-local message behind evqgpu_last_error()
void set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
const char* last_error();

struct Error {
  int status;
  std::string msg;
};

[[noreturn]] void fail(int status, const char* fmt, ...) __attribute__((format(printf, 2, 3)));

#define EVQ_CUDA(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      ::evq::fail(EVQGPU_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// run `body`, translate exceptions into a status code + thread-local message (nothing throws across the C ABI)
template <typename F>
int guarded(F&& body) {
  try {
    body();
    return EVQGPU_OK;
  } catch (const Error& e) {
    set_error("%s", e.msg.c_str());
    return e.status;
  } catch (const std::bad_alloc&) {
    set_error("out of host memory");
    return EVQGPU_ERR_NOMEM;
  } catch (const std::exception& e) {
    set_error("%s", e.what());
    return EVQGPU_ERR_RUNTIME;
  }
}

inline uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

}  // namespace evq
