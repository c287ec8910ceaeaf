// api.cc - context management and the small entry points of the C ABI (include/evqgpu.h).
#include <string.h>
#include "context.h"
#include "query.h"
#include "table.h"

using namespace evq;

extern "C" {

int evqgpu_abi_version(void) { return EVQGPU_ABI_VERSION; }

const char* evqgpu_last_error(void) { return evq::last_error(); }

int evqgpu_ctx_create(int device, uint64_t flags, evqgpu_ctx** out) {
  return guarded([&] {
    if (!out) fail(EVQGPU_ERR_ARG, "evqgpu_ctx_create: null argument");
    (void) flags;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      fail(EVQGPU_ERR_CUDA, "no CUDA device available (%s): this engine has no host execution mode",
           e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) fail(EVQGPU_ERR_ARG, "device %d out of range (0..%d)", device, count - 1);
    std::unique_ptr<evqgpu_ctx> ctx(new evqgpu_ctx());
    ctx->device = device;
    EVQ_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    EVQ_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int) prop.sharedMemPerBlockOptin;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    if (prop.major < 10)
      fail(EVQGPU_ERR_UNSUPPORTED, "device %d is sm_%d%d; the kernels are written for sm_100a (B200)", device, prop.major, prop.minor);
    EVQ_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->pool = pool_create();
    EVQ_CUDA(cudaMallocHost(&ctx->pinned_scratch, 4096));
    *out = ctx.release();
  });
}

int evqgpu_comm_destroy(evqgpu_ctx* ctx);

void evqgpu_ctx_destroy(evqgpu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  evqgpu_comm_destroy(ctx);
  ctx->jit_cache.clear();
  evq::pool_trim(ctx->pool);   // blocks still held by live tables / queries return to the (shared) pool object and go with it
  evq::set_current_ctx(nullptr);
  if (ctx->pinned_scratch) cudaFreeHost(ctx->pinned_scratch);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int evqgpu_host_alloc(evqgpu_ctx* ctx, uint64_t nbytes, void** out) {
  return guarded([&] {
    if (!ctx || !out) fail(EVQGPU_ERR_ARG, "evqgpu_host_alloc: null argument");
    use_device(ctx);
    EVQ_CUDA(cudaMallocHost(out, nbytes ? nbytes : 1));
  });
}

int evqgpu_host_free(evqgpu_ctx* ctx, void* ptr) {
  return guarded([&] {
    if (!ctx) fail(EVQGPU_ERR_ARG, "evqgpu_host_free: null context");
    use_device(ctx);
    if (ptr) EVQ_CUDA(cudaFreeHost(ptr));
  });
}

int evqgpu_host_register(evqgpu_ctx* ctx, void* ptr, uint64_t nbytes) {
  return guarded([&] {
    if (!ctx || !ptr) fail(EVQGPU_ERR_ARG, "evqgpu_host_register: null argument");
    use_device(ctx);
    EVQ_CUDA(cudaHostRegister(ptr, nbytes, cudaHostRegisterDefault));
  });
}

int evqgpu_host_unregister(evqgpu_ctx* ctx, void* ptr) {
  return guarded([&] {
    if (!ctx || !ptr) fail(EVQGPU_ERR_ARG, "evqgpu_host_unregister: null argument");
    use_device(ctx);
    EVQ_CUDA(cudaHostUnregister(ptr));
  });
}

int evqgpu_ctx_set_profiling(evqgpu_ctx* ctx, int on) {
  return guarded([&] {
    if (!ctx) fail(EVQGPU_ERR_ARG, "evqgpu_ctx_set_profiling: null context");
    ctx->profiling = on != 0;
  });
}

uint64_t evqgpu_ctx_kernel_launches(const evqgpu_ctx* ctx) { return ctx ? ctx->kernel_launches : 0; }

void* evqgpu_ctx_stream(evqgpu_ctx* ctx) { return ctx ? (void*) ctx->stream : nullptr; }

int evqgpu_ctx_synchronize(evqgpu_ctx* ctx) {
  return guarded([&] {
    if (!ctx) fail(EVQGPU_ERR_ARG, "evqgpu_ctx_synchronize: null context");
    use_device(ctx);
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

// FastCSTableScan::fetchColumn* on its own: `select <column> from t` through the scan-only path
int evqgpu_table_decode_column(evqgpu_table* tbl, const char* column, uint64_t row0, uint64_t nrows, void* dst, uint64_t cap) {
  return guarded([&] {
    if (!tbl || !column || !dst) fail(EVQGPU_ERR_ARG, "evqgpu_table_decode_column: null argument");
    const int ci = tbl->find(column);
    if (ci < 0) fail(EVQGPU_ERR_ARG, "column not found: %s", column);
    const uint32_t ty = tbl->cols[ci].sql_type;
    const uint64_t w = ty == EVQ_BOOL ? 2 : 9;
    if (row0 > tbl->num_rows) row0 = tbl->num_rows;
    nrows = std::min<uint64_t>(nrows, tbl->num_rows - row0);
    if (cap < nrows * w) fail(EVQGPU_ERR_ARG, "evqgpu_table_decode_column: buffer too small");
    evqgpu_insn in;
    memset(&in, 0, sizeof(in));
    in.op = EVQ_X_INPUT;
    in.type = (uint8_t) ty;
    evqgpu_expr sel = {&in, 1, nullptr, 0};
    const char* names[1] = {column};
    evqgpu_query_desc d;
    memset(&d, 0, sizeof(d));
    d.struct_size = sizeof(d);
    d.num_input_columns = 1;
    d.input_columns = names;
    d.num_select = 1;
    d.select = &sel;
    evqgpu_query* q = nullptr;
    int rc = evqgpu_query_create(tbl->ctx, &d, &q);
    if (rc != EVQGPU_OK) throw Error{rc, last_error()};
    std::unique_ptr<evqgpu_query, void (*)(evqgpu_query*)> guard(q, evqgpu_query_destroy);
    rc = evqgpu_query_execute(q, &tbl, 1);
    if (rc != EVQGPU_OK) throw Error{rc, last_error()};
    void* cols[1] = {dst};
    uint64_t got = 0;
    rc = evqgpu_query_fetch(q, row0, nrows, cols, &got);
    if (rc != EVQGPU_OK) throw Error{rc, last_error()};
    if (got != nrows) fail(EVQGPU_ERR_RUNTIME, "decode returned %llu of %llu rows", (unsigned long long) got, (unsigned long long) nrows);
  });
}

}  // extern "C"
