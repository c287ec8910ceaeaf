// order.cu - ORDER BY / LIMIT over the result rows of an executed query, on the device.
//
//   evqgpu_query_order_by   csql::OrderByExpression   sql/statements/select/orderby.cc:58-160
//   evqgpu_query_limit      csql::LimitExpression     sql/statements/select/limit.cc:43-112
//
// The reference collects its input rows, std::sort()s them with one typed `cmp` call per sort spec (values only, NULL
// tags ignored like everywhere in pure functions) and streams them out.  Here the result columns already sit
// in HBM in the packed SVector encoding (9 B numeric, 2 B BOOL): per sort spec, last first, the column is turned into
// order-preserving 64-bit keys (sign flip for INT64, the IEEE-754 total-order transform for FLOAT64 with -0.0 == +0.0,
// complemented for DESC), cub::DeviceRadixSort::SortPairs - stable - permutes a row index, and one gather pass rewrites
// every column.  Rows the reference may emit in any order (equal sort keys; std::sort is not stable) come out in their
// previous order.
#include <cub/cub.cuh>

#include "query.h"

namespace evq {
void finish_query(evqgpu_query& q);

__device__ __forceinline__ u64 load_packed_u64(const u8* p) {
  u64 v = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) v |= (u64) p[i] << (8 * i);
  return v;
}

// keys[i] = order-preserving key of row perm[i] of one packed result column
__global__ void k_sort_keys(const u8* __restrict__ col, u32 elem, u32 type, u32 descending, const u32* __restrict__ perm, u64 n,
                            u64* __restrict__ keys) {
  const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u8* p = col + (u64) perm[i] * elem;
  u64 k;
  if (elem == 2) {
    k = p[0] ? 1ull : 0ull;
  } else {
    const u64 v = load_packed_u64(p);
    switch (type) {
      case EVQ_INT64: k = v ^ 0x8000000000000000ull; break;
      case EVQ_FLOAT64: {
        const u64 b = v == 0x8000000000000000ull ? 0ull : v;   // -0.0 and +0.0 compare equal
        k = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
        break;
      }
      default: k = v; break;   // UINT64, TIMESTAMP64
    }
  }
  keys[i] = descending ? ~k : k;
}

__global__ void k_iota(u32* __restrict__ p, u64 n) {
  const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (u32) i;
}

// dst row i = src row perm[first + i]
__global__ void k_gather_rows(const u8* __restrict__ src, u8* __restrict__ dst, u32 elem, const u32* __restrict__ perm, u64 first,
                              u64 n) {
  const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u8* s = src + (u64) (perm ? perm[first + i] : first + i) * elem;
  u8* d = dst + i * elem;
  for (u32 b = 0; b < elem; ++b) d[b] = s[b];
}

static uint32_t elem_size(const evqgpu_query& q, size_t col) { return q.select[col].expr->type == EVQ_BOOL ? 2 : 9; }

// rewrite every result column as rows perm[first .. first + n) (perm == nullptr: the identity)
static void rewrite_rows(evqgpu_query& q, const u32* perm, uint64_t first, uint64_t n) {
  evqgpu_ctx* ctx = q.ctx;
  for (size_t c = 0; c < q.select.size(); ++c) {
    const uint32_t w = elem_size(q, c);
    DevBuf out;
    out.alloc(n * w + 16);
    if (n) {
      k_gather_rows<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(q.out_cols[c].as<u8>(), out.as<u8>(), w, perm, first, n);
      EVQ_CUDA(cudaGetLastError());
      ctx->kernel_launches++;
      q.stats.kernel_launches++;
    }
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));   // (the old buffer goes back to the pool)
    q.out_cols[c] = std::move(out);
  }
  q.num_rows_out = n;
  q.out_capacity = n;
  q.reordered = true;
}

}  // namespace evq

using namespace evq;

extern "C" {

int evqgpu_query_order_by(evqgpu_query* q, const evqgpu_sort_spec* specs, uint32_t nspecs) {
  return guarded([&] {
    if (!q || !specs) fail(EVQGPU_ERR_ARG, "evqgpu_query_order_by: null argument");
    if (nspecs == 0) fail(EVQGPU_ERR_ARG, "can't execute ORDER BY: no sort specs");   // orderby.cc:52-54
    if (q->pending) finish_query(*q);
    for (uint32_t i = 0; i < nspecs; ++i)
      if (specs[i].column >= q->select.size())
        fail(EVQGPU_ERR_ARG, "evqgpu_query_order_by: sort column %u of %zu", specs[i].column, q->select.size());
    for (uint32_t i = 0; i < nspecs; ++i)
      if (q->select[specs[i].column].is_string)   // the device column holds dictionary codes, which carry no order
        fail(EVQGPU_ERR_UNSUPPORTED, "evqgpu_query_order_by: ORDER BY a string column");
    const uint64_t n = q->num_rows_out;
    if (n >= (1ull << 32)) fail(EVQGPU_ERR_UNSUPPORTED, "ORDER BY over more than 2^32 result rows");
    if (n < 2) return;
    evqgpu_ctx* ctx = q->ctx;
    use_device(ctx);
    DevBuf perm_a, perm_b, keys_a, keys_b, tmp;
    perm_a.alloc(n * 4);
    perm_b.alloc(n * 4);
    keys_a.alloc(n * 8);
    keys_b.alloc(n * 8);
    const unsigned blocks = (unsigned) ((n + 255) / 256);
    k_iota<<<blocks, 256, 0, ctx->stream>>>(perm_a.as<u32>(), n);
    EVQ_CUDA(cudaGetLastError());
    size_t tmp_bytes = 0;
    EVQ_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_a.as<u64>(), keys_b.as<u64>(), perm_a.as<u32>(), perm_b.as<u32>(),
                                             (int64_t) n, 0, 64, ctx->stream));
    tmp.alloc(tmp_bytes);
    u32* cur = perm_a.as<u32>();
    u32* nxt = perm_b.as<u32>();
    for (uint32_t s = nspecs; s-- > 0;) {   // least significant sort spec first: the sort is stable
      const uint32_t c = specs[s].column;
      k_sort_keys<<<blocks, 256, 0, ctx->stream>>>(q->out_cols[c].as<u8>(), elem_size(*q, c), (u32) q->select[c].expr->type,
                                                    specs[s].descending ? 1u : 0u, cur, n, keys_a.as<u64>());
      EVQ_CUDA(cudaGetLastError());
      EVQ_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_a.as<u64>(), keys_b.as<u64>(), cur, nxt, (int64_t) n, 0, 64,
                                               ctx->stream));
      std::swap(cur, nxt);
      ctx->kernel_launches += 2;
      q->stats.kernel_launches += 2;
    }
    rewrite_rows(*q, cur, 0, n);
  });
}

int evqgpu_query_limit(evqgpu_query* q, uint64_t limit, uint64_t offset) {
  return guarded([&] {
    if (!q) fail(EVQGPU_ERR_ARG, "evqgpu_query_limit: null argument");
    if (q->pending) finish_query(*q);
    use_device(q->ctx);
    const uint64_t n = q->num_rows_out;
    const uint64_t first = std::min(offset, n);
    const uint64_t keep = std::min(limit, n - first);
    if (first == 0 && keep == n) return;
    if (first == 0) {   // a prefix: nothing moves
      q->num_rows_out = keep;
      return;
    }
    rewrite_rows(*q, nullptr, first, keep);
  });
}

}  // extern "C"
