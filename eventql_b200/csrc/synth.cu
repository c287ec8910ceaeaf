// synth.cu - synthetic cstable tables generated directly in HBM (BASELINE.json's configs).
// Values are a pure function of (seed, row): r = splitmix64(seed + row); v = lo + r % span [optionally mixed
// again], so any slice of a table can be regenerated anywhere (tests regenerate them with numpy).
#include "table.h"
#include <cub/device/device_scan.cuh>
#include <string.h>

namespace evq {

__host__ __device__ __forceinline__ u64 splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ull;
  u64 z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct SynthSpec {
  u64 seed, lo, span, row_offset;
  u32 transform, null_every;
};

__device__ __forceinline__ bool synth_is_null(const SynthSpec& s, u64 row) {
  return s.null_every && ((row + s.row_offset) % s.null_every) == s.null_every - 1;
}

__device__ __forceinline__ u64 synth_value(const SynthSpec& s, u64 row) {
  const u64 r = splitmix64(s.seed + row + s.row_offset);
  u64 v = s.lo + r % s.span;
  if (s.transform == 1) v = splitmix64(v);
  else if (s.transform == 2) v = (u64) __double_as_longlong((double) v / 100.0);
  return v;
}

__device__ __forceinline__ u32 leb_len(u64 v) {
  u32 n = 1;
  while (v >>= 7) ++n;
  return n;
}

// ---- dlevel stream: bit-packed b=1, libsimdcomp vertical layout; one thread per 32-bit word
__global__ void k_synth_levels(SynthSpec s, u64 num_rows, u32* __restrict__ out, u64 nwords) {
  const u64 w = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  const u64 blk = w >> 2;
  const u32 lane = w & 3u;
  u32 word = 0;
  for (u32 j = 0; j < 32; ++j) {
    const u64 row = blk * 128 + 4 * j + lane;
    if (row < num_rows && !synth_is_null(s, row)) word |= 1u << j;
  }
  out[w] = word;
}

// NULL rows are periodic, so the index of a row among the non-NULL values has a closed form
__host__ __device__ __forceinline__ u64 synth_nulls_before(const SynthSpec& s, u64 row) {
  if (!s.null_every) return 0;
  return (row + s.row_offset) / s.null_every - s.row_offset / s.null_every;
}

// ---- fixed width encodings: one thread per row
__global__ void k_synth_plain(SynthSpec s, u64 num_rows, u32 width, u8* __restrict__ out) {
  const u64 row = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= num_rows) return;
  if (synth_is_null(s, row)) return;
  const u64 idx = row - synth_nulls_before(s, row);   // index among the non-NULL values
  const u64 v = synth_value(s, row);
  if (width == 8) ((u64*) out)[idx] = v;
  else ((u32*) out)[idx] = (u32) v;
}

// ---- bit-packed data (vertical layout), width b; one thread per output word
__global__ void k_synth_bitpack(SynthSpec s, u64 num_values, u32 b, u32* __restrict__ out, u64 nwords) {
  // only used for required columns (value index == row)
  const u64 w = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  const u64 blk = w / (4ull * b);
  const u32 rem = (u32) (w % (4ull * b));
  const u32 lane = rem & 3u, wi = rem >> 2;   // word wi of lane's private stream
  u32 word = 0;
  // values j with bit range [j*b, j*b+b) intersecting [32*wi, 32*wi+32)
  const u32 j0 = (32u * wi) / b;
  for (u32 j = j0; j < 32 && j * b < 32u * (wi + 1); ++j) {
    const u64 row = blk * 128 + 4ull * j + lane;
    u64 v = row < num_values ? synth_value(s, row) : 0;
    if (b < 32) v &= (1ull << b) - 1;
    else v &= 0xffffffffull;
    const int shift = (int) (j * b) - (int) (32u * wi);
    if (shift >= 0) word |= (u32) (v << shift);
    else word |= (u32) (v >> (-shift));
  }
  out[w] = word;
}

// ---- LEB128: bytes per 1024-row group, exclusive scan, then every thread writes its group's bytes
__global__ void k_synth_leb_sizes(SynthSpec s, u64 num_rows, u64 ngroups, u64* __restrict__ sizes) {
  const u64 warp = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  if (warp >= ngroups) return;
  const u64 r0 = warp * 1024, r1 = min(num_rows, r0 + 1024);
  u32 n = 0;
  for (u64 r = r0 + lane; r < r1; r += 32)
    if (!synth_is_null(s, r)) n += leb_len(synth_value(s, r));
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if (lane == 0) sizes[warp] = n;
}

__global__ void k_synth_leb_write(SynthSpec s, u64 num_rows, u64 ngroups, const u64* __restrict__ offsets, u8* __restrict__ out) {
  // one warp per 1024-row group; lane handles 32 consecutive rows, offsets inside the group via a warp scan
  const u64 warp = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  if (warp >= ngroups) return;
  const u64 r0 = warp * 1024 + (u64) lane * 32;
  const u64 r1 = min(num_rows, r0 + 32);
  u32 n = 0;
  for (u64 r = r0; r < r1; ++r)
    if (!synth_is_null(s, r)) n += leb_len(synth_value(s, r));
  u32 incl = n;
  for (int o = 1; o < 32; o <<= 1) {
    const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (u32) o) incl += t;
  }
  u8* p = out + offsets[warp] + (incl - n);
  for (u64 r = r0; r < r1; ++r) {
    if (synth_is_null(s, r)) continue;
    u64 v = synth_value(s, r);
    do {
      u8 b = v & 0x7f;
      v >>= 7;
      if (v) b |= 0x80;
      *p++ = b;
    } while (v);
  }
}

static void scan_u64(evqgpu_ctx* ctx, u64* in, u64* out, uint64_t n) {
  size_t tmp_bytes = 0;
  EVQ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int64_t) n, ctx->stream));
  DevBuf tmp;
  tmp.alloc(tmp_bytes);
  EVQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, (int64_t) n, ctx->stream));
  EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
}

static void synth_column(evqgpu_table* t, Column& c, const evqgpu_synth_column& spec, uint64_t row_offset) {
  evqgpu_ctx* ctx = t->ctx;
  const uint64_t n = t->num_rows;
  SynthSpec s{spec.seed, spec.lo, spec.span ? spec.span : 1, row_offset, spec.transform, spec.null_every};
  const uint64_t ngroups = (n + 1023) / 1024;
  uint64_t nvalues = n;
  if (spec.null_every) {
    // dlevel stream
    const uint64_t nwords = (n + 127) / 128 * 4;
    c.dlevel.buf.alloc(round_up(nwords * 4, 256) + 256);
    EVQ_CUDA(cudaMemsetAsync(c.dlevel.buf.p, 0, c.dlevel.buf.bytes, ctx->stream));
    if (nwords) k_synth_levels<<<(unsigned) ((nwords + 255) / 256), 256, 0, ctx->stream>>>(s, n, c.dlevel.buf.as<u32>(), nwords);
    c.dlevel.nbytes = nwords * 4;
    c.dlevel.bitpack_max = 1;
    c.dlevel.present = true;
    nvalues = n - synth_nulls_before(s, n);
  }
  switch (c.data_kind) {
    case EVQ_KIND_PLAIN64:
    case EVQ_KIND_PLAIN32: {
      const uint32_t w = c.data_kind == EVQ_KIND_PLAIN64 ? 8 : 4;
      c.data.buf.alloc(round_up(nvalues * w, 256) + 256);
      EVQ_CUDA(cudaMemsetAsync(c.data.buf.p, 0, c.data.buf.bytes, ctx->stream));
      if (n) k_synth_plain<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(s, n, w, c.data.buf.as<u8>());
      c.data.nbytes = nvalues * w;
      break;
    }
    case EVQ_KIND_BITPACK: {
      if (spec.null_every) fail(EVQGPU_ERR_UNSUPPORTED, "synthetic bit-packed columns must be required");
      const uint32_t maxv = c.meta.encoding == EVQ_ENC_BOOLEAN_BITPACKED ? 1u : 0xffffffffu;   // column_writer_uint.cc:57-63
      const uint32_t b = bits_needed(maxv);
      const uint64_t nwords = (n + 127) / 128 * 4 * b;
      c.data.buf.alloc(round_up(nwords * 4, 256) + 256);
      EVQ_CUDA(cudaMemsetAsync(c.data.buf.p, 0, c.data.buf.bytes, ctx->stream));
      if (nwords) k_synth_bitpack<<<(unsigned) ((nwords + 255) / 256), 256, 0, ctx->stream>>>(s, n, b, c.data.buf.as<u32>(), nwords);
      c.data.nbytes = nwords * 4;
      c.data.bitpack_max = maxv;
      break;
    }
    case EVQ_KIND_LEB128: {
      DevBuf sizes, offs;
      sizes.alloc((ngroups + 1) * 8);
      offs.alloc((ngroups + 1) * 8);
      EVQ_CUDA(cudaMemsetAsync(sizes.p, 0, sizes.bytes, ctx->stream));
      if (ngroups) k_synth_leb_sizes<<<(unsigned) ((ngroups * 32 + 255) / 256), 256, 0, ctx->stream>>>(s, n, ngroups, sizes.as<u64>());
      scan_u64(ctx, sizes.as<u64>(), offs.as<u64>(), ngroups + 1);
      uint64_t total = 0;
      EVQ_CUDA(cudaMemcpy(&total, offs.as<u64>() + ngroups, 8, cudaMemcpyDeviceToHost));
      c.data.buf.alloc(round_up(total, 256) + 256);
      EVQ_CUDA(cudaMemsetAsync(c.data.buf.p, 0, c.data.buf.bytes, ctx->stream));
      if (ngroups) k_synth_leb_write<<<(unsigned) ((ngroups * 32 + 255) / 256), 256, 0, ctx->stream>>>(s, n, ngroups, offs.as<u64>(), c.data.buf.as<u8>());
      c.data.nbytes = total;
      EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
      break;
    }
  }
  EVQ_CUDA(cudaGetLastError());
  c.data.present = true;
  table_finish_column(t, c);
}

}  // namespace evq

using namespace evq;

extern "C" int evqgpu_table_synthesize(evqgpu_ctx* ctx, uint64_t num_rows, uint64_t row_offset, const evqgpu_synth_column* cols,
                                       uint32_t ncols, evqgpu_table** out) {
  return guarded([&] {
    if (!ctx || !cols || !out) fail(EVQGPU_ERR_ARG, "evqgpu_table_synthesize: null argument");
    use_device(ctx);
    std::unique_ptr<evqgpu_table> t(new evqgpu_table());
    t->ctx = ctx;
    t->num_rows = num_rows;
    t->meta.version = 2;
    t->meta.num_rows = num_rows;
    for (uint32_t i = 0; i < ncols; ++i) {
      Column c;
      c.meta.name = cols[i].name;
      c.meta.column_id = i + 1;
      c.meta.logical_type = cols[i].logical_type;
      c.meta.encoding = cols[i].encoding;
      c.meta.dlevel_max = cols[i].null_every ? 1 : 0;
      t->cols.push_back(std::move(c));
    }
    table_init_columns(t.get());
    for (uint32_t i = 0; i < ncols; ++i) {
      if (!t->cols[i].scannable) fail(EVQGPU_ERR_UNSUPPORTED, "synthetic column '%s': unsupported type/encoding", cols[i].name);
      synth_column(t.get(), t->cols[i], cols[i], row_offset);
    }
    *out = t.release();
  });
}
