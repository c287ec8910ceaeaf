// table.cu - table residency: host->device copies of column pages and the device-side row-tile index.
//
// Replaces CSTableReader::openFile / getColumnReader / PageManager::readPage
// (io/cstable/cstable_reader.cc:78-200, page_manager.cc:125-171): instead of pread()ing 512 KiB pages on
// demand, the pages of the referenced columns are DMA'd once into contiguous HBM streams.  Values are NOT
// decoded here - the query kernels decode the page encodings themselves; the only derived data is the small
// row-tile index that makes row i of column A line up with row i of column B (SURVEY §7 "hard parts").
#include "table.h"
#include <cub/device/device_scan.cuh>
#include <string.h>
#include <algorithm>

namespace evq {

uint32_t sql_type_of(const ColumnMeta& m) {   // sql/CSTableScanProvider.cc:79-107
  switch (m.logical_type) {
    case EVQ_COL_BOOLEAN: return EVQ_BOOL;
    case EVQ_COL_UNSIGNED_INT: return EVQ_UINT64;
    case EVQ_COL_SIGNED_INT: return EVQ_INT64;
    case EVQ_COL_FLOAT: return EVQ_FLOAT64;
    case EVQ_COL_STRING: return EVQ_STRING;
    case EVQ_COL_DATETIME: return EVQ_UINT64;
    default: return EVQ_NIL;
  }
}

static uint64_t g_next_table_uid = 1;
uint64_t next_table_uid() { return __atomic_fetch_add(&g_next_table_uid, 1, __ATOMIC_RELAXED); }

void table_init_columns(evqgpu_table* t) {
  if (t->uid == 0) t->uid = __atomic_fetch_add(&g_next_table_uid, 1, __ATOMIC_RELAXED);
  t->num_tiles = (uint32_t) ((t->num_rows + EVQ_TILE_ROWS - 1) / EVQ_TILE_ROWS);
  for (auto& c : t->cols) {
    c.sql_type = sql_type_of(c.meta);
    const bool numeric = c.sql_type == EVQ_UINT64 || c.sql_type == EVQ_FLOAT64 || c.sql_type == EVQ_BOOL;
    c.scannable = numeric && c.meta.rlevel_max == 0;
    switch (c.meta.encoding) {
      case EVQ_ENC_UINT64_PLAIN:
      case EVQ_ENC_FLOAT_IEEE754: c.data_kind = EVQ_KIND_PLAIN64; break;
      case EVQ_ENC_UINT32_PLAIN: c.data_kind = EVQ_KIND_PLAIN32; break;
      case EVQ_ENC_UINT32_BITPACKED:
      case EVQ_ENC_BOOLEAN_BITPACKED: c.data_kind = EVQ_KIND_BITPACK; break;
      case EVQ_ENC_UINT64_LEB128: c.data_kind = EVQ_KIND_LEB128; break;
      case EVQ_ENC_STRING_PLAIN:
        c.scannable = false;
        c.data_kind = EVQ_KIND_STRING_HOST;
        c.is_string = c.sql_type == EVQ_STRING && c.meta.rlevel_max == 0;
        break;
      default: c.scannable = false; break;
    }
  }
}

// ---- kernels -------------------------------------------------------------------------------------------------------

// present-value count of every row tile of an optional column (definition level == dmax)
__global__ void k_level_tile_counts(const u32* __restrict__ words, u32 bits, u32 dmax, u64 num_rows, u32 num_tiles,
                                    u64* __restrict__ counts) {
  // one warp per tile; lane handles 128-row blocks round robin
  const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  if (warp >= num_tiles) return;
  const u64 row0 = (u64) warp * EVQ_TILE_ROWS;
  const u64 rows = min((u64) EVQ_TILE_ROWS, num_rows - row0);
  u32 cnt = 0;
  for (u32 r = lane; r < rows; r += 32) {
    const u64 v = row0 + r;
    const u64 blk = v >> 7;
    const u32 i = v & 127u;
    const u32 o = (i >> 2) * bits;
    const u32* w = words + blk * 4u * bits + 4u * (o >> 5) + (i & 3u);
    const u32 sh = o & 31u;
    u32 x = w[0] >> sh;
    if (sh + bits > 32u) x |= w[4] << (32u - sh);
    if (bits < 32u) x &= (1u << bits) - 1u;
    cnt += (x == dmax);
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) counts[warp] = cnt;
}

#define EVQ_LEB_CHUNK 1024u   // bytes per terminator-count chunk

__device__ __forceinline__ u32 term_count16(uint4 c, u32 valid) {
  // valid = number of leading bytes of the 16 that belong to the payload (0..16)
  u32 w[4] = {c.x, c.y, c.z, c.w};
  u32 n = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    u32 t = ~w[k] & 0x80808080u;
    const int rem = (int) valid - 4 * k;
    if (rem <= 0) t = 0;
    else if (rem < 4) t &= (1u << (8 * rem)) - 1u;
    n += __popc(t);
  }
  return n;
}

// bit i of the result = byte i of the 16-byte chunk is a continuation byte (msb set)
__device__ __forceinline__ u32 cont_mask16(uint4 c) {
  u32 w[4] = {c.x, c.y, c.z, c.w};
  u32 m = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const u32 t = w[k] & 0x80808080u;
    m |= (((t * 0x00204081u) >> 28) & 0xfu) << (4 * k);
  }
  return m;
}

// terminator bytes (msb clear) per EVQ_LEB_CHUNK bytes of payload; also the column statistic "longest value":
// bit k-1 of *runs is set when k consecutive continuation bytes occur somewhere (k = 1..9), i.e. a value of k+1 bytes
__global__ void k_leb_chunk_counts(const uint4* __restrict__ data, u64 nbytes, u64 nchunks, u64* __restrict__ counts,
                                   unsigned int* __restrict__ runs) {
  const u64 warp = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  if (warp >= nchunks) return;
  const u64 base = warp * EVQ_LEB_CHUNK;
  u32 cnt = 0, seen = 0;
#pragma unroll
  for (u32 k = 0; k < EVQ_LEB_CHUNK / 16 / 32; ++k) {
    const u64 off = base + (u64) (k * 32 + lane) * 16;
    if (off < nbytes) {
      const u64 rem = nbytes - off;
      const uint4 q = data[off >> 4];
      cnt += term_count16(q, rem < 16 ? (u32) rem : 16u);
      // continuation bits of this chunk and the next one (streams are zero padded by >= 128 bytes: reading on is safe)
      u32 c = cont_mask16(q) | (cont_mask16(data[(off >> 4) + 1]) << 16);
      if (rem < 32) c &= (1u << rem) - 1u;   // only payload bytes
      u32 r = c;
#pragma unroll
      for (u32 len = 1; len <= 9; ++len) {
        if (r & 0xffffu) seen |= 1u << (len - 1);   // a run of `len` continuation bytes starts inside this chunk
        r &= c >> len;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    seen |= __shfl_xor_sync(0xffffffffu, seen, o);
  }
  if (lane == 0) {
    counts[warp] = cnt;
    if (seen) atomicOr(runs, seen);
  }
}

// ---- column statistics: exact minimum and maximum of the values (zone-map style; the scan kernels are specialised on them)
// out[0] = max, out[1] = ~min (both start at 0, both merged with atomicMax)

// PLAIN64 / PLAIN32 streams
template <typename T>
__global__ void k_minmax_plain(const T* __restrict__ v, u64 n, unsigned long long* __restrict__ out) {
  u64 mx = 0, mn = ~0ull;
  for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
    const u64 x = (u64) v[i];
    mx = max(mx, x);
    mn = min(mn, x);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, mx);
    atomicMax(out + 1, ~mn);
  }
}

// LEB128 streams whose values are all one byte long: value i is byte i, n = number of values
__global__ void k_minmax_bytes(const uint4* __restrict__ v, u64 n, unsigned long long* __restrict__ out) {
  u32 mx = 0, mn = 0xffffffffu;
  const u64 n16 = n >> 4;
  for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (u64) gridDim.x * blockDim.x) {
    const uint4 q = v[i];
    mx = __vmaxu4(mx, __vmaxu4(__vmaxu4(q.x, q.y), __vmaxu4(q.z, q.w)));
    mn = __vminu4(mn, __vminu4(__vminu4(q.x, q.y), __vminu4(q.z, q.w)));
  }
  mx = max(max(mx & 0xffu, (mx >> 8) & 0xffu), max((mx >> 16) & 0xffu, mx >> 24));
  mn = min(min(mn & 0xffu, (mn >> 8) & 0xffu), min((mn >> 16) & 0xffu, mn >> 24));
  if (blockIdx.x == 0 && threadIdx.x == 0) {   // the last, partial 16 bytes
    const u8* b = (const u8*) v;
    for (u64 i = n16 << 4; i < n; ++i) {
      mx = max(mx, (u32) b[i]);
      mn = min(mn, (u32) b[i]);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, (unsigned long long) mx);
    atomicMax(out + 1, ~(unsigned long long) mn);
  }
}

// variable-length LEB128 streams with a sub-index: one thread decodes the 8 values behind one sub-index entry
// (sub == nullptr: every value is `ulen` bytes long)
// (val_index != nullptr: an optional column - the tile holds val_index[tile + 1] - val_index[tile] values, the sub-index
// entries count VALUES of the tile)
__global__ void k_minmax_leb(const u8* __restrict__ data, const u64* __restrict__ off_index, const u16* __restrict__ sub, u32 ulen,
                             u32 num_tiles, u64 num_rows, const u64* __restrict__ val_index, unsigned long long* __restrict__ out) {
  const u64 idx = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  const u64 tile = idx / (EVQ_TILE_ROWS / 8);
  const u32 g = (u32) (idx % (EVQ_TILE_ROWS / 8));
  u64 mx = 0, mn = ~0ull;
  u64 tile_vals = 0;
  if (tile < num_tiles)
    tile_vals = val_index ? val_index[tile + 1] - val_index[tile] : min((u64) EVQ_TILE_ROWS, num_rows - tile * EVQ_TILE_ROWS);
  if (8ull * g < tile_vals) {
    const u32 nv = (u32) min((u64) 8, tile_vals - 8ull * g);
    const u8* p = data + off_index[tile] + (sub ? (u32) sub[tile * EVQ_SUB_ENTRIES + g * (8 / EVQ_SUB_GRAN)] : 8u * g * ulen);
    for (u32 i = 0; i < nv; ++i) {
      u64 x = 0;
      for (u32 sh = 0; sh < 70; sh += 7) {
        const u8 b = *p++;
        x |= (u64) (b & 0x7fu) << sh;          // (the 10th byte contributes its low bit only: shifts >= 64 - 7 drop the rest)
        if (!(b & 0x80u)) break;
      }
      mx = max(mx, x);
      mn = min(mn, x);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0 && mn <= mx) {
    atomicMax(out, mx);
    atomicMax(out + 1, ~mn);
  }
}

// off_index[t] = byte offset at which value number boundary(t) starts, boundary(t) = val_index[t] or t*TILE
__global__ void k_leb_select(const u8* __restrict__ data, u64 nbytes, const u64* __restrict__ chunk_base, u64 nchunks,
                             const u64* __restrict__ val_index, u64 num_rows, u32 num_tiles, u64* __restrict__ off_index) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > num_tiles) return;
  u64 V;
  if (val_index) V = val_index[t];
  else V = min((u64) t * EVQ_TILE_ROWS, num_rows);
  if (V == 0) { off_index[t] = 0; return; }
  const u64 o = V - 1;   // ordinal of the terminator that ends the previous value
  // chunk c with chunk_base[c] <= o < chunk_base[c+1]
  u64 lo = 0, hi = nchunks;
  while (hi - lo > 1) {
    const u64 mid = (lo + hi) >> 1;
    if (chunk_base[mid] <= o) lo = mid; else hi = mid;
  }
  u64 need = o - chunk_base[lo];   // terminators to skip inside the chunk
  u64 pos = lo * EVQ_LEB_CHUNK;
  const u64 end = min(nbytes, pos + EVQ_LEB_CHUNK);
  const uint4* d4 = (const uint4*) data;
  for (; pos < end; pos += 16) {
    const u64 rem = nbytes - pos;
    const uint4 q = d4[pos >> 4];
    const u32 n = term_count16(q, rem < 16 ? (u32) rem : 16u);
    if (n > need) {
      const u8* b = (const u8*) &q;
      for (u32 k = 0; k < 16; ++k) {
        if (!(b[k] & 0x80)) {
          if (need == 0) { off_index[t] = pos + k + 1; return; }
          --need;
        }
      }
    }
    need -= n;
  }
  off_index[t] = nbytes;   // fewer values than expected: the scan kernel never reads past nbytes
}

// sub_index[t][g] = byte offset, from the tile's first byte, at which value EVQ_SUB_GRAN * g of tile t starts.  One warp per tile walks
// the tile's bytes in 16-byte chunks (32 chunks per step), numbers the terminator bytes with a warp scan and records the
// byte behind every EVQ_SUB_GRAN-th one.
__global__ void k_leb_sub_index(const u8* __restrict__ data, const u64* __restrict__ off_index, u32 num_tiles, u16* __restrict__ sub) {
  const u32 tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  if (tile >= num_tiles) return;
  const u64 start = off_index[tile], end = off_index[tile + 1];
  const u64 al = start & ~15ull;
  const u32 delta = (u32) (start - al);
  const u32 tb = (u32) (end - al);                 // bytes [delta, tb) of the aligned window are the tile
  const u32 nchunks = (tb + 15u) >> 4;
  u16* out = sub + (u64) tile * EVQ_SUB_ENTRIES;
  if (lane == 0) out[0] = 0;
  u32 running = 0;
  for (u32 base = 0; base < nchunks; base += 32) {
    const u32 c = base + lane;
    u32 m = 0;
    if (c < nchunks) {
      const uint4 q = *(const uint4*) (data + al + 16ull * c);
      u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) m |= ((((~w[k] & 0x80808080u) * 0x00204081u) >> 28) & 0xfu) << (4 * k);
      const u32 pos = 16u * c;
      if (pos < delta) m &= ~((1u << (delta - pos)) - 1u);
      if (tb - pos < 16u) m &= (1u << (tb - pos)) - 1u;
    }
    const u32 cnt = __popc(m);
    u32 incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (u32) o) incl += n;
    }
    u32 j = running + incl - cnt;                  // number of this chunk's first terminator
    while (m) {
      const u32 k = __ffs(m) - 1u;
      m &= m - 1u;
      if ((j & (EVQ_SUB_GRAN - 1u)) == EVQ_SUB_GRAN - 1u) {
        const u32 g = (j + 1u) / EVQ_SUB_GRAN;
        if (g < EVQ_SUB_ENTRIES) out[g] = (u16) (16u * c + k + 1u - delta);
      }
      ++j;
    }
    running += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// max over tiles of the 16-byte aligned copy size of [idx[t]*scale, idx[t+1]*scale)
__global__ void k_max_span(const u64* __restrict__ idx, u32 num_tiles, u32 scale_bytes, u32 block_values, u32 block_bytes,
                           unsigned int* __restrict__ out) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_tiles) return;
  u64 start, end;
  if (block_values) {   // bit-packed: whole 128-value blocks
    start = (idx[t] / block_values) * block_bytes;
    end = ((idx[t + 1] + block_values - 1) / block_values) * block_bytes;
  } else {
    start = idx[t] * scale_bytes;
    end = idx[t + 1] * scale_bytes;
  }
  const u64 al = start & ~15ull;
  const u32 span = (u32) (((end - al) + 15) & ~15ull);
  atomicMax(out, span);
}

__global__ void k_mask_last_byte(u8* p, u32 keep_bits) { *p &= (u8) ((1u << keep_bits) - 1u); }

// ---- host side -----------------------------------------------------------------------------------------------------

static void exclusive_scan_u64(evqgpu_ctx* ctx, const u64* in, u64* out, uint64_t n) {
  size_t tmp_bytes = 0;
  EVQ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int64_t) n, ctx->stream));
  DevBuf tmp;
  tmp.alloc(tmp_bytes);
  EVQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, (int64_t) n, ctx->stream));
  EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->kernel_launches += 2;
}

static void upload_stream(evqgpu_table* t, DeviceStream& ds, const StreamLayout& L) {
  ds.present = L.present;
  ds.bitpack_max = L.bitpack_max;
  ds.nbytes = L.total;
  const uint64_t alloc = round_up(L.total, 256) + 256;
  ds.buf.alloc(alloc);
  // zero the tail padding so that over-reads of tile copies see defined bytes
  const uint64_t tail0 = L.total & ~255ull;
  EVQ_CUDA(cudaMemsetAsync((uint8_t*) ds.buf.p + tail0, 0, alloc - tail0, t->ctx->stream));
  uint64_t dst = 0;
  for (const auto& e : L.extents) {
    EVQ_CUDA(cudaMemcpyAsync((uint8_t*) ds.buf.p + dst, t->file + e.file_offset, e.nbytes, cudaMemcpyHostToDevice,
                             t->ctx->stream));
    dst += e.nbytes;
  }
}

void table_load_column(evqgpu_table* t, Column& c) {
  if (c.loaded) return;
  if (c.is_string) {
    // flat STRING_PLAIN column: the stream goes to the device like any other; the value index is built from the host
    // image of the stream (length prefixes are a sequential chain - strings.cu)
    if (!t->from_file) fail(EVQGPU_ERR_ARG, "column '%s' has no streams", c.meta.name.c_str());
    use_device(t->ctx);
    StreamLayout dl = stream_layout(t->meta, c.meta, EVQ_STREAM_DATA, t->file, t->file_bytes);
    upload_stream(t, c.data, dl);
    if (c.meta.dlevel_max > 0) {
      StreamLayout ll = stream_layout(t->meta, c.meta, EVQ_STREAM_DLEVEL, t->file, t->file_bytes);
      upload_stream(t, c.dlevel, ll);
      if (t->meta.version == 1) c.dlevel.bitpack_max = c.meta.dlevel_max;
    }
    if (dl.extents.size() == 1) {
      table_finish_string_column(t, c, t->file + dl.extents[0].file_offset, dl.total);
    } else {
      std::vector<uint8_t> logical(dl.total);
      uint64_t dst = 0;
      for (const auto& e : dl.extents) {
        memcpy(logical.data() + dst, t->file + e.file_offset, e.nbytes);
        dst += e.nbytes;
      }
      table_finish_string_column(t, c, logical.data(), dl.total);
    }
    return;
  }
  if (!c.scannable)
    fail(EVQGPU_ERR_UNSUPPORTED, "column '%s' (logical type %u, encoding %u, rlevel_max %u) is outside the flat numeric scan path",
         c.meta.name.c_str(), c.meta.logical_type, c.meta.encoding, c.meta.rlevel_max);
  if (!t->from_file) fail(EVQGPU_ERR_ARG, "column '%s' has no streams", c.meta.name.c_str());
  use_device(t->ctx);
  StreamLayout dl = stream_layout(t->meta, c.meta, EVQ_STREAM_DATA, t->file, t->file_bytes);
  upload_stream(t, c.data, dl);
  if (c.meta.dlevel_max > 0) {
    StreamLayout ll = stream_layout(t->meta, c.meta, EVQ_STREAM_DLEVEL, t->file, t->file_bytes);
    upload_stream(t, c.dlevel, ll);
    if (t->meta.version == 1) c.dlevel.bitpack_max = c.meta.dlevel_max;
  }
  table_finish_column(t, c);
}

void table_finish_column(evqgpu_table* t, Column& c) {
  evqgpu_ctx* ctx = t->ctx;
  use_device(ctx);
  const uint32_t ntiles = t->num_tiles;
  c.num_values = t->num_rows;
  c.level_bits = 0;
  c.level_payload_bytes = 0;

  DevBuf maxspan;
  maxspan.alloc(sizeof(unsigned int));
  c.value_min = 0;
  // exact value range of the column (k_minmax_*): {max, ~min} on the device, read back once the pass has run
  DevBuf minmax;
  minmax.alloc(16);
  auto read_minmax = [&](bool have_values) {
    u64 mm[2] = {0, 0};
    EVQ_CUDA(cudaMemcpyAsync(mm, minmax.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!have_values) return;
    c.value_max = mm[0];
    c.value_min = ~mm[1] <= mm[0] ? ~mm[1] : 0;
    c.value_bits = mm[0] ? 64 - (uint32_t) __builtin_clzll(mm[0]) : 1;
  };

  // ---- optional column: present values per tile -> val_index
  const bool nullable = c.meta.dlevel_max > 0;
  if (nullable) {
    c.level_bits = bits_needed(c.dlevel.bitpack_max);
    if (t->num_rows == 0 && c.level_bits == 0) c.level_bits = bits_needed(c.meta.dlevel_max);   // empty table: no level page at all
    if (c.level_bits == 0)
      fail(EVQGPU_ERR_FORMAT, "column '%s': definition level stream has bit width 0", c.meta.name.c_str());
    const uint64_t need = (t->num_rows + 127) / 128 * 16 * c.level_bits;
    if (c.dlevel.nbytes < need)
      fail(EVQGPU_ERR_FORMAT, "column '%s': definition level stream too short (%llu < %llu)", c.meta.name.c_str(),
           (unsigned long long) c.dlevel.nbytes, (unsigned long long) need);
    c.level_payload_bytes = need;
    c.level_tile_cap = (uint32_t) round_up((uint64_t) (EVQ_TILE_ROWS / 128) * 16 * c.level_bits, 16) + 16;
    DevBuf counts;
    counts.alloc((uint64_t) (ntiles + 1) * 8);
    c.val_index.alloc((uint64_t) (ntiles + 1) * 8);
    EVQ_CUDA(cudaMemsetAsync(counts.p, 0, counts.bytes, ctx->stream));
    if (ntiles) {
      const uint32_t threads = 256, warps_per_block = threads / 32;
      k_level_tile_counts<<<(ntiles + warps_per_block - 1) / warps_per_block, threads, 0, ctx->stream>>>(
          c.dlevel.buf.as<u32>(), c.level_bits, c.meta.dlevel_max, t->num_rows, ntiles, counts.as<u64>());
      EVQ_CUDA(cudaGetLastError());
      ctx->kernel_launches++;
    }
    exclusive_scan_u64(ctx, counts.as<u64>(), c.val_index.as<u64>(), ntiles + 1);
    u64 total = 0;
    EVQ_CUDA(cudaMemcpyAsync(&total, c.val_index.as<u64>() + ntiles, 8, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
    c.num_values = total;
  }

  // ---- data stream geometry
  const uint64_t nv = c.num_values;
  auto span_from_index = [&](const u64* idx, u32 scale, u32 block_values, u32 block_bytes) -> uint32_t {
    EVQ_CUDA(cudaMemsetAsync(maxspan.p, 0, sizeof(unsigned int), ctx->stream));
    if (ntiles) {
      k_max_span<<<(ntiles + 255) / 256, 256, 0, ctx->stream>>>(idx, ntiles, scale, block_values, block_bytes,
                                                                  maxspan.as<unsigned int>());
      EVQ_CUDA(cudaGetLastError());
      ctx->kernel_launches++;
    }
    unsigned int v = 0;
    EVQ_CUDA(cudaMemcpyAsync(&v, maxspan.p, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
    return v + 16;
  };

  switch (c.data_kind) {
    case EVQ_KIND_PLAIN64:
    case EVQ_KIND_PLAIN32: {
      const uint32_t w = c.data_kind == EVQ_KIND_PLAIN64 ? 8 : 4;
      if (c.data.nbytes < nv * w)
        fail(EVQGPU_ERR_FORMAT, "column '%s': data stream too short", c.meta.name.c_str());
      c.data_payload_bytes = nv * w;
      c.data_bits = w * 8;
      c.value_bits = w * 8;
      c.value_max = w == 8 ? ~0ull : 0xffffffffull;
      if (nv && c.sql_type != EVQ_FLOAT64) {
        EVQ_CUDA(cudaMemsetAsync(minmax.p, 0, 16, ctx->stream));
        if (w == 8) k_minmax_plain<u64><<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(c.data.buf.as<u64>(), nv, minmax.as<unsigned long long>());
        else k_minmax_plain<u32><<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(c.data.buf.as<u32>(), nv, minmax.as<unsigned long long>());
        EVQ_CUDA(cudaGetLastError());
        ctx->kernel_launches++;
        read_minmax(true);
      }
      if (nullable) c.data_tile_cap = span_from_index(c.val_index.as<u64>(), w, 0, 0);
      else c.data_tile_cap = EVQ_TILE_ROWS * w + 32;
      break;
    }
    case EVQ_KIND_BITPACK: {
      c.data_bits = bits_needed(c.data.bitpack_max);
      c.value_bits = std::max<uint32_t>(1, c.data_bits);
      c.value_max = c.value_bits >= 64 ? ~0ull : (1ull << c.value_bits) - 1;
      const uint64_t need = (nv + 127) / 128 * 16 * c.data_bits;
      if (c.data_bits && c.data.nbytes < need)
        fail(EVQGPU_ERR_FORMAT, "column '%s': bit-packed stream too short", c.meta.name.c_str());
      c.data_payload_bytes = need + (c.data_bits ? 4 : 0);
      if (nullable) c.data_tile_cap = span_from_index(c.val_index.as<u64>(), 0, 128, 16 * c.data_bits);
      else c.data_tile_cap = (EVQ_TILE_ROWS / 128) * 16 * c.data_bits + 32;
      break;
    }
    case EVQ_KIND_LEB128: {
      const uint64_t nchunks = std::max<uint64_t>(1, (c.data.nbytes + EVQ_LEB_CHUNK - 1) / EVQ_LEB_CHUNK);
      DevBuf counts, base;
      counts.alloc((nchunks + 1) * 8);
      base.alloc((nchunks + 1) * 8);
      EVQ_CUDA(cudaMemsetAsync(counts.p, 0, counts.bytes, ctx->stream));
      EVQ_CUDA(cudaMemsetAsync(maxspan.p, 0, sizeof(unsigned int), ctx->stream));
      {
        const uint32_t threads = 256;
        const uint64_t blocks = (nchunks * 32 + threads - 1) / threads;
        k_leb_chunk_counts<<<(unsigned) blocks, threads, 0, ctx->stream>>>(c.data.buf.as<uint4>(), c.data.nbytes, nchunks,
                                                                           counts.as<u64>(), maxspan.as<unsigned int>());
        EVQ_CUDA(cudaGetLastError());
        ctx->kernel_launches++;
      }
      {
        // longest value of the column in bytes -> static decode width of the query kernels
        unsigned int runs = 0;
        EVQ_CUDA(cudaMemcpyAsync(&runs, maxspan.p, sizeof(runs), cudaMemcpyDeviceToHost, ctx->stream));
        EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
        c.leb_max_len = 1;
        while (c.leb_max_len < 10 && (runs >> (c.leb_max_len - 1)) & 1u) ++c.leb_max_len;
        c.value_bits = std::min<uint32_t>(64, 7 * c.leb_max_len);
        c.value_max = c.value_bits >= 64 ? ~0ull : (1ull << c.value_bits) - 1;
        if (c.leb_max_len == 1 && nv && c.data.nbytes >= nv) {
          // all values are single bytes: value i is byte i, the exact range is one more cheap pass
          EVQ_CUDA(cudaMemsetAsync(minmax.p, 0, 16, ctx->stream));
          k_minmax_bytes<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(c.data.buf.as<uint4>(), nv, minmax.as<unsigned long long>());
          EVQ_CUDA(cudaGetLastError());
          ctx->kernel_launches++;
          read_minmax(true);
        }
      }
      exclusive_scan_u64(ctx, counts.as<u64>(), base.as<u64>(), nchunks + 1);
      u64 total_terms = 0;
      EVQ_CUDA(cudaMemcpyAsync(&total_terms, base.as<u64>() + nchunks, 8, cudaMemcpyDeviceToHost, ctx->stream));
      EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
      if (total_terms < nv)
        fail(EVQGPU_ERR_FORMAT, "column '%s': LEB128 stream holds %llu values, %llu expected", c.meta.name.c_str(),
             (unsigned long long) total_terms, (unsigned long long) nv);
      c.off_index.alloc((uint64_t) (ntiles + 1) * 8);
      k_leb_select<<<(ntiles + 1 + 127) / 128, 128, 0, ctx->stream>>>(
          c.data.buf.as<u8>(), c.data.nbytes, base.as<u64>(), nchunks, nullable ? c.val_index.as<u64>() : nullptr,
          t->num_rows, ntiles, c.off_index.as<u64>());
      EVQ_CUDA(cudaGetLastError());
      ctx->kernel_launches++;
      u64 used = 0;
      EVQ_CUDA(cudaMemcpyAsync(&used, c.off_index.as<u64>() + ntiles, 8, cudaMemcpyDeviceToHost, ctx->stream));
      EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
      c.data_payload_bytes = used;
      c.data_bits = 0;
      c.data_tile_cap = span_from_index(c.off_index.as<u64>(), 1, 0, 0);
      c.leb_uniform = nv > 0 && used == (u64) c.leb_max_len * nv;
      if (c.leb_max_len >= 2 && ntiles) {
        if (!c.leb_uniform) {
          // where the values of the column differ in length: starts of every EVQ_SUB_GRAN-th value (the fast kernel's decode
          // entry points)
          c.sub_index.alloc((uint64_t) ntiles * EVQ_SUB_ENTRIES * 2 + 256);
          EVQ_CUDA(cudaMemsetAsync(c.sub_index.p, 0, c.sub_index.bytes, ctx->stream));
          k_leb_sub_index<<<(unsigned) (((uint64_t) ntiles * 32 + 255) / 256), 256, 0, ctx->stream>>>(
              c.data.buf.as<u8>(), c.off_index.as<u64>(), ntiles, c.sub_index.as<u16>());
          EVQ_CUDA(cudaGetLastError());
          ctx->kernel_launches++;
        }
        // ... which also make the exact value range one cheap pass (8 values per thread; of an optional column: its present values)
        EVQ_CUDA(cudaMemsetAsync(minmax.p, 0, 16, ctx->stream));
        const uint64_t groups = (uint64_t) ntiles * (EVQ_TILE_ROWS / 8);
        k_minmax_leb<<<(unsigned) ((groups + 255) / 256), 256, 0, ctx->stream>>>(
            c.data.buf.as<u8>(), c.off_index.as<u64>(), c.leb_uniform ? nullptr : c.sub_index.as<u16>(), c.leb_max_len, ntiles,
            t->num_rows, nullable ? c.val_index.as<u64>() : nullptr, minmax.as<unsigned long long>());
        EVQ_CUDA(cudaGetLastError());
        ctx->kernel_launches++;
        read_minmax(nv > 0);
      }
      break;
    }
  }
  // NULLs read as value 0 wherever the tag is ignored
  c.value_min_present = c.value_min;   // over the values that are in the data stream (what a decoder of the stream meets)
  if (nullable && c.num_values < t->num_rows) c.value_min = 0;
  c.loaded = true;
}

}  // namespace evq

// ---- C ABI ---------------------------------------------------------------------------------------------------------
using namespace evq;

extern "C" {

int evqgpu_table_open(evqgpu_ctx* ctx, const void* file, uint64_t nbytes, evqgpu_table** out) {
  return guarded([&] {
    if (!ctx || !file || !out) fail(EVQGPU_ERR_ARG, "evqgpu_table_open: null argument");
    std::unique_ptr<evqgpu_table> t(new evqgpu_table());
    t->ctx = ctx;
    t->file = (const uint8_t*) file;
    t->file_bytes = nbytes;
    t->from_file = true;
    t->meta = parse_cstable(t->file, nbytes);
    t->num_rows = t->meta.num_rows;
    for (const auto& cm : t->meta.columns) {
      Column c;
      c.meta = cm;
      t->cols.push_back(std::move(c));
    }
    table_init_columns(t.get());
    *out = t.release();
  });
}

int evqgpu_table_create(evqgpu_ctx* ctx, uint64_t num_rows, evqgpu_table** out) {
  return guarded([&] {
    if (!ctx || !out) fail(EVQGPU_ERR_ARG, "evqgpu_table_create: null argument");
    std::unique_ptr<evqgpu_table> t(new evqgpu_table());
    t->ctx = ctx;
    t->num_rows = num_rows;
    t->meta.version = 2;
    t->meta.num_rows = num_rows;
    table_init_columns(t.get());
    *out = t.release();
  });
}

int evqgpu_table_add_column(evqgpu_table* tbl, const char* name, uint32_t logical_type, uint32_t encoding,
                            uint32_t rlevel_max, uint32_t dlevel_max) {
  int idx = -1;
  int rc = guarded([&] {
    if (!tbl || !name) fail(EVQGPU_ERR_ARG, "evqgpu_table_add_column: null argument");
    if (tbl->find(name) >= 0) fail(EVQGPU_ERR_ARG, "duplicate column '%s'", name);
    Column c;
    c.meta.name = name;
    c.meta.column_id = (uint32_t) tbl->cols.size() + 1;   // TableSchema.cc:55-70: ids 1.. in add order
    c.meta.logical_type = logical_type;
    c.meta.encoding = encoding;
    c.meta.rlevel_max = rlevel_max;
    c.meta.dlevel_max = dlevel_max;
    tbl->cols.push_back(std::move(c));
    table_init_columns(tbl);
    idx = (int) tbl->cols.size() - 1;
  });
  return rc == EVQGPU_OK ? idx : -rc;
}

int evqgpu_table_add_stream(evqgpu_table* tbl, const char* column, uint32_t kind, const void* ptr, uint64_t nbytes,
                            uint32_t bitpack_max, uint32_t flags) {
  return guarded([&] {
    if (!tbl || !column || (!ptr && nbytes)) fail(EVQGPU_ERR_ARG, "evqgpu_table_add_stream: null argument");
    const int ci = tbl->find(column);
    if (ci < 0) fail(EVQGPU_ERR_ARG, "unknown column '%s'", column);
    Column& c = tbl->cols[ci];
    if (kind != EVQ_STREAM_DATA && kind != EVQ_STREAM_DLEVEL)
      fail(EVQGPU_ERR_UNSUPPORTED, "only DATA and DLEVEL streams are supported (flat columns)");
    use_device(tbl->ctx);
    DeviceStream& ds = kind == EVQ_STREAM_DATA ? c.data : c.dlevel;
    ds.present = true;
    ds.bitpack_max = bitpack_max;
    ds.nbytes = nbytes;
    const uint64_t alloc = round_up(nbytes, 256) + 256;
    ds.buf.alloc(alloc);
    const uint64_t tail0 = nbytes & ~255ull;
    EVQ_CUDA(cudaMemsetAsync((uint8_t*) ds.buf.p + tail0, 0, alloc - tail0, tbl->ctx->stream));
    if (nbytes)
      EVQ_CUDA(cudaMemcpyAsync(ds.buf.p, ptr, nbytes,
                               (flags & EVQGPU_STREAM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                               tbl->ctx->stream));
    const bool ready = c.data.present && (c.meta.dlevel_max == 0 || c.dlevel.present);
    if (ready && c.is_string) {
      std::vector<uint8_t> logical(c.data.nbytes);
      if (c.data.nbytes)
        EVQ_CUDA(cudaMemcpyAsync(logical.data(), c.data.buf.p, c.data.nbytes, cudaMemcpyDeviceToHost, tbl->ctx->stream));
      EVQ_CUDA(cudaStreamSynchronize(tbl->ctx->stream));
      table_finish_string_column(tbl, c, logical.data(), c.data.nbytes);
    } else if (ready) {
      if (!c.scannable) fail(EVQGPU_ERR_UNSUPPORTED, "column '%s' is outside the flat numeric scan path", column);
      table_finish_column(tbl, c);
    }
  });
}

int evqgpu_table_set_filter(evqgpu_table* tbl, const void* bits, uint64_t nrows, uint32_t flags) {
  return guarded([&] {
    if (!tbl) fail(EVQGPU_ERR_ARG, "evqgpu_table_set_filter: null table");
    use_device(tbl->ctx);
    if (!bits) {
      tbl->filter.release();
      tbl->has_filter = false;
      return;
    }
    if (nrows != tbl->num_rows)
      fail(EVQGPU_ERR_ARG, "evqgpu_table_set_filter: the filter has %llu rows, the table %llu", (unsigned long long) nrows,
           (unsigned long long) tbl->num_rows);
    // one 128-byte line per row tile, zero padded: rows behind the end are dropped anyway
    const uint64_t alloc = round_up((uint64_t) tbl->num_tiles * (EVQ_TILE_ROWS / 8), 256) + 256;
    tbl->filter.alloc(alloc);
    EVQ_CUDA(cudaMemsetAsync(tbl->filter.p, 0, tbl->filter.bytes, tbl->ctx->stream));
    const uint64_t nbytes = (nrows + 7) / 8;
    if (nbytes)
      EVQ_CUDA(cudaMemcpyAsync(tbl->filter.p, bits, nbytes, (flags & EVQGPU_STREAM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                               tbl->ctx->stream));
    if (nrows % 8) {   // clear the bits behind the last row in the last byte
      k_mask_last_byte<<<1, 1, 0, tbl->ctx->stream>>>(tbl->filter.as<u8>() + nbytes - 1, (u32) (nrows % 8));
      EVQ_CUDA(cudaGetLastError());
    }
    EVQ_CUDA(cudaStreamSynchronize(tbl->ctx->stream));   // the caller's buffer may go away
    tbl->has_filter = true;
  });
}

void evqgpu_table_destroy(evqgpu_table* tbl) {
  if (!tbl) return;
  cudaSetDevice(tbl->ctx->device);
  auto& cc = tbl->ctx->code_columns;
  cc.erase(std::remove_if(cc.begin(), cc.end(), [&](const std::pair<void*, void*>& e) { return e.first == (void*) tbl; }), cc.end());
  delete tbl;
}

uint64_t evqgpu_table_num_rows(const evqgpu_table* tbl) { return tbl ? tbl->num_rows : 0; }
uint32_t evqgpu_table_num_columns(const evqgpu_table* tbl) { return tbl ? (uint32_t) tbl->cols.size() : 0; }

int evqgpu_table_column_info(const evqgpu_table* tbl, uint32_t idx, evqgpu_column_info* out) {
  return guarded([&] {
    if (!tbl || !out || idx >= tbl->cols.size()) fail(EVQGPU_ERR_ARG, "evqgpu_table_column_info: bad argument");
    const Column& c = tbl->cols[idx];
    memset(out, 0, sizeof(*out));
    out->name = c.meta.name.c_str();
    out->column_id = c.meta.column_id;
    out->logical_type = c.meta.logical_type;
    out->encoding = c.meta.encoding;
    out->rlevel_max = c.meta.rlevel_max;
    out->dlevel_max = c.meta.dlevel_max;
    out->sql_type = c.sql_type;
    out->loaded = c.loaded;
    out->data_bytes = c.data_payload_bytes;
    out->level_bytes = c.level_payload_bytes;
    out->num_values = c.num_values;
    out->value_bits = c.value_bits;
    out->leb_max_len = c.leb_max_len;
    out->value_min = c.value_min;
    out->value_max = c.value_max;
  });
}

int evqgpu_table_find_column(const evqgpu_table* tbl, const char* name) { return (tbl && name) ? tbl->find(name) : -1; }

int evqgpu_table_load_columns(evqgpu_table* tbl, const char* const* names, uint32_t n) {
  return guarded([&] {
    if (!tbl) fail(EVQGPU_ERR_ARG, "evqgpu_table_load_columns: null table");
    if (!names) {
      for (auto& c : tbl->cols)
        if (c.scannable && !c.loaded) table_load_column(tbl, c);
      return;
    }
    for (uint32_t i = 0; i < n; ++i) {
      const int ci = tbl->find(names[i]);
      if (ci < 0) fail(EVQGPU_ERR_ARG, "column not found: %s", names[i]);
      table_load_column(tbl, tbl->cols[ci]);
    }
  });
}

int evqgpu_table_read_stream(evqgpu_table* tbl, const char* column, uint32_t kind, void* dst, uint64_t cap,
                             uint64_t* nbytes_out, uint32_t* bitpack_max_out) {
  return guarded([&] {
    if (!tbl || !column || !nbytes_out) fail(EVQGPU_ERR_ARG, "evqgpu_table_read_stream: null argument");
    const int ci = tbl->find(column);
    if (ci < 0) fail(EVQGPU_ERR_ARG, "column not found: %s", column);
    Column& c = tbl->cols[ci];
    DeviceStream& ds = kind == EVQ_STREAM_DATA ? c.data : c.dlevel;
    if (!ds.present) { *nbytes_out = 0; return; }
    uint64_t n = ds.nbytes;
    if (kind == EVQ_STREAM_DATA && c.loaded && c.data_kind == EVQ_KIND_LEB128) n = c.data_payload_bytes;
    if (kind == EVQ_STREAM_DLEVEL && c.loaded) n = c.level_payload_bytes;
    *nbytes_out = n;
    if (bitpack_max_out) *bitpack_max_out = ds.bitpack_max;
    if (dst && cap >= n && n) {
      use_device(tbl->ctx);
      EVQ_CUDA(cudaMemcpyAsync(dst, ds.buf.p, n, cudaMemcpyDeviceToHost, tbl->ctx->stream));
      EVQ_CUDA(cudaStreamSynchronize(tbl->ctx->stream));
    }
  });
}

int evqgpu_table_write_file(evqgpu_table* tbl, const char* path) {
  return guarded([&] {
    if (!tbl || !path) fail(EVQGPU_ERR_ARG, "evqgpu_table_write_file: null argument");
    use_device(tbl->ctx);
    std::vector<ColumnMeta> metas;
    std::vector<std::vector<uint8_t>> payloads;
    std::vector<WriteStream> streams;
    payloads.reserve(tbl->cols.size() * 2);
    for (auto& c : tbl->cols) {
      if (!c.loaded) fail(EVQGPU_ERR_ARG, "column '%s' is not resident", c.meta.name.c_str());
      metas.push_back(c.meta);
      auto fetch = [&](DeviceStream& ds, uint64_t n) -> const uint8_t* {
        payloads.emplace_back(n);
        if (n) EVQ_CUDA(cudaMemcpyAsync(payloads.back().data(), ds.buf.p, n, cudaMemcpyDeviceToHost, tbl->ctx->stream));
        return payloads.back().data();
      };
      if (c.meta.dlevel_max > 0) {
        const uint8_t* p = fetch(c.dlevel, c.level_payload_bytes);
        streams.push_back({EVQ_STREAM_DLEVEL, c.meta.column_id, p, c.level_payload_bytes, true, c.dlevel.bitpack_max});
      }
      const bool bp = c.data_kind == EVQ_KIND_BITPACK;
      const uint64_t n = bp ? (c.data_payload_bytes >= 4 ? c.data_payload_bytes - 4 : 0) : c.data_payload_bytes;
      const uint8_t* p = fetch(c.data, n);
      streams.push_back({EVQ_STREAM_DATA, c.meta.column_id, p, n, bp, c.data.bitpack_max});
    }
    EVQ_CUDA(cudaStreamSynchronize(tbl->ctx->stream));
    write_cstable_v2(path, tbl->num_rows, metas, streams);
  });
}

}  // extern "C"
