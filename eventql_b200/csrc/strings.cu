// strings.cu - flat STRING_PLAIN columns on the device, and the LSM visibility filter built from them.
//
// (1) String columns.  Replaces StringColumnReader::readString over LenencStringPageReader
//     (io/cstable/columns/column_reader_string.cc, page_reader_lenencstring.cc:37-62; v0.1.0:
//     columns/v1/StringColumnReader.cc:94-113) and FastCSTableScan::fetchColumnString (sql/CSTableScan.cc:970-995).
//     The data stream is resident like any other column's.  A value's position depends on every length prefix before it
//     (`varuint length + bytes`, not self-synchronising), so the value index {start, length} is built by ONE sequential
//     walk over the prefixes when the column is loaded - on the host, which holds the file image anyway - and uploaded.
//     Everything per row (definition levels -> value ordinal, output offsets, the byte gather into the packed SVector
//     encoding) runs on the device.
//
// (2) LSM visibility filter.  Replaces the filter loop of PartitionCursor::openNextTable
//     (server/sql/partition_cursor.cc:157-194): over the segments of a partition, newest first, a row is dropped if it is
//     skipped or if an earlier row was a visible update with the same 20-byte __lsm_id; visible update rows record their
//     id.  The reference walks the rows with a std::set<SHA1Hash>.  That loop is order dependent only through "an earlier
//     row": id X is in the set before row i  <=>  some row j < i has id X, is_update and is not skipped (the smallest such
//     j is either visible, and inserts X, or hidden because X is in the set already).  So the filter is a semi-join of
//     the rows with the non-skipped update rows on (id equal, ordinal smaller), done here with a sort:
//       gather   per row: the id (5 words), a 64-bit hash of it, the flags {is_update, skip}
//       sort     (hash, global row ordinal) pairs, stable radix sort -> rows of one id are adjacent, ordinals ascending
//       resolve  a row is hidden iff an entry before it in its hash run has the same id bytes and is a non-skipped update
//       pack     1 bit per row into the table's filter stream (the layout evqgpu_table_set_filter produces)
//     Exact: hash collisions only make runs longer, ids are compared byte for byte.
//
// (3) String predicates and string GROUP BY keys.  eq / neq between string columns and literals
//     (sql/expressions/boolean.cc:235-257, 355-377) and bare string columns as keys / select items run in the scan kernels on
//     dictionary codes: a UINT32_PLAIN shadow column per string column, codes from one dictionary per context (section
//     "dictionary codes" below; the rewrite of the plan is query.cu: lower_strings).
#include "table.h"
#include "expr.h"
#include <cub/cub.cuh>
#include <string.h>
#include <algorithm>
#include <memory>

namespace evq {

// ---- definition levels -> value ordinal of every record ---------------------------------------------------------------

__device__ __forceinline__ u32 vertical_get(const u32* __restrict__ words, u32 bits, u64 v) {
  // libsimdcomp vertical layout: 128-value blocks of 4 * bits words, value i of a block in lane i & 3
  const u64 blk = v >> 7;
  const u32 i = (u32) (v & 127u);
  const u32 o = (i >> 2) * bits;
  const u32* w = words + blk * 4u * bits + 4u * (o >> 5) + (i & 3u);
  const u32 sh = o & 31u;
  u32 x = w[0] >> sh;
  if (sh + bits > 32u) x |= w[4] << (32u - sh);
  if (bits < 32u) x &= (1u << bits) - 1u;
  return x;
}

// one warp per row tile; records in order, 32 at a time
__global__ void k_row_values(const u32* __restrict__ words, u32 bits, u32 dmax, u64 num_rows, u32 num_tiles,
                             const u64* __restrict__ val_index, u32* __restrict__ row_value) {
  const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  if (warp >= num_tiles) return;
  const u64 row0 = (u64) warp * EVQ_TILE_ROWS;
  u32 base = (u32) val_index[warp];
  for (u32 k = 0; k < EVQ_TILE_ROWS / 32; ++k) {
    const u64 row = row0 + k * 32 + lane;
    const bool valid = row < num_rows;
    const bool present = valid && vertical_get(words, bits, row) == dmax;
    const u32 b = __ballot_sync(0xffffffffu, present);
    if (valid) row_value[row] = present ? base + __popc(b & ((1u << lane) - 1u)) : 0xffffffffu;
    base += __popc(b);
  }
}

// ---- fetchColumnString: rows [row0, row0 + n) as packed STRING SVector elements ----------------------------------------

__global__ void k_str_sizes(const u32* __restrict__ row_value, const u32* __restrict__ len, u64 row0, u64 n, u64* __restrict__ sizes) {
  const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { sizes[i] = 0; return; }
  const u32 v = row_value ? row_value[row0 + i] : (u32) (row0 + i);
  sizes[i] = 5ull + (v == 0xffffffffu ? 0u : len[v]);   // [u32 length][bytes][tag]
}

// one warp per row
__global__ void k_str_emit(const u8* __restrict__ data, const u64* __restrict__ start, const u32* __restrict__ len,
                           const u32* __restrict__ row_value, u64 row0, u64 n, const u64* __restrict__ out_off, u8* __restrict__ out) {
  const u64 i = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  if (i >= n) return;
  const u32 v = row_value ? row_value[row0 + i] : (u32) (row0 + i);
  const bool null = v == 0xffffffffu;
  const u32 l = null ? 0u : len[v];
  u8* o = out + out_off[i];
  if (lane < 4) o[lane] = (u8) (l >> (8 * lane));
  if (lane == 4) o[4 + l] = null ? (u8) EVQ_STAG_NULL : (u8) 0;
  if (l) {
    const u8* s = data + start[v];
    for (u32 k = lane; k < l; k += 32) o[4 + k] = s[k];
  }
}

// ---- host: the value index ---------------------------------------------------------------------------------------------

void table_finish_string_column(evqgpu_table* t, Column& c, const uint8_t* host, uint64_t nbytes) {
  evqgpu_ctx* ctx = t->ctx;
  use_device(ctx);
  // definition levels: val_index + number of values (the generic part of table_finish_column; no data geometry for
  // data_kind EVQ_KIND_STRING_HOST)
  table_finish_column(t, c);
  c.loaded = false;
  const uint64_t nv = c.num_values;
  if (t->num_rows >= 0xffffffffull) fail(EVQGPU_ERR_UNSUPPORTED, "string column '%s': more than 2^32 - 1 rows in one table", c.meta.name.c_str());
  std::vector<uint64_t> start(nv);
  std::vector<uint32_t> len(nv);
  uint64_t pos = 0, longest = 0;
  const bool v1 = t->meta.version == 1;
  for (uint64_t i = 0; i < nv; ++i) {
    uint64_t l = 0;
    if (v1) {   // u32 length
      if (pos + 4 > nbytes) fail(EVQGPU_ERR_FORMAT, "column '%s': end of column reached", c.meta.name.c_str());
      uint32_t l32;
      memcpy(&l32, host + pos, 4);
      l = l32;
      pos += 4;
    } else {    // varuint length
      for (int k = 0;; ++k) {
        if (pos >= nbytes) fail(EVQGPU_ERR_FORMAT, "column '%s': end of column reached", c.meta.name.c_str());
        const uint8_t b = host[pos++];
        if (k < 10) l |= (uint64_t) (b & 0x7f) << (7 * k);
        if (!(b & 0x80)) break;
      }
    }
    if (l > nbytes - pos) fail(EVQGPU_ERR_FORMAT, "column '%s': end of column reached", c.meta.name.c_str());
    if (l > 0xfffffff0ull) fail(EVQGPU_ERR_UNSUPPORTED, "column '%s': string value longer than 4 GiB", c.meta.name.c_str());
    start[i] = pos;
    len[i] = (uint32_t) l;
    longest = std::max(longest, l);
    pos += l;
  }
  c.data_payload_bytes = pos;
  c.value_bits = 0;
  c.value_min = 0;
  c.value_max = longest;   // statistic of a string column: the longest value in bytes
  c.str_start.alloc(nv * 8);
  c.str_len.alloc(nv * 4);
  if (nv) {
    EVQ_CUDA(cudaMemcpyAsync(c.str_start.p, start.data(), nv * 8, cudaMemcpyHostToDevice, ctx->stream));
    EVQ_CUDA(cudaMemcpyAsync(c.str_len.p, len.data(), nv * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (c.meta.dlevel_max > 0 && t->num_rows) {
    c.row_value.alloc(t->num_rows * 4);
    const uint32_t threads = 256, warps_per_block = threads / 32;
    k_row_values<<<(t->num_tiles + warps_per_block - 1) / warps_per_block, threads, 0, ctx->stream>>>(
        c.dlevel.buf.as<u32>(), c.level_bits, c.meta.dlevel_max, t->num_rows, t->num_tiles, c.val_index.as<u64>(), c.row_value.as<u32>());
    EVQ_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
  }
  EVQ_CUDA(cudaStreamSynchronize(ctx->stream));   // the host vectors go away
  c.loaded = true;
}

// ---- dictionary codes: string predicates and string group keys ---------------------------------------------------------
// eq / neq between string columns and literals, and GROUP BY / select of a bare string column, run in the scan kernels on
// dense codes: equality of codes <=> equality of bytes because every table of a context shares one dictionary.  The codes
// are assigned by a host pass over the column's values when the first query needs them (a hash-map lookup per value, once
// per column); the shadow column then goes through the same index / statistics pass as any UINT32_PLAIN column, so the
// kernels see narrow keys with an exact range (a handful of distinct flags -> the dense register tier).

uint32_t string_code(evqgpu_ctx* ctx, const std::string& s) {
  if (ctx->code_strings.empty()) {
    ctx->code_strings.push_back(std::string());
    ctx->string_codes.emplace(std::string(), 0u);
  }
  auto it = ctx->string_codes.find(s);
  if (it != ctx->string_codes.end()) return it->second;
  if (ctx->code_strings.size() >= 0xfffffff0ull) fail(EVQGPU_ERR_UNSUPPORTED, "string dictionary of the context is full");
  const uint32_t code = (uint32_t) ctx->code_strings.size();
  ctx->code_strings.push_back(s);
  ctx->string_codes.emplace(s, code);
  return code;
}

const Column* ensure_code_column(evqgpu_table* t, Column& c) {
  if (c.code_col) return c.code_col.get();
  if (!c.is_string) fail(EVQGPU_ERR_ARG, "column '%s' is not a string column", c.meta.name.c_str());
  if (!c.loaded) table_load_column(t, c);
  evqgpu_ctx* ctx = t->ctx;
  use_device(ctx);
  const uint64_t nv = c.num_values;
  std::vector<uint8_t> data(c.data_payload_bytes);
  std::vector<uint64_t> start(nv);
  std::vector<uint32_t> len(nv), codes(nv);
  if (!data.empty()) EVQ_CUDA(cudaMemcpyAsync(data.data(), c.data.buf.p, data.size(), cudaMemcpyDeviceToHost, ctx->stream));
  if (nv) {
    EVQ_CUDA(cudaMemcpyAsync(start.data(), c.str_start.p, nv * 8, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaMemcpyAsync(len.data(), c.str_len.p, nv * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
  std::string key;
  for (uint64_t i = 0; i < nv; ++i) {
    key.assign((const char*) data.data() + start[i], len[i]);
    codes[i] = string_code(ctx, key);
  }
  std::unique_ptr<Column> sh(new Column());
  sh->meta = c.meta;
  sh->meta.logical_type = EVQ_COL_UNSIGNED_INT;
  sh->meta.encoding = EVQ_ENC_UINT32_PLAIN;
  sh->sql_type = EVQ_UINT64;
  sh->scannable = true;
  sh->data_kind = EVQ_KIND_PLAIN32;
  sh->data.present = true;
  sh->data.nbytes = nv * 4;
  const uint64_t alloc = round_up(nv * 4, 256) + 256;
  sh->data.buf.alloc(alloc);
  const uint64_t tail0 = (nv * 4) & ~255ull;
  EVQ_CUDA(cudaMemsetAsync((uint8_t*) sh->data.buf.p + tail0, 0, alloc - tail0, ctx->stream));
  if (nv) EVQ_CUDA(cudaMemcpyAsync(sh->data.buf.p, codes.data(), nv * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (c.meta.dlevel_max > 0) {
    sh->dlevel.present = c.dlevel.present;
    sh->dlevel.nbytes = c.dlevel.nbytes;
    sh->dlevel.bitpack_max = c.dlevel.bitpack_max;
    sh->dlevel.buf.alloc(c.dlevel.buf.bytes);
    EVQ_CUDA(cudaMemcpyAsync(sh->dlevel.buf.p, c.dlevel.buf.p, c.dlevel.buf.bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  EVQ_CUDA(cudaStreamSynchronize(ctx->stream));   // `codes` goes away
  table_finish_column(t, *sh);
  c.code_col = std::move(sh);
  ctx->code_columns.push_back({(void*) t, (void*) &c});
  return c.code_col.get();
}

__global__ void k_remap_codes(u32* __restrict__ codes, u64 n, const u32* __restrict__ remap, u32 first) {
  for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
    const u32 c = codes[i];
    if (c >= first) codes[i] = remap[c - first];
  }
}

std::vector<uint64_t> comm_all_gather_host(evqgpu_ctx* ctx, const std::vector<uint64_t>& mine);   // comm.cc

bool sync_dictionary(evqgpu_ctx* ctx) {
  if (ctx->code_strings.empty()) string_code(ctx, std::string());
  if (ctx->dict_agreed == 0) ctx->dict_agreed = 1;   // code 0 = "" everywhere
  const uint32_t first = ctx->dict_agreed;
  // 1. my provisional strings, serialised [u32 length][bytes], padded to 8
  std::vector<uint8_t> blob;
  for (size_t i = first; i < ctx->code_strings.size(); ++i) {
    const uint32_t len = (uint32_t) ctx->code_strings[i].size();
    blob.insert(blob.end(), (const uint8_t*) &len, (const uint8_t*) &len + 4);
    blob.insert(blob.end(), ctx->code_strings[i].begin(), ctx->code_strings[i].end());
  }
  const uint64_t my_n = ctx->code_strings.size() - first, my_bytes = blob.size();
  const std::vector<uint64_t> sizes = comm_all_gather_host(ctx, std::vector<uint64_t>{my_n, my_bytes});
  uint64_t max_bytes = 0, total_new = 0;
  for (int r = 0; r < ctx->nranks; ++r) { max_bytes = std::max(max_bytes, sizes[2 * r + 1]); total_new += sizes[2 * r]; }
  if (total_new == 0) return false;   // (every rank sees the same sizes: the same decision everywhere)
  const size_t words = (size_t) (max_bytes + 7) / 8;
  std::vector<uint64_t> mine(words, 0);
  if (!blob.empty()) memcpy(mine.data(), blob.data(), blob.size());
  const std::vector<uint64_t> all = comm_all_gather_host(ctx, mine);
  // 2. the agreed prefix + the union of everybody's new strings in rank order
  std::vector<std::string> provisional(ctx->code_strings.begin() + first, ctx->code_strings.end());
  ctx->code_strings.resize(first);
  ctx->string_codes.clear();
  for (uint32_t i = 0; i < first; ++i) ctx->string_codes.emplace(ctx->code_strings[i], i);
  for (int r = 0; r < ctx->nranks; ++r) {
    const uint8_t* p = (const uint8_t*) (all.data() + (size_t) r * words);
    uint64_t pos = 0;
    for (uint64_t i = 0; i < sizes[2 * r]; ++i) {
      uint32_t len;
      memcpy(&len, p + pos, 4);
      std::string s((const char*) p + pos + 4, len);
      pos += 4 + len;
      if (ctx->string_codes.find(s) == ctx->string_codes.end()) {
        ctx->string_codes.emplace(s, (uint32_t) ctx->code_strings.size());
        ctx->code_strings.push_back(std::move(s));
      }
    }
  }
  ctx->dict_agreed = (uint32_t) ctx->code_strings.size();
  // 3. renumber my provisional codes where they moved
  std::vector<uint32_t> remap(provisional.size());
  bool changed = false;
  for (size_t i = 0; i < provisional.size(); ++i) {
    remap[i] = ctx->string_codes.at(provisional[i]);
    changed = changed || remap[i] != first + i;
  }
  if (!changed || provisional.empty()) return changed;
  use_device(ctx);
  DevBuf dremap;
  dremap.alloc(remap.size() * 4);
  EVQ_CUDA(cudaMemcpyAsync(dremap.p, remap.data(), remap.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  for (auto& tc : ctx->code_columns) {
    evqgpu_table* t = (evqgpu_table*) tc.first;
    Column* c = (Column*) tc.second;
    if (!c->code_col) continue;
    Column& cc = *c->code_col;
    const uint64_t nv = cc.num_values;
    if (nv) {
      k_remap_codes<<<(unsigned) std::min<uint64_t>((nv + 255) / 256, (uint64_t) ctx->sm_count * 16), 256, 0, ctx->stream>>>(
          cc.data.buf.as<u32>(), nv, dremap.as<u32>(), first);
      EVQ_CUDA(cudaGetLastError());
      ctx->kernel_launches++;
    }
    table_finish_column(t, cc);   // the value range of the codes changed
    t->uid = next_table_uid();    // (cached key bounds of queries over this table are stale)
  }
  EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
  return true;
}

// ---- string predicates evaluated once per dictionary entry --------------------------------------------------------------
// lt / lte / gt / gte (expressions/boolean.cc:439-710: strncmp over the shorter length, then the lengths - embedded NUL
// bytes end the comparison like they do there) and startswith / endswith (expressions/string.cc:52-74, StringUtil) between a
// string column and a literal.
bool string_predicate_eval(const StringPredicate& p, const std::string& value) {
  const std::string& left = p.column_first ? value : p.literal;
  const std::string& right = p.column_first ? p.literal : value;
  switch ((Fn) p.fn) {
    case Fn::STARTSWITH: return left.size() >= right.size() && left.compare(0, right.size(), right) == 0;
    case Fn::ENDSWITH: return left.size() >= right.size() && left.compare(left.size() - right.size(), right.size(), right) == 0;
    default: break;
  }
  const int cmp = strncmp(left.c_str(), right.c_str(), std::min(left.size(), right.size()));
  switch ((Fn) p.fn) {
    case Fn::LT: return cmp < 0 || (cmp == 0 && left.size() < right.size());
    case Fn::LTE: return cmp < 0 || (cmp == 0 && left.size() <= right.size());
    case Fn::GT: return cmp > 0 || (cmp == 0 && left.size() > right.size());
    case Fn::GTE: return cmp > 0 || (cmp == 0 && left.size() >= right.size());
    default: fail(EVQGPU_ERR_UNSUPPORTED, "string predicate %d", p.fn);
  }
}

__global__ void k_lut_apply(const u32* __restrict__ codes, u64 n, const u8* __restrict__ lut, u32* __restrict__ out) {
  for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) out[i] = lut[codes[i]];
}

// The verdict column: a UINT32_PLAIN shadow (0 / 1 per value, the string column's definition levels) that the scan reads
// as a BOOL input.  The predicate runs on the host once per DICTIONARY entry - not per row - and a device pass maps the
// value codes through the verdict table.
const Column* ensure_pred_column(evqgpu_table* t, Column& c, const StringPredicate& p) {
  const std::string key = std::to_string(p.fn) + (p.column_first ? "c" : "l") + (p.invert ? "!" : "=") + p.literal;
  auto it = c.pred_cols.find(key);
  if (it != c.pred_cols.end()) return it->second.get();
  const Column* codes = ensure_code_column(t, c);
  evqgpu_ctx* ctx = t->ctx;
  use_device(ctx);
  const auto& dict = ctx->code_strings;
  std::vector<uint8_t> lut(std::max<size_t>(1, dict.size()));
  for (size_t i = 0; i < dict.size(); ++i) lut[i] = (string_predicate_eval(p, dict[i]) != p.invert) ? 1 : 0;
  DevBuf dlut;
  dlut.alloc(lut.size());
  EVQ_CUDA(cudaMemcpyAsync(dlut.p, lut.data(), lut.size(), cudaMemcpyHostToDevice, ctx->stream));
  const uint64_t nv = codes->num_values;
  std::unique_ptr<Column> sh(new Column());
  sh->meta = c.meta;
  sh->meta.logical_type = EVQ_COL_BOOLEAN;
  sh->meta.encoding = EVQ_ENC_UINT32_PLAIN;
  sh->sql_type = EVQ_BOOL;
  sh->scannable = true;
  sh->data_kind = EVQ_KIND_PLAIN32;
  sh->data.present = true;
  sh->data.nbytes = nv * 4;
  const uint64_t alloc = round_up(nv * 4, 256) + 256;
  sh->data.buf.alloc(alloc);
  const uint64_t tail0 = (nv * 4) & ~255ull;
  EVQ_CUDA(cudaMemsetAsync((uint8_t*) sh->data.buf.p + tail0, 0, alloc - tail0, ctx->stream));
  if (nv) {
    k_lut_apply<<<(unsigned) std::min<uint64_t>((nv + 255) / 256, (uint64_t) ctx->sm_count * 16), 256, 0, ctx->stream>>>(
        codes->data.buf.as<u32>(), nv, dlut.as<u8>(), sh->data.buf.as<u32>());
    EVQ_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
  }
  if (c.meta.dlevel_max > 0) {
    sh->dlevel.present = c.dlevel.present;
    sh->dlevel.nbytes = c.dlevel.nbytes;
    sh->dlevel.bitpack_max = c.dlevel.bitpack_max;
    sh->dlevel.buf.alloc(c.dlevel.buf.bytes);
    EVQ_CUDA(cudaMemcpyAsync(sh->dlevel.buf.p, c.dlevel.buf.p, c.dlevel.buf.bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  EVQ_CUDA(cudaStreamSynchronize(ctx->stream));   // `lut` / `dlut` go away
  table_finish_column(t, *sh);
  const Column* out = sh.get();
  c.pred_cols[key] = std::move(sh);
  return out;
}

static Column& string_column(evqgpu_table* tbl, const char* name) {
  const int ci = tbl->find(name);
  if (ci < 0) fail(EVQGPU_ERR_ARG, "column not found: %s", name);
  Column& c = tbl->cols[ci];
  if (!c.is_string)
    fail(EVQGPU_ERR_UNSUPPORTED, "column '%s' (logical type %u, encoding %u, rlevel_max %u) is not a flat string column", name,
         c.meta.logical_type, c.meta.encoding, c.meta.rlevel_max);
  if (!c.loaded) table_load_column(tbl, c);
  return c;
}

static void exclusive_scan(evqgpu_ctx* ctx, const u64* in, u64* out, uint64_t n) {
  size_t tmp_bytes = 0;
  EVQ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int64_t) n, ctx->stream));
  DevBuf tmp;
  tmp.alloc(tmp_bytes);
  EVQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, (int64_t) n, ctx->stream));
  EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->kernel_launches += 2;
}

// ---- LSM visibility ----------------------------------------------------------------------------------------------------

#define LSM_UPDATE 1u
#define LSM_SKIP 2u
#define LSM_ERR_ID_LENGTH 1u

struct LsmSegmentArgs {
  const u8* id_data;
  const u64* id_start;
  const u32* id_len;
  const u32* id_row_value;   // optional id column, else nullptr
  const u32* upd_words;      // __lsm_is_update, bit-packed (required column)
  const u32* skip_words;     // __lsm_skip column or nullptr
  const u8* skiplist;        // arena skiplist, 1 bit per row (overrides the column) or nullptr
  u32 upd_bits, skip_bits;
  u64 num_rows;
  u64 base;                  // global ordinal of row 0
};

__device__ __forceinline__ u64 mix64(u64 x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

__global__ void k_lsm_gather(LsmSegmentArgs a, u32* __restrict__ ids, u64* __restrict__ keys, u32* __restrict__ ords,
                             u8* __restrict__ flags, u32* __restrict__ status, unsigned long long* __restrict__ live_updates) {
  const u64 r = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = r < a.num_rows;
  // rows that would record their id (update, not skipped): counted per segment for the cursor's needs_filter rule
  bool live = false;
  if (in_range) {
    live = a.upd_bits && vertical_get(a.upd_words, a.upd_bits, r) > 0;
    bool sk = a.skip_words && a.skip_bits && vertical_get(a.skip_words, a.skip_bits, r) > 0;
    if (a.skiplist) sk = (a.skiplist[r >> 3] >> (r & 7)) & 1;
    live = live && !sk;
  }
  const u32 nlive = __popc(__ballot_sync(0xffffffffu, live));
  if ((threadIdx.x & 31) == 0 && nlive) atomicAdd(live_updates, (unsigned long long) nlive);
  if (!in_range) return;
  const u64 g = a.base + r;
  const u32 v = a.id_row_value ? a.id_row_value[r] : (u32) r;
  u32 w[5] = {0, 0, 0, 0, 0};
  if (v == 0xffffffffu || a.id_len[v] != 20u) {
    atomicOr(status, LSM_ERR_ID_LENGTH);   // SHA1Hash(const void*, size_t) raises "invalid SHA1Hash" (util/SHA1.cc:79-85)
  } else {
    const u8* s = a.id_data + a.id_start[v];
#pragma unroll
    for (int k = 0; k < 20; ++k) w[k >> 2] |= (u32) s[k] << (8 * (k & 3));
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) ids[g * 5 + k] = w[k];
  u64 h = mix64(((u64) w[1] << 32 | w[0]) ^ 0x9e3779b97f4a7c15ull);
  h = mix64(h ^ ((u64) w[3] << 32 | w[2]));
  h = mix64(h ^ w[4]);
  u32 f = 0;
  if (a.upd_bits && vertical_get(a.upd_words, a.upd_bits, r) > 0) f |= LSM_UPDATE;   // readBoolean: value > 0
  bool skip = a.skip_words && a.skip_bits && vertical_get(a.skip_words, a.skip_bits, r) > 0;
  if (a.skiplist) skip = (a.skiplist[r >> 3] >> (r & 7)) & 1;   // partition_cursor.cc:181-183: the arena skiplist wins
  if (skip) f |= LSM_SKIP;
  keys[g] = h;
  ords[g] = (u32) g;
  flags[g] = (u8) f;
}

__global__ void k_lsm_resolve(const u64* __restrict__ keys, const u32* __restrict__ ords, const u32* __restrict__ ids,
                              const u8* __restrict__ flags, u64 n, u8* __restrict__ visible) {
  const u64 p = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const u32 g = ords[p];
  bool vis = !(flags[g] & LSM_SKIP);
  if (vis) {
    const u64 h = keys[p];
    const u32* me = ids + (u64) g * 5;
    for (u64 q = p; q-- > 0 && keys[q] == h;) {
      const u32 gq = ords[q];
      if ((flags[gq] & (LSM_UPDATE | LSM_SKIP)) != LSM_UPDATE) continue;
      const u32* o = ids + (u64) gq * 5;
      if (o[0] == me[0] && o[1] == me[1] && o[2] == me[2] && o[3] == me[3] && o[4] == me[4]) { vis = false; break; }
    }
  }
  visible[g] = vis ? 1 : 0;
}

// 8 rows -> one filter byte (LSB first); counts the visible rows
__global__ void k_lsm_pack(const u8* __restrict__ visible, u64 base, u64 num_rows, u8* __restrict__ filter,
                           unsigned long long* __restrict__ count) {
  const u64 b = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  u32 m = 0;
  if (b * 8 < num_rows) {
    const u32 k = (u32) min((u64) 8, num_rows - b * 8);
    for (u32 i = 0; i < k; ++i) m |= (u32) (visible[base + b * 8 + i] & 1u) << i;
    filter[b] = (u8) m;
  }
  u32 c = __popc(m);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (unsigned long long) c);
}

static Column& lsm_bool_column(evqgpu_table* t, const char* name) {
  const int ci = t->find(name);
  if (ci < 0) fail(EVQGPU_ERR_ARG, "column not found: %s", name);
  Column& c = t->cols[ci];
  if (c.data_kind != EVQ_KIND_BITPACK || c.meta.dlevel_max != 0 || c.meta.rlevel_max != 0)
    fail(EVQGPU_ERR_UNSUPPORTED, "column '%s': the visibility filter expects a required bit-packed column", name);
  if (!c.loaded) table_load_column(t, c);
  return c;
}

}  // namespace evq

using namespace evq;

extern "C" {

int evqgpu_table_decode_string_column(evqgpu_table* tbl, const char* column, uint64_t row0, uint64_t nrows, void* dst,
                                      uint64_t cap, uint64_t* nbytes_out) {
  return guarded([&] {
    if (!tbl || !column || !nbytes_out) fail(EVQGPU_ERR_ARG, "evqgpu_table_decode_string_column: null argument");
    evqgpu_ctx* ctx = tbl->ctx;
    use_device(ctx);
    Column& c = string_column(tbl, column);
    if (row0 > tbl->num_rows) row0 = tbl->num_rows;
    nrows = std::min<uint64_t>(nrows, tbl->num_rows - row0);
    *nbytes_out = 0;
    if (nrows == 0) return;
    const u32* rv = c.meta.dlevel_max > 0 ? c.row_value.as<u32>() : nullptr;
    DevBuf sizes, offs;
    sizes.alloc((nrows + 1) * 8);
    offs.alloc((nrows + 1) * 8);
    k_str_sizes<<<(unsigned) ((nrows + 1 + 255) / 256), 256, 0, ctx->stream>>>(rv, c.str_len.as<u32>(), row0, nrows, sizes.as<u64>());
    EVQ_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
    exclusive_scan(ctx, sizes.as<u64>(), offs.as<u64>(), nrows + 1);
    u64 total = 0;
    EVQ_CUDA(cudaMemcpyAsync(&total, offs.as<u64>() + nrows, 8, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
    *nbytes_out = total;
    if (!dst || cap < total) return;
    DevBuf out;
    out.alloc(total);
    k_str_emit<<<(unsigned) ((nrows * 32 + 255) / 256), 256, 0, ctx->stream>>>(c.data.buf.as<u8>(), c.str_start.as<u64>(), c.str_len.as<u32>(),
                                                                             rv, row0, nrows, offs.as<u64>(), out.as<u8>());
    EVQ_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
    EVQ_CUDA(cudaMemcpyAsync(dst, out.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int evqgpu_table_get_filter(evqgpu_table* tbl, void* bits, uint64_t cap_bytes, int* has_filter_out) {
  return guarded([&] {
    if (!tbl || !has_filter_out) fail(EVQGPU_ERR_ARG, "evqgpu_table_get_filter: null argument");
    *has_filter_out = tbl->has_filter ? 1 : 0;
    if (!tbl->has_filter || !bits) return;
    const uint64_t nbytes = (tbl->num_rows + 7) / 8;
    if (cap_bytes < nbytes) fail(EVQGPU_ERR_ARG, "evqgpu_table_get_filter: buffer too small");
    use_device(tbl->ctx);
    if (nbytes) EVQ_CUDA(cudaMemcpyAsync(bits, tbl->filter.p, nbytes, cudaMemcpyDeviceToHost, tbl->ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(tbl->ctx->stream));
  });
}

int evqgpu_lsm_build_filters(evqgpu_ctx* ctx, evqgpu_lsm_segment* segs, uint32_t nsegs) {
  return guarded([&] {
    if (!ctx || (!segs && nsegs)) fail(EVQGPU_ERR_ARG, "evqgpu_lsm_build_filters: null argument");
    use_device(ctx);
    uint64_t total = 0;
    for (uint32_t i = 0; i < nsegs; ++i) {
      if (!segs[i].table) fail(EVQGPU_ERR_ARG, "evqgpu_lsm_build_filters: segment %u has no table", i);
      if (segs[i].table->ctx != ctx) fail(EVQGPU_ERR_ARG, "evqgpu_lsm_build_filters: segment %u lives on another context", i);
      segs[i].visible_rows = segs[i].table->num_rows;
      segs[i].filtered = 0;
      if (!(segs[i].flags & EVQGPU_LSM_NO_FILTER)) total += segs[i].table->num_rows;
    }
    if (total >= 0xffffffffull) fail(EVQGPU_ERR_UNSUPPORTED, "visibility filter over more than 2^32 - 1 rows");
    DevBuf ids, keys_a, keys_b, ords_a, ords_b, flags, visible, status, live, errs, tmp;
    ids.alloc(total * 20);
    keys_a.alloc(total * 8);
    keys_b.alloc(total * 8);
    ords_a.alloc(total * 4);
    ords_b.alloc(total * 4);
    flags.alloc(total);
    visible.alloc(total);
    status.alloc(16);   // [8] visible-row counter (u64)
    live.alloc((uint64_t) (nsegs + 1) * 8);   // per segment: rows that record their id (update, not skipped)
    errs.alloc((uint64_t) (nsegs + 1) * 4);   // per segment: error bits
    EVQ_CUDA(cudaMemsetAsync(status.p, 0, 16, ctx->stream));
    EVQ_CUDA(cudaMemsetAsync(errs.p, 0, (uint64_t) (nsegs + 1) * 4, ctx->stream));
    EVQ_CUDA(cudaMemsetAsync(live.p, 0, (uint64_t) (nsegs + 1) * 8, ctx->stream));
    std::vector<DevBuf> skiplists(nsegs);
    std::vector<uint64_t> bases(nsegs, 0);
    uint64_t base = 0;
    for (uint32_t i = 0; i < nsegs; ++i) {
      evqgpu_table* t = segs[i].table;
      if (segs[i].flags & EVQGPU_LSM_NO_FILTER) continue;
      bases[i] = base;
      if (t->num_rows == 0) continue;
      Column& id = string_column(t, "__lsm_id");
      Column& upd = lsm_bool_column(t, "__lsm_is_update");
      LsmSegmentArgs a;
      memset(&a, 0, sizeof(a));
      a.id_data = id.data.buf.as<u8>();
      a.id_start = id.str_start.as<u64>();
      a.id_len = id.str_len.as<u32>();
      a.id_row_value = id.meta.dlevel_max > 0 ? id.row_value.as<u32>() : nullptr;
      a.upd_words = upd.data.buf.as<u32>();
      a.upd_bits = upd.data_bits;
      if (segs[i].flags & EVQGPU_LSM_SKIP_COLUMN) {
        Column& sk = lsm_bool_column(t, "__lsm_skip");
        a.skip_words = sk.data.buf.as<u32>();
        a.skip_bits = sk.data_bits;
      }
      if (segs[i].skiplist) {
        const uint64_t nb = (t->num_rows + 7) / 8;
        skiplists[i].alloc(nb);
        EVQ_CUDA(cudaMemcpyAsync(skiplists[i].p, segs[i].skiplist, nb, cudaMemcpyHostToDevice, ctx->stream));
        a.skiplist = skiplists[i].as<u8>();
      }
      a.num_rows = t->num_rows;
      a.base = base;
      k_lsm_gather<<<(unsigned) ((t->num_rows + 255) / 256), 256, 0, ctx->stream>>>(a, ids.as<u32>(), keys_a.as<u64>(), ords_a.as<u32>(),
                                                                                  flags.as<u8>(), errs.as<u32>() + i,
                                                                                  live.as<unsigned long long>() + i);
      EVQ_CUDA(cudaGetLastError());
      ctx->kernel_launches++;
      base += t->num_rows;
    }
    // which segments the cursor filters (partition_cursor.cc:139-155): with EVQGPU_LSM_AUTO a table without a skiplist is
    // scanned unfiltered - and records no ids - when no id was recorded before it and it is the partition's oldest table or
    // has no updates.  "No id recorded yet" == no filtered segment before it holds a non-skipped update row.
    std::vector<uint64_t> live_h(nsegs + 1, 0);
    std::vector<u32> errs_h(nsegs + 1, 0);
    EVQ_CUDA(cudaMemcpyAsync(live_h.data(), live.p, (uint64_t) (nsegs + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaMemcpyAsync(errs_h.data(), errs.p, (uint64_t) (nsegs + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<char> filtered(nsegs, 0);
    bool id_set_empty = true;
    for (uint32_t i = 0; i < nsegs; ++i) {
      const uint32_t f = segs[i].flags;
      if (f & EVQGPU_LSM_NO_FILTER) continue;
      bool needs = true;
      if (f & EVQGPU_LSM_AUTO) {
        const bool has_skiplist = (f & EVQGPU_LSM_SKIP_COLUMN) || segs[i].skiplist;
        if (!has_skiplist && (f & EVQGPU_LSM_OLDEST) && id_set_empty) needs = false;
        if (!has_skiplist && !(f & EVQGPU_LSM_HAS_UPDATES) && id_set_empty) needs = false;
      }
      filtered[i] = needs;
      if (needs && live_h[i] > 0) id_set_empty = false;
      if (!needs && segs[i].table->num_rows)   // its rows take no part: as if skipped
        EVQ_CUDA(cudaMemsetAsync(flags.as<u8>() + bases[i], LSM_SKIP, segs[i].table->num_rows, ctx->stream));
    }
    // the reference converts every id it reads; the ids of a segment it does not filter are never read
    for (uint32_t i = 0; i < nsegs; ++i)
      if (filtered[i] && (errs_h[i] & LSM_ERR_ID_LENGTH)) fail(EVQGPU_ERR_RUNTIME, "invalid SHA1Hash");
    if (total) {
      size_t tmp_bytes = 0;
      EVQ_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_a.as<u64>(), keys_b.as<u64>(), ords_a.as<u32>(), ords_b.as<u32>(),
                                               (int64_t) total, 0, 64, ctx->stream));
      tmp.alloc(tmp_bytes);
      EVQ_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_a.as<u64>(), keys_b.as<u64>(), ords_a.as<u32>(), ords_b.as<u32>(),
                                               (int64_t) total, 0, 64, ctx->stream));
      ctx->kernel_launches += 8;
      k_lsm_resolve<<<(unsigned) ((total + 255) / 256), 256, 0, ctx->stream>>>(keys_b.as<u64>(), ords_b.as<u32>(), ids.as<u32>(),
                                                                             flags.as<u8>(), total, visible.as<u8>());
      EVQ_CUDA(cudaGetLastError());
      ctx->kernel_launches++;
    }
    for (uint32_t i = 0; i < nsegs; ++i) {
      evqgpu_table* t = segs[i].table;
      if (!filtered[i]) {   // partition_cursor.cc:216-218: no setFilter call
        t->filter.release();
        t->has_filter = false;
        continue;
      }
      t->filter.alloc(round_up((uint64_t) t->num_tiles * (EVQ_TILE_ROWS / 8), 256) + 256);
      EVQ_CUDA(cudaMemsetAsync(t->filter.p, 0, t->filter.bytes, ctx->stream));
      unsigned long long cnt = 0;
      if (t->num_rows) {
        EVQ_CUDA(cudaMemsetAsync((u8*) status.p + 8, 0, 8, ctx->stream));
        const uint64_t nb = (t->num_rows + 7) / 8;
        k_lsm_pack<<<(unsigned) ((nb + 255) / 256), 256, 0, ctx->stream>>>(visible.as<u8>(), bases[i], t->num_rows, t->filter.as<u8>(),
                                                                         (unsigned long long*) ((u8*) status.p + 8));
        EVQ_CUDA(cudaGetLastError());
        ctx->kernel_launches++;
        EVQ_CUDA(cudaMemcpyAsync(&cnt, (u8*) status.p + 8, 8, cudaMemcpyDeviceToHost, ctx->stream));
      }
      EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
      segs[i].visible_rows = cnt;
      segs[i].filtered = 1;
      t->has_filter = true;
    }
  });
}

}  // extern "C"
