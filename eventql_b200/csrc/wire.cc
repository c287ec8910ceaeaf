// wire.cc - evqgpu_query_fetch_partial: the groups of an aggregate plan as the rows the reference's
// PartialGroupByExpression::nextBatch produces (sql/statements/select/groupby.cc:411-445), i.e. what a shard of a cluster
// query sends to GroupByMergeExpression (groupby.cc:553-615) and writes to its query cache (groupby.cc:379-405).
//
// The device side (evq_emit with EVQGPU_QUERY_WIRE, csrc/codegen.cc) leaves three things per group in HBM: the packed
// result columns, the SHA-1 of the key tuple and the raw aggregate state words.  Here they are copied to the host and
// spelled in the reference's serialisation: varuints (util/io/outputstream.cc appendVarUInt) for count / sum, the raw
// structs of the extension aggregates (oracle/ref_tools/ext_aggregates.cc), SValue::encode (svalue.cc:306-309) for items
// that are not aggregates.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "query.h"

namespace evq {
void finish_query(evqgpu_query& q);

static void put_varuint(std::vector<uint8_t>& out, uint64_t v) {   // OutputStream::appendVarUInt: LEB128
  do {
    uint8_t b = v & 0x7f;
    v >>= 7;
    if (v) b |= 0x80;
    out.push_back(b);
  } while (v);
}

static void put_raw(std::vector<uint8_t>& out, const void* p, size_t n) {
  const uint8_t* b = (const uint8_t*) p;
  out.insert(out.end(), b, b + n);
}

}  // namespace evq

using namespace evq;

extern "C" int evqgpu_query_fetch_partial(evqgpu_query* q, uint64_t row0, uint64_t max_rows, void* keys, void* data, uint64_t data_cap,
                                          uint64_t* data_offsets, uint64_t* nrows_out, uint64_t* data_bytes_out) {
  return guarded([&] {
    if (!q || !keys || !data_offsets || !nrows_out || !data_bytes_out) fail(EVQGPU_ERR_ARG, "evqgpu_query_fetch_partial: null argument");
    if (!(q->flags & EVQGPU_QUERY_WIRE) || !(q->flags & EVQGPU_QUERY_GROUPBY))
      fail(EVQGPU_ERR_ARG, "evqgpu_query_fetch_partial: the plan was not created with EVQGPU_QUERY_GROUPBY | EVQGPU_QUERY_WIRE");
    if (q->pending) finish_query(*q);
    if (!q->emitted) fail(EVQGPU_ERR_ARG, "evqgpu_query_fetch_partial: the partial results of a multi-rank job are merged, not fetched");
    if (q->reordered) fail(EVQGPU_ERR_ARG, "evqgpu_query_fetch_partial after ORDER BY / LIMIT: the key hashes no longer match the rows");
    use_device(q->ctx);
    uint64_t n = 0;
    if (row0 < q->num_rows_out) n = std::min<uint64_t>(max_rows, q->num_rows_out - row0);
    *nrows_out = n;
    data_offsets[0] = 0;
    *data_bytes_out = 0;
    if (n == 0) return;
    const size_t nstate = std::max<size_t>(1, q->state_ops.size());
    std::vector<uint64_t> st(n * nstate);
    std::vector<std::vector<uint8_t>> cols(q->select.size());
    cudaStream_t s = q->ctx->stream;
    const size_t kstride = wire_key_stride(*q);
    std::vector<uint8_t> rawkeys;
    if (q->string_keys) {
      rawkeys.resize(n * kstride);
      EVQ_CUDA(cudaMemcpyAsync(rawkeys.data(), q->out_sha.as<u8>() + row0 * kstride, n * kstride, cudaMemcpyDeviceToHost, s));
    } else {
      EVQ_CUDA(cudaMemcpyAsync(keys, q->out_sha.as<u8>() + row0 * 20, n * 20, cudaMemcpyDeviceToHost, s));
    }
    EVQ_CUDA(cudaMemcpyAsync(st.data(), q->out_state.as<u64>() + row0 * nstate, n * nstate * 8, cudaMemcpyDeviceToHost, s));
    for (size_t i = 0; i < q->select.size(); ++i) {
      if (q->select[i].agg) continue;
      const uint64_t w = q->select[i].expr->type == EVQ_BOOL ? 2 : 9;
      cols[i].resize(n * w);
      EVQ_CUDA(cudaMemcpyAsync(cols[i].data(), q->out_cols[i].as<u8>() + row0 * w, n * w, cudaMemcpyDeviceToHost, s));
    }
    EVQ_CUDA(cudaStreamSynchronize(s));
    const auto& dict = q->ctx->code_strings;
    // a string as it lies on the VM stack and in an SVector: [u32 length][bytes][tag] (svalue.cc:1139-1177); a NULL has length 0
    auto put_string = [&](std::vector<uint8_t>& b, uint64_t code, uint8_t tag) {
      if (code >= dict.size()) fail(EVQGPU_ERR_RUNTIME, "string code %llu outside the dictionary", (unsigned long long) code);
      const std::string empty;
      const std::string& str = (tag & EVQ_STAG_NULL) ? empty : dict[code];
      const uint32_t len = (uint32_t) str.size();
      put_raw(b, &len, 4);
      put_raw(b, str.data(), len);
      b.push_back(tag);
    };
    // count_distinct: the saved state is the set (aggregate.cc:110-116: its size, then its members in std::set order).  The
    // sets live in the (group id, value) tables of the scan; the emit kernel left every group's id in the state word
    std::vector<std::unordered_map<uint64_t, std::vector<uint64_t>>> sets(q->distinct_args.size());
    for (size_t d = 0; d < q->distinct_args.size(); ++d) {
      std::vector<uint64_t> tab(q->dt_cap * 4);
      EVQ_CUDA(cudaMemcpyAsync(tab.data(), q->dt_slots[d].p, tab.size() * 8, cudaMemcpyDeviceToHost, s));
      EVQ_CUDA(cudaStreamSynchronize(s));
      for (uint64_t i = 0; i < q->dt_cap; ++i)
        if (tab[i * 4]) sets[d][tab[i * 4 + 1]].push_back(tab[i * 4 + 2]);
      for (auto& kv : sets[d]) std::sort(kv.second.begin(), kv.second.end());
    }
    if (q->string_keys) {
      // the group key: SHA-1 of the group expressions' stack bytes, last expression first (groupby.cc:112-135); the device sent
      // the tuple with dictionary codes in place of the strings
      std::vector<uint8_t> msg;
      for (uint64_t r = 0; r < n; ++r) {
        const uint8_t* raw = &rawkeys[r * kstride];
        msg.clear();
        for (size_t i = q->group.size(); i-- > 0;) {
          if (q->group[i]->type == EVQ_BOOL) {
            put_raw(msg, raw, 2);
            raw += 2;
            continue;
          }
          if (i < q->group_is_string.size() && q->group_is_string[i]) {
            uint64_t code;
            memcpy(&code, raw, 8);
            put_string(msg, code, raw[8]);
          } else {
            put_raw(msg, raw, 9);
          }
          raw += 9;
        }
        sha1(msg.data(), msg.size(), (uint8_t*) keys + r * 20);
      }
    }
    std::vector<uint8_t> out;
    out.reserve(n * 32);
    for (uint64_t r = 0; r < n; ++r) {
      const uint64_t* g = &st[r * nstate];
      for (size_t i = 0; i < q->select.size(); ++i) {
        const SelectItem& item = q->select[i];
        if (!item.agg && item.is_string) {
          // SValue::encode of a string: type, length, [u32 length][bytes][tag].  A packed value of exactly
          // SValue::kInlineDataSize = 16 bytes carries STAG_INLINE in its tag byte: the tag of the packed string and the
          // SValue's own flag byte are the same byte then (svalue.cc:346-365, svalue.h:55,66) - reproduced, it is on the wire
          uint64_t code;
          memcpy(&code, &cols[i][r * 9], 8);
          std::vector<uint8_t> packed;
          put_string(packed, code, cols[i][r * 9 + 8]);
          if (packed.size() == 16) packed.back() |= 128;
          out.push_back((uint8_t) EVQ_STRING);
          put_varuint(out, packed.size());
          put_raw(out, packed.data(), packed.size());
          continue;
        }
        if (!item.agg) {   // SValue::encode: type, length, packed value
          const uint64_t w = item.expr->type == EVQ_BOOL ? 2 : 9;
          out.push_back((uint8_t) item.expr->type);
          put_varuint(out, w);
          put_raw(out, &cols[i][r * w], w);
          continue;
        }
        const FnInfo& fi = item.agg->info();
        const uint64_t s0 = item.state0 >= 0 ? g[item.state0] : 0;
        const uint64_t seen = item.state_seen >= 0 ? g[item.state_seen] : 0;
        switch (fi.fn) {
          case Fn::COUNT: put_varuint(out, g[0]); break;                               // count_save
          case Fn::SUM:
            if (fi.ret == EVQ_FLOAT64) put_raw(out, &s0, 8);                           // SumF64::save
            else put_varuint(out, s0);                                                 // sum_uint64_save / sum_int64_save
            break;
          case Fn::MIN:
          case Fn::MAX: {                                                              // MinMaxState {value, seen}
            const uint64_t have = seen ? 1 : 0, v = have ? s0 : 0;
            put_raw(out, &v, 8);
            put_raw(out, &have, 8);
            break;
          }
          case Fn::MEAN: {                                                             // MeanState {double sum, n}
            double sum;
            if (item.state_carry >= 0) sum = (double) g[item.state_carry] * 18446744073709551616.0 + (double) s0;
            else memcpy(&sum, &s0, 8);
            put_raw(out, &sum, 8);
            put_raw(out, &seen, 8);
            break;
          }
          case Fn::COUNT_DISTINCT: {                                                   // count_distinct_uint64_save
            static const std::vector<uint64_t> none;
            const auto& set = sets[(size_t) item.distinct];
            const auto it = set.find(s0);   // (s0 = the group's id, see above)
            const std::vector<uint64_t>& members = it == set.end() ? none : it->second;
            put_varuint(out, members.size());
            for (uint64_t v : members) put_varuint(out, v);
            break;
          }
          default: fail(EVQGPU_ERR_UNSUPPORTED, "aggregate %s has no partial state format", fi.symbol.c_str());
        }
      }
      data_offsets[r + 1] = out.size();
    }
    *data_bytes_out = out.size();
    if (data && out.size() <= data_cap) memcpy(data, out.data(), out.size());
  });
}

// ---- the query cache entry of a partial aggregation -------------------------------------------------------------------
// PartialGroupByExpression::execute stores its groups under QueryCache (sql/runtime/query_cache.cc:58-75) as
//   u8 0x01 | u64 number of groups | per group: 20-byte group key | the saved states of the select items
// (sql/statements/select/groupby.cc:411-432; read back by :262-292) in the file <cache dir>/<key>.qc, where
//   key = SHA1(hex(input cache key) + hex(expression fingerprint))            (groupby.cc:474-483, util/SHA1.cc:87-89).
// The per-group bytes are exactly the `data` slices evqgpu_query_fetch_partial produces.

extern "C" int evqgpu_partial_cache_encode(const void* keys, const void* data, const uint64_t* data_offsets, uint64_t ngroups,
                                           void* dst, uint64_t cap, uint64_t* nbytes_out) {
  return guarded([&] {
    if (!nbytes_out || (ngroups && (!keys || !data_offsets))) fail(EVQGPU_ERR_ARG, "evqgpu_partial_cache_encode: null argument");
    uint64_t need = 9 + 20 * ngroups;
    for (uint64_t i = 0; i < ngroups; ++i) {
      if (data_offsets[i + 1] < data_offsets[i]) fail(EVQGPU_ERR_ARG, "evqgpu_partial_cache_encode: data offsets must ascend");
      need += data_offsets[i + 1] - data_offsets[i];
    }
    *nbytes_out = need;
    if (!dst || cap < need) return;
    if (ngroups && !data && data_offsets[ngroups] != data_offsets[0]) fail(EVQGPU_ERR_ARG, "evqgpu_partial_cache_encode: null data");
    uint8_t* o = (uint8_t*) dst;
    *o++ = 0x01;
    memcpy(o, &ngroups, 8);
    o += 8;
    for (uint64_t i = 0; i < ngroups; ++i) {
      memcpy(o, (const uint8_t*) keys + 20 * i, 20);
      o += 20;
      const uint64_t n = data_offsets[i + 1] - data_offsets[i];
      if (n) memcpy(o, (const uint8_t*) data + data_offsets[i], n);
      o += n;
    }
  });
}

extern "C" int evqgpu_partial_cache_filename(const void* input_cache_key, const void* expression_fingerprint, char* out, uint64_t cap) {
  return guarded([&] {
    if (!input_cache_key || !expression_fingerprint || !out || cap < 44) fail(EVQGPU_ERR_ARG, "evqgpu_partial_cache_filename: bad argument");
    static const char* digits = "0123456789abcdef";
    char text[80];
    const uint8_t* parts[2] = {(const uint8_t*) input_cache_key, (const uint8_t*) expression_fingerprint};
    for (int p = 0; p < 2; ++p)
      for (int i = 0; i < 20; ++i) {
        text[40 * p + 2 * i] = digits[parts[p][i] >> 4];
        text[40 * p + 2 * i + 1] = digits[parts[p][i] & 15];
      }
    uint8_t h[20];
    sha1((const uint8_t*) text, 80, h);
    for (int i = 0; i < 20; ++i) {
      out[2 * i] = digits[h[i] >> 4];
      out[2 * i + 1] = digits[h[i] & 15];
    }
    memcpy(out + 40, ".qc", 4);
  });
}

extern "C" int evqgpu_query_store_cache(evqgpu_query* q, const char* path) {
  return guarded([&] {
    if (!q || !path) fail(EVQGPU_ERR_ARG, "evqgpu_query_store_cache: null argument");
    uint64_t n = 0;
    if (evqgpu_query_num_rows(q, &n) != EVQGPU_OK) throw Error{EVQGPU_ERR_RUNTIME, last_error()};
    std::vector<uint8_t> keys(20 * n + 1);
    std::vector<uint64_t> offs(n + 1, 0);
    uint64_t got = 0, need = 0;
    int rc = evqgpu_query_fetch_partial(q, 0, n, keys.data(), nullptr, 0, offs.data(), &got, &need);
    if (rc != EVQGPU_OK) throw Error{rc, last_error()};
    std::vector<uint8_t> data(need + 1);
    rc = evqgpu_query_fetch_partial(q, 0, n, keys.data(), data.data(), need, offs.data(), &got, &need);
    if (rc != EVQGPU_OK) throw Error{rc, last_error()};
    if (got != n) fail(EVQGPU_ERR_RUNTIME, "evqgpu_query_store_cache: fetched %llu of %llu groups", (unsigned long long) got, (unsigned long long) n);
    uint64_t bytes = 0;
    evqgpu_partial_cache_encode(keys.data(), data.data(), offs.data(), n, nullptr, 0, &bytes);
    std::vector<uint8_t> entry(bytes);
    rc = evqgpu_partial_cache_encode(keys.data(), data.data(), offs.data(), n, entry.data(), bytes, &bytes);
    if (rc != EVQGPU_OK) throw Error{rc, last_error()};
    // QueryCache::storeEntry: write beside the target, then rename (query_cache.cc:64-74)
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) fail(EVQGPU_ERR_ARG, "evqgpu_query_store_cache: cannot create %s", tmp.c_str());
    const bool ok = fwrite(entry.data(), 1, entry.size(), f) == entry.size();
    if (fclose(f) != 0 || !ok) { remove(tmp.c_str()); fail(EVQGPU_ERR_RUNTIME, "evqgpu_query_store_cache: short write to %s", tmp.c_str()); }
    if (rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); fail(EVQGPU_ERR_RUNTIME, "evqgpu_query_store_cache: cannot rename to %s", path); }
  });
}

// ---- QUERY_PARTIALAGGR_RESULT frames ------------------------------------------------------------------------------------
// What a shard answers a coordinator's QUERY_PARTIALAGGR with (transport/native/ops/query_partialaggr.cc:83-124): frames of
//   8-byte header: u16 opcode 0x0102 | u16 flags (EVQL_ENDOFREQUEST = 1 on the last) | u32 payload length, big endian
//                  (TCPConnection::writeFrameHeaderAsync, transport/native/connection_tcp.cc:238-251)
//   payload:       varuint flags (0) | varuint number of rows | body  (QueryPartialAggrResultFrame::writeTo,
//                  transport/native/frames/query_partialaggr_result.cc:56-60)
//   body:          per row the 20-byte group key and the saved states, back to back (appendString writes raw bytes)
// A frame is closed once its body exceeds the soft maximum (8 MiB in the reference); if that happens on the last row an
// empty frame carries the end-of-request flag, as the reference's loop does.
extern "C" int evqgpu_partial_frames_encode(const void* keys, const void* data, const uint64_t* data_offsets, uint64_t ngroups,
                                            uint64_t soft_max_body, void* dst, uint64_t cap, uint64_t* nbytes_out,
                                            uint64_t* nframes_out) {
  return guarded([&] {
    if (!nbytes_out || (ngroups && (!keys || !data_offsets))) fail(EVQGPU_ERR_ARG, "evqgpu_partial_frames_encode: null argument");
    if (soft_max_body == 0) soft_max_body = 8ull << 20;
    std::vector<uint8_t> out;
    uint64_t frames = 0, i = 0;
    for (bool eof = false; !eof;) {
      std::vector<uint8_t> body;
      uint64_t num_rows = 0;
      while ((eof = (i >= ngroups)) == false) {
        ++num_rows;
        put_raw(body, (const uint8_t*) keys + 20 * i, 20);
        if (data_offsets[i + 1] < data_offsets[i]) fail(EVQGPU_ERR_ARG, "evqgpu_partial_frames_encode: data offsets must ascend");
        if (data_offsets[i + 1] > data_offsets[i]) {
          if (!data) fail(EVQGPU_ERR_ARG, "evqgpu_partial_frames_encode: null data");
          put_raw(body, (const uint8_t*) data + data_offsets[i], data_offsets[i + 1] - data_offsets[i]);
        }
        ++i;
        if (body.size() > soft_max_body) break;
      }
      std::vector<uint8_t> payload;
      put_varuint(payload, 0);
      put_varuint(payload, num_rows);
      payload.insert(payload.end(), body.begin(), body.end());
      if (payload.size() > 0xffffffffull) fail(EVQGPU_ERR_UNSUPPORTED, "evqgpu_partial_frames_encode: frame payload over 4 GiB");
      const uint16_t opcode = 0x0102, flags = eof ? 1 : 0;
      const uint32_t len = (uint32_t) payload.size();
      const uint8_t hdr[8] = {(uint8_t) (opcode >> 8), (uint8_t) opcode, (uint8_t) (flags >> 8), (uint8_t) flags,
                              (uint8_t) (len >> 24), (uint8_t) (len >> 16), (uint8_t) (len >> 8), (uint8_t) len};
      put_raw(out, hdr, 8);
      out.insert(out.end(), payload.begin(), payload.end());
      ++frames;
    }
    *nbytes_out = out.size();
    if (nframes_out) *nframes_out = frames;
    if (dst && cap >= out.size()) memcpy(dst, out.data(), out.size());
  });
}

// ---- reading partial rows back (the coordinator's side of the formats above) --------------------------------------------
// GroupByMergeExpression (groupby.cc:553-615) and the cache load of PartialGroupByExpression (groupby.cc:262-292) walk a
// body of `20-byte key | saved states` rows by loading every select item's state in turn (VM::loadInstanceState /
// SValue::decode): the length of a row is only known from the plan.  evqgpu_partial_rows_split does that walk for a plan
// description and returns where every row starts; the containers (.qc entry, result frames) are unwrapped by the two
// functions below it.  Pure host functions: parsing only, no aggregation happens here.
namespace evq {
void query_intake(evqgpu_query* q, const evqgpu_query_desc* desc);

static uint64_t get_varuint(const uint8_t* p, uint64_t n, uint64_t& pos) {
  uint64_t v = 0;
  for (int k = 0;; ++k) {
    if (pos >= n) fail(EVQGPU_ERR_FORMAT, "partial rows: truncated varuint");
    const uint8_t b = p[pos++];
    if (k < 10) v |= (uint64_t) (b & 0x7f) << (7 * k);
    if (!(b & 0x80)) return v;
  }
}

// length of one row's saved states starting at p[pos]
static void skip_states(const evqgpu_query& q, const uint8_t* p, uint64_t n, uint64_t& pos) {
  for (const auto& item : q.select) {
    if (!item.agg) {   // SValue::encode (svalue.cc:306-309): type byte, varuint length, packed value
      if (pos >= n) fail(EVQGPU_ERR_FORMAT, "partial rows: truncated value");
      ++pos;
      const uint64_t len = get_varuint(p, n, pos);
      if (len > n - pos) fail(EVQGPU_ERR_FORMAT, "partial rows: truncated value");
      pos += len;
      continue;
    }
    const FnInfo& fi = item.agg->info();
    uint64_t fixed = 0;
    switch (fi.fn) {
      case Fn::COUNT: get_varuint(p, n, pos); break;
      case Fn::COUNT_DISTINCT: {   // the set: its size, then its members
        const uint64_t members = get_varuint(p, n, pos);
        for (uint64_t i = 0; i < members; ++i) get_varuint(p, n, pos);
        break;
      }
      case Fn::SUM:
        if (fi.ret == EVQ_FLOAT64) fixed = 8;
        else get_varuint(p, n, pos);
        break;
      case Fn::MIN:
      case Fn::MAX:
      case Fn::MEAN: fixed = 16; break;
      default: fail(EVQGPU_ERR_UNSUPPORTED, "aggregate %s has no partial state format", fi.symbol.c_str());
    }
    if (fixed > n - pos) fail(EVQGPU_ERR_FORMAT, "partial rows: truncated state");
    pos += fixed;
  }
}

// one shard row -> one merge record [3 key words][tag word][state words in the coordinator layout]: loadInstanceState of
// every select item (groupby.cc:577-612), SValue::decode for the others (svalue.cc:311-315)
static void parse_row(evqgpu_query& q, const uint8_t* p, uint64_t n, uint64_t arrival, uint64_t* rec) {
  const size_t nstate = q.state_ops.size();
  if (n < 20) fail(EVQGPU_ERR_FORMAT, "partial rows: truncated group key");
  rec[0] = rec[1] = rec[2] = rec[3] = 0;
  memcpy(&rec[0], p, 8);
  memcpy(&rec[1], p + 8, 8);
  memcpy(&rec[2], p + 16, 4);
  uint64_t* st = rec + 4;
  for (size_t w = 0; w < nstate; ++w) {   // the merge kernel skips identities
    switch (q.state_ops[w]) {
      case OP_MIN_U64: case OP_FIRST_ORD: st[w] = ~0ull; break;
      case OP_MIN_I64: st[w] = 0x7fffffffffffffffull; break;
      case OP_MAX_I64: st[w] = 0x8000000000000000ull; break;
      case OP_MIN_F64: st[w] = 0x7ff0000000000000ull; break;
      case OP_MAX_F64: st[w] = 0xfff0000000000000ull; break;
      default: st[w] = 0; break;
    }
  }
  uint64_t pos = 20;
  auto raw8 = [&]() -> uint64_t {
    if (n - pos < 8) fail(EVQGPU_ERR_FORMAT, "partial rows: truncated state");
    uint64_t v;
    memcpy(&v, p + pos, 8);
    pos += 8;
    return v;
  };
  for (const auto& item : q.select) {
    if (!item.agg) {
      if (pos >= n) fail(EVQGPU_ERR_FORMAT, "partial rows: truncated value");
      const uint8_t type = p[pos++];
      const uint64_t len = get_varuint(p, n, pos);
      if (len > n - pos) fail(EVQGPU_ERR_FORMAT, "partial rows: truncated value");
      if (item.is_string) {
        // a string value: [u32 length][bytes][tag] (the tag may carry STAG_INLINE, svalue.cc:346-365); it enters the
        // context's dictionary and travels on as its code, like a string column's value
        uint32_t slen = 0;
        if (type != (uint8_t) EVQ_STRING || len < 5) fail(EVQGPU_ERR_FORMAT, "partial rows: a value of type %u where the plan has a string", type);
        memcpy(&slen, p + pos, 4);
        if ((uint64_t) slen + 5 != len) fail(EVQGPU_ERR_FORMAT, "partial rows: string of %u bytes in a value of %llu", slen, (unsigned long long) len);
        const uint64_t tag = p[pos + len - 1] & 1u;
        const uint64_t code = tag ? 0 : string_code(q.ctx, std::string((const char*) p + pos + 4, slen));
        pos += len;
        st[item.state_first] = (arrival << 1) | tag;
        st[item.state_first + 1] = code;
        continue;
      }
      const uint64_t want = item.expr->type == EVQ_BOOL ? 2 : 9;
      if (type != (uint8_t) item.expr->type || len != want)
        fail(EVQGPU_ERR_FORMAT, "partial rows: a value of type %u / %llu bytes where the plan has type %u", type, (unsigned long long) len, item.expr->type);
      uint64_t v = 0;
      memcpy(&v, p + pos, want - 1);
      const uint64_t tag = p[pos + want - 1] & 1u;
      pos += len;
      st[item.state_first] = (arrival << 1) | tag;
      st[item.state_first + 1] = tag ? 0 : v;
      continue;
    }
    const FnInfo& fi = item.agg->info();
    switch (fi.fn) {
      case Fn::COUNT: st[item.state0] = get_varuint(p, n, pos); break;
      case Fn::SUM: st[item.state0] = fi.ret == EVQ_FLOAT64 ? raw8() : get_varuint(p, n, pos); break;
      case Fn::MIN:
      case Fn::MAX: {
        const uint64_t v = raw8(), have = raw8();
        if (have) { st[item.state0] = v; st[item.state_seen] = 1; }
        break;
      }
      case Fn::MEAN: st[item.state0] = raw8(); st[item.state_seen] = raw8(); break;
      case Fn::COUNT_DISTINCT: {
        // count_distinct_uint64_load (aggregate.cc:118-124): the shard's set.  Its members are kept as (group key, value)
        // pairs; merge_finish unites them on the device and counts every group's distinct pairs into this (zero) word
        if (q.coord_pairs.size() < q.distinct_args.size()) q.coord_pairs.resize(q.distinct_args.size());
        std::vector<uint64_t>& pairs = q.coord_pairs[(size_t) item.distinct];
        const uint64_t members = get_varuint(p, n, pos);
        if (members > n - pos) fail(EVQGPU_ERR_FORMAT, "partial rows: a set of %llu members in %llu bytes", (unsigned long long) members, (unsigned long long) (n - pos));
        for (uint64_t i = 0; i < members; ++i) {
          const uint64_t v = get_varuint(p, n, pos);
          pairs.insert(pairs.end(), {rec[0], rec[1], rec[2], v});
        }
        break;
      }
      default: fail(EVQGPU_ERR_UNSUPPORTED, "aggregate %s has no partial state format", fi.symbol.c_str());
    }
  }
  if (pos != n) fail(EVQGPU_ERR_FORMAT, "partial rows: %llu bytes behind the last state of a row", (unsigned long long) (n - pos));
}

void coordinator_parse_rows(evqgpu_query& q, const uint8_t* base, const uint64_t* row_starts, const uint64_t* row_ends, uint64_t nrows) {
  const size_t rec = 4 + q.state_ops.size();
  const size_t at = q.coord_records.size();
  q.coord_records.resize(at + nrows * rec);
  for (uint64_t i = 0; i < nrows; ++i) {
    if (row_ends[i] < row_starts[i]) fail(EVQGPU_ERR_ARG, "evqgpu_query_merge_rows: row %llu ends before it starts", (unsigned long long) i);
    parse_row(q, base + row_starts[i], row_ends[i] - row_starts[i], q.coord_nrecords + i, &q.coord_records[at + i * rec]);
  }
  q.coord_nrecords += nrows;
}

static void split_rows(const evqgpu_query& q, const uint8_t* body, uint64_t n, uint64_t base, uint64_t expect_rows, bool exact,
                       std::vector<uint64_t>& starts) {
  uint64_t pos = 0, rows = 0;
  while (pos < n && (!exact || rows < expect_rows)) {
    if (n - pos < 20) fail(EVQGPU_ERR_FORMAT, "partial rows: truncated group key");
    starts.push_back(base + pos);
    pos += 20;
    skip_states(q, body, n, pos);
    ++rows;
  }
  if (exact && (rows != expect_rows || pos != n))
    fail(EVQGPU_ERR_FORMAT, "partial rows: %llu rows announced, %llu found (%llu of %llu bytes used)", (unsigned long long) expect_rows,
         (unsigned long long) rows, (unsigned long long) pos, (unsigned long long) n);
}

static void copy_out(const std::vector<uint64_t>& starts, uint64_t end, uint64_t* row_offsets, uint64_t cap_rows, uint64_t* nrows_out) {
  *nrows_out = starts.size();
  if (!row_offsets || cap_rows < starts.size()) return;
  for (size_t i = 0; i < starts.size(); ++i) row_offsets[i] = starts[i];
  row_offsets[starts.size()] = end;
}
}  // namespace evq

extern "C" int evqgpu_partial_rows_split(const evqgpu_query_desc* desc, const void* body, uint64_t nbytes, uint64_t* row_offsets,
                                         uint64_t cap_rows, uint64_t* nrows_out) {
  return guarded([&] {
    if (!desc || (!body && nbytes) || !nrows_out) fail(EVQGPU_ERR_ARG, "evqgpu_partial_rows_split: null argument");
    evqgpu_query q;
    query_intake(&q, desc);
    std::vector<uint64_t> starts;
    split_rows(q, (const uint8_t*) body, nbytes, 0, 0, false, starts);
    copy_out(starts, nbytes, row_offsets, cap_rows, nrows_out);
  });
}

extern "C" int evqgpu_partial_cache_decode(const evqgpu_query_desc* desc, const void* entry, uint64_t nbytes, uint64_t* row_offsets,
                                           uint64_t cap_rows, uint64_t* nrows_out) {
  return guarded([&] {
    if (!desc || !entry || !nrows_out) fail(EVQGPU_ERR_ARG, "evqgpu_partial_cache_decode: null argument");
    const uint8_t* p = (const uint8_t*) entry;
    if (nbytes < 9 || p[0] != 0x01) fail(EVQGPU_ERR_FORMAT, "not a partial aggregation cache entry");
    uint64_t ngroups;
    memcpy(&ngroups, p + 1, 8);
    evqgpu_query q;
    query_intake(&q, desc);
    std::vector<uint64_t> starts;
    split_rows(q, p + 9, nbytes - 9, 9, ngroups, true, starts);
    copy_out(starts, nbytes, row_offsets, cap_rows, nrows_out);
  });
}

extern "C" int evqgpu_partial_frames_decode(const evqgpu_query_desc* desc, const void* frames, uint64_t nbytes, uint64_t* row_offsets,
                                            uint64_t cap_rows, uint64_t* nrows_out, uint64_t* nframes_out, int* end_of_request_out) {
  return guarded([&] {
    if (!desc || (!frames && nbytes) || !nrows_out) fail(EVQGPU_ERR_ARG, "evqgpu_partial_frames_decode: null argument");
    const uint8_t* p = (const uint8_t*) frames;
    evqgpu_query q;
    query_intake(&q, desc);
    // rows of consecutive frames are not adjacent in the buffer: every row gets its own [start, end) pair
    std::vector<uint64_t> starts, ends;
    uint64_t pos = 0, nframes = 0;
    int eor = 0;
    while (pos < nbytes) {
      if (nbytes - pos < 8) fail(EVQGPU_ERR_FORMAT, "partial result frames: truncated header");
      const uint32_t opcode = (uint32_t) p[pos] << 8 | p[pos + 1], flags = (uint32_t) p[pos + 2] << 8 | p[pos + 3];
      const uint64_t len = (uint64_t) p[pos + 4] << 24 | (uint64_t) p[pos + 5] << 16 | (uint64_t) p[pos + 6] << 8 | p[pos + 7];
      if (opcode != 0x0102) fail(EVQGPU_ERR_FORMAT, "partial result frames: opcode 0x%04x", opcode);
      if (eor) fail(EVQGPU_ERR_FORMAT, "partial result frames: data behind the end-of-request frame");
      pos += 8;
      if (len > nbytes - pos) fail(EVQGPU_ERR_FORMAT, "partial result frames: truncated payload");
      uint64_t q0 = pos;
      get_varuint(p, pos + len, q0);                      // frame flags
      const uint64_t rows = get_varuint(p, pos + len, q0);
      std::vector<uint64_t> s;
      split_rows(q, p + q0, pos + len - q0, q0, rows, true, s);
      for (size_t i = 0; i < s.size(); ++i) {
        starts.push_back(s[i]);
        ends.push_back(i + 1 < s.size() ? s[i + 1] : pos + len);
      }
      eor = (flags & 1) ? 1 : 0;
      pos += len;
      ++nframes;
    }
    *nrows_out = starts.size();
    if (nframes_out) *nframes_out = nframes;
    if (end_of_request_out) *end_of_request_out = eor;
    if (row_offsets && cap_rows >= starts.size())
      for (size_t i = 0; i < starts.size(); ++i) {
        row_offsets[2 * i] = starts[i];
        row_offsets[2 * i + 1] = ends[i];
      }
  });
}
