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

#include <string>
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
    for (const auto& item : q->select)
      if (item.agg && item.agg->info().fn == Fn::COUNT_DISTINCT)
        fail(EVQGPU_ERR_UNSUPPORTED, "count_distinct keeps no value sets: its partial state cannot be serialised");
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
    EVQ_CUDA(cudaMemcpyAsync(keys, q->out_sha.as<u8>() + row0 * 20, n * 20, cudaMemcpyDeviceToHost, s));
    EVQ_CUDA(cudaMemcpyAsync(st.data(), q->out_state.as<u64>() + row0 * nstate, n * nstate * 8, cudaMemcpyDeviceToHost, s));
    for (size_t i = 0; i < q->select.size(); ++i) {
      if (q->select[i].agg) continue;
      const uint64_t w = q->select[i].expr->type == EVQ_BOOL ? 2 : 9;
      cols[i].resize(n * w);
      EVQ_CUDA(cudaMemcpyAsync(cols[i].data(), q->out_cols[i].as<u8>() + row0 * w, n * w, cudaMemcpyDeviceToHost, s));
    }
    EVQ_CUDA(cudaStreamSynchronize(s));
    std::vector<uint8_t> out;
    out.reserve(n * 32);
    for (uint64_t r = 0; r < n; ++r) {
      const uint64_t* g = &st[r * nstate];
      for (size_t i = 0; i < q->select.size(); ++i) {
        const SelectItem& item = q->select[i];
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
