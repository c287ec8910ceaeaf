#include "cstable_format.h"
#include <string.h>
#include <stdio.h>
#include <algorithm>
#include "util.h"

namespace evq {

static const uint8_t kMagic[4] = {0x23, 0x17, 0x23, 0x17};
static const uint64_t kMetaPos = 14, kMetaSize = 48, kSector = 512, kPageSize = 512 * 1024;

uint32_t bits_needed(uint32_t v) { return v == 0 ? 0 : 32 - __builtin_clz(v); }

// ---- SHA-1 (FIPS 180-1); used only for the 28-byte metablock checksum ----
void sha1(const uint8_t* data, size_t len, uint8_t out[20]) {
  uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
  std::vector<uint8_t> msg(data, data + len);
  msg.push_back(0x80);
  while (msg.size() % 64 != 56) msg.push_back(0);
  uint64_t bitlen = (uint64_t) len * 8;
  for (int i = 7; i >= 0; --i) msg.push_back((uint8_t) (bitlen >> (8 * i)));
  auto rol = [](uint32_t x, int n) { return (x << n) | (x >> (32 - n)); };
  for (size_t off = 0; off < msg.size(); off += 64) {
    uint32_t w[80];
    for (int i = 0; i < 16; ++i)
      w[i] = ((uint32_t) msg[off + 4 * i] << 24) | ((uint32_t) msg[off + 4 * i + 1] << 16) |
             ((uint32_t) msg[off + 4 * i + 2] << 8) | (uint32_t) msg[off + 4 * i + 3];
    for (int i = 16; i < 80; ++i) w[i] = rol(w[i - 3] ^ w[i - 8] ^ w[i - 14] ^ w[i - 16], 1);
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
    for (int i = 0; i < 80; ++i) {
      uint32_t f, k;
      if (i < 20) { f = (b & c) | (~b & d); k = 0x5A827999u; }
      else if (i < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1u; }
      else if (i < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8F1BBCDCu; }
      else { f = b ^ c ^ d; k = 0xCA62C1D6u; }
      uint32_t t = rol(a, 5) + f + e + k + w[i];
      e = d; d = c; c = rol(b, 30); b = a; a = t;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
  }
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 4; ++j) out[4 * i + j] = (uint8_t) (h[i] >> (24 - 8 * j));
}

namespace {

// [off, off + size) inside a file of nbytes - written so that file-supplied 64-bit values cannot wrap the sum
inline bool out_of_range(uint64_t off, uint64_t size, uint64_t nbytes) { return size > nbytes || off > nbytes - size; }

struct Cursor {
  const uint8_t* p;
  uint64_t n, pos;
  void need(uint64_t k) const {
    if (pos > n || k > n - pos) fail(EVQGPU_ERR_FORMAT, "cstable: truncated file (need %llu bytes at offset %llu)",
                          (unsigned long long) k, (unsigned long long) pos);
  }
  template <typename T>
  T rd() {
    need(sizeof(T));
    T v;
    memcpy(&v, p + pos, sizeof(T));
    pos += sizeof(T);
    return v;
  }
  uint64_t varuint() {
    uint64_t v = 0;
    for (int shift = 0;; shift += 7) {
      need(1);
      uint8_t b = p[pos++];
      if (shift < 64) v |= (uint64_t) (b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
      if (shift > 70) fail(EVQGPU_ERR_FORMAT, "cstable: malformed varint");
    }
  }
  std::string str(uint64_t len) {
    need(len);
    std::string s((const char*) p + pos, len);
    pos += len;
    return s;
  }
};

uint32_t logical_from_storage(uint32_t enc) {   // cstable.cc:100-117
  switch (enc) {
    case EVQ_ENC_BOOLEAN_BITPACKED: return EVQ_COL_BOOLEAN;
    case EVQ_ENC_FLOAT_IEEE754: return EVQ_COL_FLOAT;
    case EVQ_ENC_STRING_PLAIN: return EVQ_COL_STRING;
    default: return EVQ_COL_UNSIGNED_INT;
  }
}

}  // namespace

FileMeta parse_cstable(const uint8_t* file, uint64_t nbytes) {
  FileMeta meta;
  if (nbytes < 6 || memcmp(file, kMagic, 4) != 0) fail(EVQGPU_ERR_FORMAT, "not a valid cstable file");
  Cursor c{file, nbytes, 6};
  const uint8_t vnum = file[4];
  if (vnum == 1) {
    meta.version = 1;
    c.rd<uint64_t>();   // flags
    meta.num_rows = c.rd<uint64_t>();
    uint32_t ncols = c.rd<uint32_t>();
    for (uint32_t i = 0; i < ncols; ++i) {
      ColumnMeta col;
      col.encoding = c.rd<uint32_t>();
      col.logical_type = logical_from_storage(col.encoding);
      uint32_t nlen = c.rd<uint32_t>();
      col.name = c.str(nlen);
      col.rlevel_max = c.rd<uint32_t>();
      col.dlevel_max = c.rd<uint32_t>();
      col.body_offset = c.rd<uint64_t>();
      col.body_size = c.rd<uint64_t>();
      if (out_of_range(col.body_offset, col.body_size, nbytes)) fail(EVQGPU_ERR_FORMAT, "cstable: column body out of range");
      meta.columns.push_back(col);
    }
    std::sort(meta.columns.begin(), meta.columns.end(),
              [](const ColumnMeta& a, const ColumnMeta& b) { return a.name < b.name; });
    return meta;
  }
  if (vnum != 2) fail(EVQGPU_ERR_FORMAT, "unsupported cstable version: %u", (unsigned) vnum);
  meta.version = 2;
  if (nbytes < kMetaPos + 2 * kMetaSize + 128) fail(EVQGPU_ERR_FORMAT, "cstable: truncated header");
  // two metablocks; a block is valid iff its SHA-1 matches; the larger transaction id wins (cstable.cc:64-76)
  bool have = false;
  uint64_t best_txid = 0, index_offset = 0, index_size = 0;
  for (int k = 0; k < 2; ++k) {
    const uint8_t* mb = file + kMetaPos + k * kMetaSize;
    uint8_t h[20];
    sha1(mb, 28, h);
    if (memcmp(h, mb + 28, 20) != 0) continue;
    uint64_t txid, nrows, ioff;
    uint32_t isz;
    memcpy(&txid, mb, 8);
    memcpy(&nrows, mb + 8, 8);
    memcpy(&ioff, mb + 16, 8);
    memcpy(&isz, mb + 24, 4);
    // cstable.cc:64-68: block 0 wins only if its txid is strictly larger
    if (!have || txid >= best_txid) {
      have = true;
      best_txid = txid;
      meta.num_rows = nrows;
      index_offset = ioff;
      index_size = isz;
    }
  }
  if (!have) fail(EVQGPU_ERR_FORMAT, "can't open cstable: no valid metablocks found");
  c.pos = kMetaPos + 2 * kMetaSize + 128;
  uint64_t ncols = c.varuint();
  for (uint64_t i = 0; i < ncols; ++i) {
    ColumnMeta col;
    col.logical_type = (uint32_t) c.varuint();
    col.encoding = (uint32_t) c.varuint();
    col.column_id = (uint32_t) c.varuint();
    col.name = c.str(c.varuint());
    col.rlevel_max = (uint32_t) c.varuint();
    col.dlevel_max = (uint32_t) c.varuint();
    meta.columns.push_back(col);
  }
  if (out_of_range(index_offset, index_size, nbytes)) fail(EVQGPU_ERR_FORMAT, "cstable: page index out of range");
  Cursor ic{file, index_offset + index_size, index_offset};
  uint64_t n = ic.varuint();
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t type = ic.varuint(), cid = ic.varuint(), off = ic.varuint(), size = ic.varuint();
    if (out_of_range(off, size, nbytes)) fail(EVQGPU_ERR_FORMAT, "cstable: page out of range");
    for (auto& col : meta.columns) {
      if (col.column_id != cid) continue;
      PageRef pr{off, size};
      if (type == EVQ_STREAM_DATA) col.data_pages.push_back(pr);
      else if (type == EVQ_STREAM_RLEVEL) col.rlevel_pages.push_back(pr);
      else if (type == EVQ_STREAM_DLEVEL) col.dlevel_pages.push_back(pr);
    }
  }
  return meta;
}

StreamLayout stream_layout(const FileMeta& meta, const ColumnMeta& col, uint32_t kind, const uint8_t* file,
                           uint64_t nbytes) {
  StreamLayout L;
  const bool data_bitpacked = col.encoding == EVQ_ENC_UINT32_BITPACKED || col.encoding == EVQ_ENC_BOOLEAN_BITPACKED;
  if (meta.version == 2) {
    const std::vector<PageRef>* pages = kind == EVQ_STREAM_DATA ? &col.data_pages
                                       : kind == EVQ_STREAM_DLEVEL ? &col.dlevel_pages : &col.rlevel_pages;
    const bool bp = kind != EVQ_STREAM_DATA || data_bitpacked;
    for (size_t k = 0; k < pages->size(); ++k) {
      uint64_t off = (*pages)[k].offset, sz = (*pages)[k].size;
      if (bp && k == 0) {   // page_reader_bitpacked.cc:41-44: first page starts with u32 max_value
        if (sz < 4) fail(EVQGPU_ERR_FORMAT, "cstable: bit-packed page too small");
        memcpy(&L.bitpack_max, file + off, 4);
        off += 4;
        sz -= 4;
      }
      if (!L.extents.empty() && L.extents.back().file_offset + L.extents.back().nbytes == off) {
        L.extents.back().nbytes += sz;   // adjacent pages: one copy
      } else {
        L.extents.push_back({off, sz});
      }
      L.total += sz;
    }
    L.present = !pages->empty();
    return L;
  }
  // v0.1.0 body: u64 num_vals | u64 rlvl_size | u64 dlvl_size | u64 data_size | rlvl | dlvl | data
  // (columns/v1/ColumnReader.h:37-55); level streams carry no header, width = bits(level_max)
  if (col.body_size < 32) fail(EVQGPU_ERR_FORMAT, "cstable v1: column body too small");
  uint64_t hdr[4];
  memcpy(hdr, file + col.body_offset, 32);
  const uint64_t rsz = hdr[1], dsz = hdr[2], datasz = hdr[3];
  {   // (each size against what is left of the body: the sum of three file-supplied u64 may wrap)
    uint64_t left = col.body_size - 32;
    for (uint64_t part : {rsz, dsz, datasz}) {
      if (part > left) fail(EVQGPU_ERR_FORMAT, "cstable v1: stream sizes exceed column body");
      left -= part;
    }
    if (out_of_range(col.body_offset, col.body_size, nbytes)) fail(EVQGPU_ERR_FORMAT, "cstable v1: column body out of range");
  }
  uint64_t off, sz;
  if (kind == EVQ_STREAM_RLEVEL) { off = 32; sz = rsz; L.bitpack_max = col.rlevel_max; }
  else if (kind == EVQ_STREAM_DLEVEL) { off = 32 + rsz; sz = dsz; L.bitpack_max = col.dlevel_max; }
  else {
    off = 32 + rsz + dsz; sz = datasz;
    if (col.encoding == EVQ_ENC_UINT32_BITPACKED) {   // v1/BitPackedIntColumnReader.cc:31-43
      if (sz >= 4) { memcpy(&L.bitpack_max, file + col.body_offset + off, 4); off += 4; sz -= 4; }
    } else if (col.encoding == EVQ_ENC_BOOLEAN_BITPACKED) {
      L.bitpack_max = 1;                              // v1/BooleanColumnReader.cc:31-38
    }
  }
  if (sz) L.extents.push_back({col.body_offset + off, sz});
  L.total = sz;
  L.present = sz > 0;
  return L;
}

// ---- writer ----
namespace {
void put_varuint(std::vector<uint8_t>& b, uint64_t v) {
  do {
    uint8_t x = v & 0x7f;
    v >>= 7;
    if (v) x |= 0x80;
    b.push_back(x);
  } while (v);
}
template <typename T>
void put(std::vector<uint8_t>& b, T v) {
  const uint8_t* p = (const uint8_t*) &v;
  b.insert(b.end(), p, p + sizeof(T));
}
}  // namespace

void write_cstable_v2(const std::string& path, uint64_t num_rows, const std::vector<ColumnMeta>& columns,
                      const std::vector<WriteStream>& streams) {
  std::vector<uint8_t> hdr;
  hdr.insert(hdr.end(), kMagic, kMagic + 4);
  put<uint16_t>(hdr, 2);
  put<uint64_t>(hdr, 0);
  hdr.resize(hdr.size() + 2 * kMetaSize + 128, 0);
  put_varuint(hdr, columns.size());
  for (const auto& c : columns) {
    put_varuint(hdr, c.logical_type);
    put_varuint(hdr, c.encoding);
    put_varuint(hdr, c.column_id);
    put_varuint(hdr, c.name.size());
    hdr.insert(hdr.end(), c.name.begin(), c.name.end());
    put_varuint(hdr, c.rlevel_max);
    put_varuint(hdr, c.dlevel_max);
  }
  hdr.resize(round_up(hdr.size(), kSector), 0);

  FILE* f = fopen(path.c_str(), "wb");
  if (!f) fail(EVQGPU_ERR_ARG, "cannot create %s", path.c_str());
  uint64_t off = hdr.size();
  fwrite(hdr.data(), 1, hdr.size(), f);
  std::vector<uint8_t> index_entries;
  uint64_t nentries = 0;
  std::vector<uint8_t> zeros(kPageSize, 0);
  for (const auto& s : streams) {
    uint64_t done = 0;
    bool first = true;
    const uint32_t b = bits_needed(s.bitpack_max);
    if (s.bitpacked && b == 0) continue;   // page_writer_bitpacked.cc:41-43: no page is ever allocated
    const uint64_t psz = s.bitpacked ? (uint64_t) 16 * b * 1024 : kPageSize;
    while (done < s.nbytes || (first && s.bitpacked && s.nbytes == 0 && false)) {
      const uint64_t chunk = std::min(psz, s.nbytes - done);
      uint64_t page_bytes = psz;
      if (s.bitpacked && first) {
        fwrite(&s.bitpack_max, 4, 1, f);
        page_bytes += 4;
      }
      fwrite(s.payload + done, 1, chunk, f);
      for (uint64_t pad = psz - chunk; pad > 0;) {
        const uint64_t k = std::min<uint64_t>(pad, zeros.size());
        fwrite(zeros.data(), 1, k, f);
        pad -= k;
      }
      put_varuint(index_entries, s.kind);
      put_varuint(index_entries, s.column_id);
      put_varuint(index_entries, off);
      put_varuint(index_entries, page_bytes);
      ++nentries;
      off += page_bytes;
      done += chunk;
      first = false;
    }
  }
  std::vector<uint8_t> index;
  put_varuint(index, nentries);
  index.insert(index.end(), index_entries.begin(), index_entries.end());
  fwrite(index.data(), 1, index.size(), f);
  // metablock: txid 1 -> slot 1 (cstable_file.cc:172-174)
  std::vector<uint8_t> mb;
  put<uint64_t>(mb, 1);
  put<uint64_t>(mb, num_rows);
  put<uint64_t>(mb, off);
  put<uint32_t>(mb, (uint32_t) index.size());
  uint8_t h[20];
  sha1(mb.data(), mb.size(), h);
  mb.insert(mb.end(), h, h + 20);
  fseek(f, (long) (kMetaPos + kMetaSize), SEEK_SET);
  fwrite(mb.data(), 1, mb.size(), f);
  if (fclose(f) != 0) fail(EVQGPU_ERR_ARG, "write to %s failed", path.c_str());
}

}  // namespace evq
