// cstable_format.h - host-side parsing / writing of the cstable container (header, metablocks, page index).
// Only metadata is touched on the host; page payloads are decoded on the device.
//
// Format references (reference tree, src/eventql/io/cstable/): cstable.h:35-110 (spec comment),
// cstable.cc:35-84 (readHeader), :89-132 (v0.1.0), :138-255 (v0.2.0 metablock, header, index),
// page_manager.cc:45-75 (page allocation), cstable_writer.cc:267-293 (commit order).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace evq {

struct PageRef {
  uint64_t offset;
  uint64_t size;
};

struct ColumnMeta {
  std::string name;
  uint32_t column_id = 0;
  uint32_t logical_type = 0;   // EVQ_COL_*
  uint32_t encoding = 0;       // EVQ_ENC_*
  uint32_t rlevel_max = 0;
  uint32_t dlevel_max = 0;
  // v0.2.0: page lists per stream kind in index order
  std::vector<PageRef> data_pages, rlevel_pages, dlevel_pages;
  // v0.1.0: one contiguous body
  uint64_t body_offset = 0, body_size = 0;
};

struct FileMeta {
  int version = 0;   // 1 = v0.1.0, 2 = v0.2.0
  uint64_t num_rows = 0;
  std::vector<ColumnMeta> columns;
};

// throws evq::Error(EVQGPU_ERR_FORMAT) on malformed input
FileMeta parse_cstable(const uint8_t* file, uint64_t nbytes);

// a byte range of a logical stream inside the file image
struct Extent {
  uint64_t file_offset;
  uint64_t nbytes;
};

struct StreamLayout {
  std::vector<Extent> extents;   // concatenate these to get the logical stream
  uint64_t total = 0;
  uint32_t bitpack_max = 0;      // max_value header of bit-packed streams
  bool present = false;
};

// Resolve the logical stream (pages in index order, minus the bit-packed header) of one column.
// kind: EVQ_STREAM_*.  For v0.1.0 the stream sizes come from the column body header.
StreamLayout stream_layout(const FileMeta& meta, const ColumnMeta& col, uint32_t kind, const uint8_t* file,
                           uint64_t nbytes);

void sha1(const uint8_t* data, size_t len, uint8_t out[20]);

uint32_t bits_needed(uint32_t v);   // libsimdcomp bits(): 32 - clz, 0 for 0

// ---- writer (v0.2.0) ----
struct WriteStream {
  uint32_t kind;                 // EVQ_STREAM_*
  uint32_t column_id;
  const uint8_t* payload;        // logical stream bytes (bit-packed: without the max_value header)
  uint64_t nbytes;
  bool bitpacked;
  uint32_t bitpack_max;
};

void write_cstable_v2(const std::string& path, uint64_t num_rows, const std::vector<ColumnMeta>& columns,
                      const std::vector<WriteStream>& streams);

}  // namespace evq
