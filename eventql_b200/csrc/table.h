// table.h - evqgpu_table: one cstable file / partition segment resident in HBM.
//
// HBM layout per loaded column:
//   data    : the column's DATA pages concatenated in index order (page padding dropped, bit-packed header
//             stripped), 256-byte aligned, zero-padded by >= 128 bytes so tile copies may over-read
//   dlevel  : same for the DLEVEL pages of optional columns (bit-packed, libsimdcomp vertical layout)
//   val_index[t] : number of non-NULL values before row tile t        (optional columns)   u64[ntiles+1]
//   off_index[t] : byte offset of the first value of row tile t       (LEB128 columns)     u64[ntiles+1]
//   sub_index[t][g] : byte offset inside tile t of its value 4g        (LEB128 columns whose values differ in length)
//                     u16[ntiles][256] = 0.5 B per row; lets the fast scan kernel start two independent 4-value decode
//                     chains per thread without searching value boundaries
// Flat STRING_PLAIN columns (strings.cu) keep the data stream plus a value index built when the column is loaded:
//   str_start[v] : byte offset of the payload of value v (behind its length prefix)   u64[nvalues]
//   str_len[v]   : its length                                                         u32[nvalues]
//   row_value[r] : value ordinal of record r, 0xffffffff = NULL                       u32[nrows]  (optional columns)
// A row tile is EVQ_TILE_ROWS = 1024 records.
#pragma once
#include <memory>
#include <string>
#include <vector>
#include "context.h"
#include "cstable_format.h"
#include "kernels/evq_abi.h"

namespace evq {

struct DeviceStream {
  DevBuf buf;
  uint64_t nbytes = 0;         // payload bytes (logical stream length, may include the zero tail of the last page)
  uint32_t bitpack_max = 0;
  bool present = false;
};

struct Column {
  ColumnMeta meta;
  uint32_t sql_type = 0;
  bool loaded = false;
  bool scannable = false;      // flat, numeric
  DeviceStream data, dlevel;
  DevBuf off_index, val_index;
  DevBuf sub_index;            // variable-length LEB128 columns: u16[ntiles][256] byte offset (from the tile's first byte) of every 4th value
  uint64_t num_values = 0;
  uint64_t data_payload_bytes = 0;    // algorithmic bytes of the DATA stream
  uint64_t level_payload_bytes = 0;
  uint32_t data_kind = 0;      // EVQ_KIND_*
  uint32_t data_bits = 0;
  uint32_t level_bits = 0;
  uint32_t value_bits = 64;    // statistic: every value of the column is < 2^value_bits (computed when the column is loaded)
  uint32_t leb_max_len = 10;   // LEB128: the longest value in bytes
  bool leb_uniform = false;    // LEB128: every value is leb_max_len bytes long (value i starts at byte i * leb_max_len: no sub-index)
  uint64_t value_max = ~0ull;  // statistic: largest value (exact for plain columns, 1-byte LEB128 and sub-indexed LEB128 columns; a bound otherwise)
  uint64_t value_min = 0;      // statistic: smallest value (exact for the same columns, 0 otherwise; 0 when NULLs are present)
  uint64_t value_min_present = 0;   // ... over the non-NULL values only: bounds the encoded length of every value in the stream
  uint32_t data_tile_cap = 0;  // max bytes one tile copy of the data stream can need (multiple of 16)
  uint32_t level_tile_cap = 0;
  // flat string columns
  bool is_string = false;
  DevBuf str_start, str_len, row_value;
  // ... and, once a query compares or groups by the column, its values as dictionary codes (evqgpu_ctx::string_codes): a
  // UINT32_PLAIN shadow column with the same definition levels, which is what the scan kernels read
  std::unique_ptr<Column> code_col;
  // verdict columns of string predicates over this column (key: function, operand order, literal): value i = predicate of
  // the string value i, looked up through the dictionary code (strings.cu ensure_pred_column)
  std::map<std::string, std::unique_ptr<Column>> pred_cols;
};

#define EVQ_KIND_STRING_HOST 255u   // Column::data_kind of string columns (host side only: the scan kernels never see them)

}  // namespace evq

struct evqgpu_table {
  evqgpu_ctx* ctx = nullptr;
  evq::FileMeta meta;
  const uint8_t* file = nullptr;
  uint64_t file_bytes = 0;
  uint64_t num_rows = 0;
  uint32_t num_tiles = 0;
  std::vector<evq::Column> cols;
  bool from_file = false;
  evq::DevBuf filter;        // external row filter (FastCSTableScan::setFilter): 1 bit per row, LSB first, zero padded
  bool has_filter = false;
  uint64_t uid = 0;   // unique per table object (pointers can be reused after destroy)

  int find(const char* name) const {
    for (size_t i = 0; i < cols.size(); ++i)
      if (cols[i].meta.name == name) return (int) i;
    return -1;
  }
};

namespace evq {
uint32_t sql_type_of(const ColumnMeta& m);
void table_init_columns(evqgpu_table* t);
void table_load_column(evqgpu_table* t, Column& c);
void table_finish_column(evqgpu_table* t, Column& c);   // indexes + tile caps after the streams are on the device
// strings.cu: value index of a string column whose streams are on the device; `host_stream` is the logical DATA stream
void table_finish_string_column(evqgpu_table* t, Column& c, const uint8_t* host_stream, uint64_t nbytes);
uint32_t string_code(evqgpu_ctx* ctx, const std::string& s);        // code of a string in the context's dictionary (inserted if new)
struct StringPredicate {
  int fn = 0;              // evq::Fn as int: LT / LTE / GT / GTE / STARTSWITH / ENDSWITH
  bool column_first = true;   // fn(column, literal) or fn(literal, column)
  std::string literal;
  bool invert = false;     // the verdict column holds NOT predicate (chosen so that a NULL row, which reads 0, gets predicate(""))
};
bool string_predicate_eval(const StringPredicate& p, const std::string& value);   // the reference's semantics, on the host
const Column* ensure_pred_column(evqgpu_table* t, Column& c, const StringPredicate& p);
const Column* ensure_code_column(evqgpu_table* t, Column& c);
// COLLECTIVE (multi-rank jobs): make the string dictionaries of all ranks identical - the ranks exchange the strings they
// added since the last call, append their union in rank order, and renumber their provisional codes (code columns included).
// Returns true when codes of this rank changed.
bool sync_dictionary(evqgpu_ctx* ctx);
uint64_t next_table_uid();       // the dictionary-coded shadow of a loaded string column
}  // namespace evq
