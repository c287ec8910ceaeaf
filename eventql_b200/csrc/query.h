// query.h - evqgpu_query: a fused FastCSTableScan (+ GroupByExpression) plan and its device state.
#pragma once
#include <memory>
#include <string>
#include <vector>
#include "context.h"
#include "expr.h"
#include "jit.h"
#include "kernels/evq_abi.h"
#include "table.h"

namespace evq {

// state-word op codes (kernels/evq_prelude.cuh EVQ_OP_*)
enum { OP_ADD_U64 = 0, OP_ADD_F64 = 1, OP_MIN_U64 = 2, OP_MAX_U64 = 3, OP_MIN_I64 = 4, OP_MAX_I64 = 5, OP_MIN_F64 = 6, OP_MAX_F64 = 7,
       OP_FIRST_ORD = 8, OP_FIRST_VAL = 9 };

struct SelectItem {
  ExprPtr expr;
  const Expr* agg = nullptr;   // the (single) aggregate call inside expr, if any
  // aggregate state words (indices into evqgpu_query::state_ops; decided per execution by layout_states()):
  int state0 = -1;             // count: rows word; sum: the sum; min/max: the extremum; mean: the sum (u64 low word or f64)
  int state_seen = -1;         // min/max/mean: number of non-NULL arguments (the rows word when the argument cannot be NULL)
  int state_carry = -1;        // mean over a uint64 argument: wraps of the 64-bit sum word (exact 128-bit integer sum)
  int distinct = -1;           // count_distinct: index into evqgpu_query::distinct_args
  bool is_string = false;      // a bare string column: the device column holds dictionary codes, fetched with evqgpu_query_fetch_strings
  // a non-aggregate item that is NOT a function of the GROUP BY key (e.g. the hidden columns the planner appends for
  // `ORDER BY <expression>`): the reference boxes its value for the group's first row (groupby.cc:161-172); here the pair
  // of state words state_first (row ordinal | tag, smallest wins) and state_first + 1 (value bits of that row)
  bool first = false;
  int state_first = -1;
};

// how one input column is laid out on the device (part of the kernel's specialisation key)
struct ColSig {
  bool used = false;
  uint32_t sql_type = 0;
  uint32_t kind = 0;       // EVQ_KIND_* of the DATA stream
  bool nullable = false;
  uint32_t dmax = 0;
  int data_stream = -1, level_stream = -1;
  int leb_slot = -1, null_slot = -1;
  uint32_t bits = 64;      // every value < 2^bits (max over the scanned tables)
  uint64_t vmax = ~0ull;   // every value <= vmax (max over the scanned tables; coarsened, see stat_ceil)
  uint64_t vmin = 0;       // every value >= vmin (min over the scanned tables; coarsened, see stat_floor)
  uint64_t vmin_present = 0;   // every non-NULL value >= this (vmin is 0 as soon as NULLs are present)
  uint32_t leb_len = 10;   // LEB128: longest value in bytes (max over the scanned tables)
  bool leb_uniform = false;   // LEB128: every value of every scanned table has exactly leb_len bytes
  int gen_slot = -1;       // fast kernel: index among the LEB128 columns that may need the boundary search (leb_len >= 2)
  bool packed = false;     // fast kernel: keep the column's raw bytes (4 rows per word) for the dp4a aggregates
  bool presence = false;   // fast kernel: keep the optional column's presence bytes (1 = not NULL, 4 rows per word): "seen" counters
  int sub_stream = -1;     // fast kernel: stream of the column's sub-index (entry points of every 8th value), -1 = none
  int nv_slot = -1;        // fast kernel, optional variable-length LEB128 column of <= 4 bytes: its values are decoded by VALUE ordinal
                           // (the required-column decode) into a shared staging array and gathered per row; index of that array
};

struct DenseMap {
  uint64_t key_min[EVQ_MAX_KEYS] = {0}, key_stride[EVQ_MAX_KEYS] = {0}, key_null_idx[EVQ_MAX_KEYS] = {0},
           key_range[EVQ_MAX_KEYS] = {0};
  uint64_t slots = 1;
};

struct KernelShape {
  std::vector<ColSig> cols;
  int tier = 1;        // 0 count pass, 3 projection pass (scan-only); 1 dense / single group; 2 global hash table
  int g1 = 1;          // dense slots (power of two >= groups) for tier 1
  int ncons = 256;     // consumer threads
  int nstages = 3;
  int kt = 1;          // fast kernel: consecutive row tiles per pipeline stage (one bulk copy per stream and stage)
  int min_ctas = 1;
  int nstreams = 0, nleb = 0, nnull = 0;
  bool dense_global = false;   // tier 2 without a hash table: the groups are a direct-addressed array (key bounds known), updated
                               // with atomics at L2 - for key ranges beyond the thread-private dense tier
  bool fast = false;   // all referenced columns are required: kernels/evq_scan_fast.cuh (4 consecutive rows per thread)
  int ngen = 0;        // fast kernel: LEB128 columns with leb_len >= 2 whose value boundaries are searched in the kernel
  int nnv = 0;         // fast kernel: optional columns decoded through a staging array (ColSig::nv_slot)
  int part_bits = 0;   // tier 2, fast kernel: > 0 = partitioned aggregation with 2^part_bits record partitions (query.cu)
  int slice_slots = 0; // ... > 0 = second partitioning level + table slices of (at most) this many slots in shared memory
  std::vector<int> rec_cols;   // ... the input columns a record carries (read by the GROUP BY expressions and aggregate arguments)
  int filter_stream = -1;    // stream of the tables' external row filter (FastCSTableScan::setFilter), -1 = none
  bool use_subidx = false;   // fast kernel: variable-length columns take their decode entry points from Column::sub_index
  DenseMap dense;      // tier 1 with g1 > 1: the key -> slot map is baked into the kernel text as constants
};

// parameter block of the emit kernel (mirrors the struct spelled in the generated text)
struct EmitParams {
  const u64* dense_state;
  EvqHashTable ht;
  u64 key_min[EVQ_MAX_KEYS];
  u64 key_stride[EVQ_MAX_KEYS];
  u64 key_null_idx[EVQ_MAX_KEYS];
  u64 key_range[EVQ_MAX_KEYS];
  u64 slots;              // dense slots or hash capacity
  u64* out_count;
  u64 out_capacity;
  u8* out_cols[EVQ_MAX_STREAMS];
  u8* out_sha;            // EVQGPU_QUERY_WIRE: 20-byte SHA-1 of every group's key tuple bytes (groupby.cc:129-135)
  u64* out_state;         // EVQGPU_QUERY_WIRE: the raw aggregate state words of every group [rows][nstate]
};

// parameter block of the tail kernel of a dense-tier execution (mirrors the generated struct EvqTailParams)
#define EVQ_P2P_XSLOT_WORDS 8192
#define EVQ_P2P_FLAGS_OFFSET (2ull * 16 * EVQ_P2P_XSLOT_WORDS * 8)   // bytes: the flag words sit behind the state slots
#define EVQ_P2P_BYTES (EVQ_P2P_FLAGS_OFFSET + 4096)
struct TailParams {
  EmitParams E;
  u64* dense_state;
  u64* ctl;            // [0..5] live status / counters (zeroed by the tail), [8..13] their values of this execution, [14] rows, [15] tails run
  u32 nranks, rank;
  u64 epoch;
  u64* xbuf_local;
  u64* flags_local;
  u64* xbuf_peer[16];
  u64* flags_peer[16];
  u64* merged_out;     // EVQGPU_QUERY_WIRE / debugging: the merged state words, or null
};

// parameter block of pass 2 of the partitioned aggregation (mirrors the generated struct EvqAggParams)
struct AggParams {
  EvqHashTable ht;
  const u64* part_buf;
  const u32* part_cursor;
  u64 part_cap;
  u32 nparts;
  u32 nseg;
  u32* status;
  u64* counters;
  u32* bar;       // CTAs done per partition, summed (zeroed before the launch)
  u32 window;     // partitions a CTA may run ahead of the slowest one
  u32 pad;
};

struct RepartParams {   // codegen.cc EvqRepartParams
  EvqHashTable ht;
  const u64* in;
  const u32* in_cursor;
  u64 in_cap;
  u64* out;
  u32* out_cursor;
  u64 out_cap;
  u32* status;
  u32 nparts, sub_bits, sub_shift, pad;
};

struct AggSmemParams {   // codegen.cc EvqAggSmemParams
  EvqHashTable ht;
  const u64* buf;
  const u32* cursor;
  u64 cap;
  u32* status;
  u32 nsub_total, slice_slots;
};

struct InitParams {
  u64* dense_state;
  EvqHashTable ht;
  u64 slots;
};

}  // namespace evq

struct evqgpu_query {
  evqgpu_ctx* ctx = nullptr;
  uint32_t flags = 0;
  std::vector<std::string> input_columns;
  evq::ExprPtr where;
  std::vector<evq::ExprPtr> group;
  std::vector<evq::SelectItem> select;
  std::vector<int> state_ops;       // op of every aggregate state word; word 0 = rows per group
  std::vector<std::string> state_keys;   // what the word accumulates ("sum:<expr>" ...): equal keys share one word
  std::vector<int> state_carry_of;  // for a carry word: index of the sum word whose wraps it counts, else -1
  std::vector<int> state_smem;      // index of the word among the thread-private accumulators (-1 for carry words)
  int nstate_smem = 0;
  // byte-wide aggregates of the fast dense kernel (rows counter, sums of 1-byte columns): accumulated with dp4a into
  // thread-private u32 registers instead of shared memory (codegen.cc: layout_narrow)
  std::vector<bool> state_global;   // words updated in global memory only (carry words, count_distinct counters)
  std::vector<const evq::Expr*> distinct_args;   // count_distinct: the distinct arguments (one (group, value) set each)
  std::vector<int> distinct_word;   // ... and the state word that counts the set's members per group
  std::vector<evq::DevBuf> dt_slots;
  uint64_t dt_cap = 0;
  std::vector<int> state_narrow;    // index into plane_sums, -1 otherwise
  std::vector<int> narrow_col;      // input columns whose raw bytes (4 rows per word) the kernel keeps next to the values
  int nnarrow = 0;                  // byte planes over all plane sums: EVQ_NNARROW * EVQ_NG u32 accumulators per thread
  int plane_groups = 0;             // EVQ_NG: groups the plane accumulators are kept for (the dense slots: <= 4 as a power of two,
                                    // 5..7 exactly - nibble 7 of the quad selector means "row did not pass")
  // Byte-plane sums (codegen.cc: layout_narrow): sum(W * B) per group with W < 2^32 split into byte planes and B <= 255
  // (or absent), accumulated 4 rows at a time with dp4a: sum = SUM_p 256^p * dp4a(plane_p(W), B & mask_of_group)
  struct PlaneOperand {
    const evq::Expr* expr = nullptr;   // null: the constant 1 (rows counter)
    int packed_col = -1;               // >= 0: bare 1-byte LEB128 column (its raw bytes are the one plane / the byte operand)
    int nplanes = 1;
    int swar = 0;                      // byte operand lit - col (1), lit + col (2) on the packed bytes of packed_col
    uint64_t swar_lit = 0;
    int presence_col = -1;             // >= 0: the byte operand is the presence (1 = not NULL) of this optional column: the
                                       // "seen" counters of min / max / mean over it
  };
  struct PlaneSum {
    int word = 0;                      // state word
    int w = -1, b = -1;                // indexes into plane_w / plane_b (-1: constant 1)
    int nplanes = 1;
    int plane0 = 0;                    // first plane: accumulators nacc[(plane0 + p) * G1 + g]
  };
  std::vector<PlaneOperand> plane_w, plane_b;
  std::vector<PlaneSum> plane_sums;
  bool swar_slots = false;             // dense slots of 4 rows computed on the packed key bytes
  std::string plane_sig;               // text form of the above for the kernel signature
  std::string merge_checked_layout;  // state layout all ranks were last verified to share
  uint64_t expected_groups = 0;
  std::vector<bool> col_used;
  std::vector<bool> col_is_string;   // plan input columns read as dictionary codes of a string column (strings.cu)
  std::vector<std::pair<evq::Expr*, std::string>> string_literals;   // string literals read as dictionary codes (imm), with their text
  std::vector<int> col_pred;         // plan input columns that are the verdict column of string_preds[i] (-1: not)
  std::vector<evq::StringPredicate> string_preds;
  bool string_keys = false;          // a GROUP BY expression is a string column
  std::vector<bool> group_is_string; // ... which ones
  bool has_first = false;            // some select item takes the value of its group's first row (SelectItem::first)
  bool coordinator = false;          // EVQGPU_QUERY_COORDINATOR: no scan; merges shards' partial rows (merge.cu coordinator_*)
  std::vector<uint64_t> coord_records;   // parsed rows: [3 key words][tag word][state words], kept until merge_finish
  uint64_t coord_nrecords = 0;
  std::vector<std::vector<uint64_t>> coord_pairs;   // per count_distinct argument: [3 key words][value] of every set member received
  u64* dense_base = nullptr;         // dense_state, offset by one word where that makes the first-row pairs 16-byte aligned

  // device state, reused across executions
  evq::DevBuf merge_recv, merge_send, merge_slots, merge_counts, merge_status;
  evq::DevBuf part_buf2, part_cursor2;   // ... second level (sub-partitions)
  evq::DevBuf part_buf, part_cursor;   // partitioned aggregation: the records of pass 1 and their per-partition counts
  bool no_partition = false;           // a partition overflowed once (skewed keys): this query keeps the direct hash tier
  uint64_t merge_cap = 0;           // capacity the merged table last needed (kept across executions)
  evq::DevBuf dense_state, ht_slots, status, counters, out_count, tile_counts, tile_base;
  // dense tier: status / counters / row count live in one control block that the tail kernel publishes and re-arms
  evq::DevBuf ctl;
  bool use_tail = false;            // the last execution ends in evq_tail (dense tier without count_distinct)
  bool armed = false;               // the state words and the control block are known to be identities / zero: the tail of
                                    // the previous execution re-armed them and nothing has touched them since
  std::string armed_sig;            // ... for the kernel of this signature (the state layout)
  bool tail_done = false;
  std::vector<evq::DevBuf> out_cols;
  evq::DevBuf out_sha, out_state;   // EVQGPU_QUERY_WIRE (PartialGroupByExpression rows)
  bool reordered = false;           // ORDER BY / LIMIT rewrote the result rows (the per-group side buffers no longer line up)
  uint64_t out_capacity = 0;
  uint64_t ht_cap = 0;

  // last execution
  evq::KernelShape shape;
  evq::DenseMap dense;
  bool dense_cache_valid = false, dense_cache_ok = false;
  evq::DenseMap dense_cache_map;            // (q.dense is the map of the LAST execution, which may have fallen back to the hash tier)
  bool prepared = false;                    // multi-rank: evqgpu_query_prepare agreed on slot assignment + layout ...
  std::vector<uint64_t> prepared_uids;      // ... for this set of tables
  std::vector<uint64_t> dense_cache_uids;   // bounds of the GROUP BY expressions are cached per set of (immutable) tables
  std::shared_ptr<evq::JitModule> module;
  std::string module_sig;   // what `module` was specialised for (group-by plans)
  std::vector<evqgpu_table*> tables;
  bool pending = false;
  bool merged = false;
  uint64_t num_rows_out = 0;
  evqgpu_query_stats stats = {};
  std::string kernel_source;
  float jit_ms_total = 0;
  evq::EmitParams emit;             // parameters of the (possibly deferred) emit kernel
  uint64_t emit_total_rows = 0;
  bool emitted = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;   // around scan launches, summed at finish
  ~evqgpu_query() {
    for (auto& e : prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  }
};

namespace evq {
// codegen.cc
// EVQGPU_QUERY_WIRE: bytes per group in out_sha - the 20-byte SHA-1 of the key tuple, or (string GROUP BY keys: the hash covers
// the string bytes, which only the host's dictionary has) the key tuple itself, 9 bytes per key, hashed in wire.cc
inline size_t wire_key_stride(const evqgpu_query& q) { return q.string_keys ? std::max<size_t>(20, 9 * q.group.size()) : 20; }
std::string generate_source(const evqgpu_query& q, const KernelShape& shape);
struct RecordField { int col, word, shift, bits; bool is_tag; };
struct RecordLayout { std::vector<RecordField> fields; size_t nwords = 1; };
RecordLayout record_layout(const KernelShape& shape);   // the packed record of the partitioned hash tier (codegen.cc)
int part_bin_records(int part_bits, size_t nrec);   // records per shared-memory bin of pass 1 of the partitioned hash tier
std::string generate_coordinator_source(const evqgpu_query& q);   // evq_emit over a table keyed by the 20-byte group keys
void layout_states(evqgpu_query& q, const KernelShape& shape);
void layout_narrow(evqgpu_query& q, const KernelShape& shape);   // after tier / g1 are known
int gen_chunks(const KernelShape& shape);
// query.cu
void prepare_query(evqgpu_query& q, std::vector<evqgpu_table*>& tables);
void emit_results(evqgpu_query& q);
void launch_tail(evqgpu_query& q, bool merge);
void finish_query(evqgpu_query& q);
// merge.cu
void merge_query(evqgpu_query& q);
void coordinator_finish(evqgpu_query& q);
// wire.cc
void coordinator_parse_rows(evqgpu_query& q, const uint8_t* base, const uint64_t* row_starts, const uint64_t* row_ends, uint64_t nrows);
// comm.cc
std::vector<uint64_t> comm_all_gather_host(evqgpu_ctx* ctx, const std::vector<uint64_t>& mine);   // [rank][mine.size()]
void comm_all_gather(evqgpu_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank);
void comm_all_reduce_sum_u64(evqgpu_ctx* ctx, void* buf, size_t nwords);   // in place
void comm_all_to_all(evqgpu_ctx* ctx, const void* send, const uint64_t* send_off, const uint64_t* send_bytes, void* recv,
                     const uint64_t* recv_off, const uint64_t* recv_bytes);
}  // namespace evq
