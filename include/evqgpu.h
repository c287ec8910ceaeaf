/*
 * evqgpu.h - C ABI of the B200-native columnar scan / filter / GROUP BY engine.
 *
 * This is the drop-in boundary.  Everything above it is host C++ that
 * keeps EventQL's operator surface (csql::TableExpression / FastCSTableScan /
 * GroupByExpression / TableProvider - see eventql_b200/host/); everything below it is
 * hand-written CUDA for sm_100a.  Plain pointers and sizes only: no C++ types, no
 * torch types, no exceptions cross this line.
 *
 * Conventions
 *   - every call returns an evqgpu_status (0 = ok); on failure the message is available
 *     from evqgpu_last_error() (per calling thread)
 *   - handles are opaque; the caller owns host buffers, the library owns device memory
 *   - one host thread per context at a time (single-threaded pull, like
 *     csql::TableExpression: sql/table_expression.h:35-50 in the reference)
 *   - there is NO CPU fallback: a query the device path cannot run fails with
 *     EVQGPU_ERR_UNSUPPORTED
 *
 * Reference interfaces replaced (paths relative to the reference's src/eventql/):
 *   evqgpu_table_open / _column_info     io/cstable/cstable_reader.cc:133-200 (CSTableReader::openFile),
 *                                        io/cstable/cstable.cc:35-84,200-255  (readHeader/readIndex)
 *   evqgpu_table_load_columns            io/cstable/cstable_reader.cc:78-131  (openColumnV2: bind level +
 *                                        data page readers), io/cstable/page_manager.cc:125-171
 *   evqgpu_table_create/_add_*           io/cstable/cstable_file.cc (in-memory cstable arena)
 *   evqgpu_query_create                  sql/CSTableScan.cc:726-755 (FastCSTableScan::execute: compile
 *                                        select list + WHERE) and sql/runtime/compiler.cc:50-104
 *                                        (split of a select item into accumulate / get programs)
 *   evqgpu_query_execute                 sql/CSTableScan.cc:757-858 (nextBatch: decode, WHERE, project) fused
 *                                        with sql/statements/select/groupby.cc:69-185 (GroupByExpression::execute)
 *   evqgpu_query_fetch                   sql/statements/select/groupby.cc:187-220 (GroupByExpression::nextBatch)
 *                                        output in the packed SVector encoding of sql/svalue.cc:533-549
 *   evqgpu_query_merge                   sql/statements/select/groupby.cc:528-637 (GroupByMergeExpression)
 *   evqgpu_function_lookup               sql/runtime/symboltable.cc:33-39,162-175 (symbol strings)
 *   evqgpu_table_decode_string_column    io/cstable/columns/column_reader_string.cc + page_reader_lenencstring.cc:37-62
 *                                        (StringColumnReader::readString), sql/CSTableScan.cc:970-995 (fetchColumnString)
 *   evqgpu_lsm_build_filters             server/sql/partition_cursor.cc:157-194,216-218 (the visibility filter loop of
 *                                        PartitionCursor::openNextTable and its setFilter call)
 */
#ifndef EVQGPU_H
#define EVQGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVQGPU_ABI_VERSION 5

#if defined(__GNUC__)
#define EVQGPU_API __attribute__((visibility("default")))
#else
#define EVQGPU_API
#endif

typedef struct evqgpu_ctx evqgpu_ctx;
typedef struct evqgpu_table evqgpu_table;
typedef struct evqgpu_query evqgpu_query;

typedef enum evqgpu_status {
  EVQGPU_OK = 0,
  EVQGPU_ERR_ARG = 1,          /* malformed argument / plan */
  EVQGPU_ERR_UNSUPPORTED = 2,  /* valid, but not runnable on the device path (no CPU fallback) */
  EVQGPU_ERR_CUDA = 3,         /* CUDA / NVRTC / NCCL failure, incl. "no device" */
  EVQGPU_ERR_RUNTIME = 4,      /* query-time error the reference raises too (division by zero ...) */
  EVQGPU_ERR_FORMAT = 5,       /* not a valid cstable file */
  EVQGPU_ERR_NOMEM = 6
} evqgpu_status;

/* value types == csql::SType (sql/svalue.h:41-49) */
enum {
  EVQ_NIL = 0,
  EVQ_UINT64 = 1,
  EVQ_INT64 = 2,
  EVQ_FLOAT64 = 3,
  EVQ_BOOL = 4,
  EVQ_STRING = 5,
  EVQ_TIMESTAMP64 = 6
};

/* value tag == csql::STag (sql/svalue.h:51-56) */
#define EVQ_STAG_NULL 1

/* cstable logical column types == cstable::ColumnType (io/cstable/cstable.h:112-120) */
enum {
  EVQ_COL_SUBRECORD = 0,
  EVQ_COL_BOOLEAN = 1,
  EVQ_COL_UNSIGNED_INT = 2,
  EVQ_COL_SIGNED_INT = 3,
  EVQ_COL_STRING = 4,
  EVQ_COL_FLOAT = 5,
  EVQ_COL_DATETIME = 6
};

/* cstable storage encodings == cstable::ColumnEncoding (io/cstable/cstable.h:122-130) */
enum {
  EVQ_ENC_BOOLEAN_BITPACKED = 1,
  EVQ_ENC_UINT32_BITPACKED = 10,
  EVQ_ENC_UINT32_PLAIN = 11,
  EVQ_ENC_UINT64_PLAIN = 12,
  EVQ_ENC_UINT64_LEB128 = 13,
  EVQ_ENC_FLOAT_IEEE754 = 14,
  EVQ_ENC_STRING_PLAIN = 100
};

/* page stream kinds == cstable::PageIndexEntryType (io/cstable/cstable.h:186-190) */
enum { EVQ_STREAM_DATA = 1, EVQ_STREAM_RLEVEL = 2, EVQ_STREAM_DLEVEL = 3 };

/* ------------------------------------------------------------------------------------------
 * context
 * ---------------------------------------------------------------------------------------- */

/* Bind a context to one CUDA device (one process per GPU; the device must exist - there is
 * no host execution mode).  flags: reserved, pass 0. */
EVQGPU_API int evqgpu_ctx_create(int device, uint64_t flags, evqgpu_ctx** out);
EVQGPU_API void evqgpu_ctx_destroy(evqgpu_ctx* ctx);

/* Message of the last failed call on this thread ("" if none). Never NULL. */
EVQGPU_API const char* evqgpu_last_error(void);

/* ABI version of the loaded library (== EVQGPU_ABI_VERSION it was built with). */
EVQGPU_API int evqgpu_abi_version(void);

/* Pinned host memory helpers (so that evqgpu_table_load_columns can DMA straight from the
 * caller's file image).  evqgpu_host_register pins an existing allocation in place. */
EVQGPU_API int evqgpu_host_alloc(evqgpu_ctx* ctx, uint64_t nbytes, void** out);
EVQGPU_API int evqgpu_host_free(evqgpu_ctx* ctx, void* ptr);
EVQGPU_API int evqgpu_host_register(evqgpu_ctx* ctx, void* ptr, uint64_t nbytes);
EVQGPU_API int evqgpu_host_unregister(evqgpu_ctx* ctx, void* ptr);

/* The CUDA stream (cudaStream_t) all work of this context is issued on. */
EVQGPU_API void* evqgpu_ctx_stream(evqgpu_ctx* ctx);
EVQGPU_API int evqgpu_ctx_synchronize(evqgpu_ctx* ctx);

/* Profiling: when on, every scan kernel launch is bracketed by CUDA events on the context stream and
 * evqgpu_query_stats.scan_ms reports their summed device time (bench.py's roofline figure). */
EVQGPU_API int evqgpu_ctx_set_profiling(evqgpu_ctx* ctx, int on);
/* Kernels launched through this context since it was created (bench.py's gpu_launches claim). */
EVQGPU_API uint64_t evqgpu_ctx_kernel_launches(const evqgpu_ctx* ctx);

/* ------------------------------------------------------------------------------------------
 * tables (one cstable file / partition segment, resident in HBM)
 * ---------------------------------------------------------------------------------------- */

typedef struct evqgpu_column_info {
  const char* name;         /* owned by the table */
  uint32_t column_id;
  uint32_t logical_type;    /* EVQ_COL_* */
  uint32_t encoding;        /* EVQ_ENC_* */
  uint32_t rlevel_max;
  uint32_t dlevel_max;
  uint32_t sql_type;        /* EVQ_* SType as mapped by sql/CSTableScanProvider.cc:79-107 */
  uint32_t loaded;          /* 1 once resident in HBM */
  uint64_t data_bytes;      /* payload bytes of the DATA stream (algorithmic bytes, padding excluded) */
  uint64_t level_bytes;     /* payload bytes of the DLEVEL stream, 0 if the column is required */
  uint64_t num_values;      /* non-NULL values (known once loaded) */
  uint32_t value_bits;      /* statistic: every value < 2^value_bits (known once loaded) */
  uint32_t leb_max_len;     /* statistic: longest LEB128 value in bytes (known once loaded) */
  uint64_t value_min;       /* statistic: value range of the non-NULL values (exact where the loader can decode the
                               column in one pass, else [0, 2^value_bits - 1]); value_min is 0 when NULLs are present */
  uint64_t value_max;
} evqgpu_column_info;

/* Parse header + page index of a cstable file image (v0.1.0 and v0.2.0).  Host only: nothing
 * is copied yet.  `file` must stay valid until the columns of interest are loaded. */
EVQGPU_API int evqgpu_table_open(evqgpu_ctx* ctx, const void* file, uint64_t nbytes, evqgpu_table** out);

/* Build a table from logical streams instead of a file image (in-memory arena, generators).
 * Streams are the concatenation of a column's pages in index order, without the bit-packed
 * page's 4-byte max_value header (pass it as `bitpack_max`; ignored for other encodings).
 * ptr may be a host pointer or - with EVQGPU_STREAM_DEVICE - a device pointer (copied). */
EVQGPU_API int evqgpu_table_create(evqgpu_ctx* ctx, uint64_t num_rows, evqgpu_table** out);
EVQGPU_API int evqgpu_table_add_column(evqgpu_table* tbl, const char* name, uint32_t logical_type,
                                       uint32_t encoding, uint32_t rlevel_max, uint32_t dlevel_max);
#define EVQGPU_STREAM_DEVICE 1u
EVQGPU_API int evqgpu_table_add_stream(evqgpu_table* tbl, const char* column, uint32_t kind, const void* ptr,
                                       uint64_t nbytes, uint32_t bitpack_max, uint32_t flags);

/* External row filter of a table: FastCSTableScan::setFilter (sql/CSTableScan.h:36-41, CSTableScan.cc:1006-1009), the
 * LSM visibility bitmap of a partition segment.  `bits` holds one bit per row, LSB first (row i = bit i%8 of byte i/8,
 * 1 = keep); the scan evaluates WHERE on every row and then drops the rows whose bit is 0 (CSTableScan.cc:826-833).
 * ptr may be a host pointer or - with EVQGPU_STREAM_DEVICE - a device pointer; it is copied.  bits == NULL removes the
 * filter.  nrows must equal the table's row count. */
EVQGPU_API int evqgpu_table_set_filter(evqgpu_table* tbl, const void* bits, uint64_t nrows, uint32_t flags);

EVQGPU_API void evqgpu_table_destroy(evqgpu_table* tbl);

EVQGPU_API uint64_t evqgpu_table_num_rows(const evqgpu_table* tbl);
EVQGPU_API uint32_t evqgpu_table_num_columns(const evqgpu_table* tbl);
EVQGPU_API int evqgpu_table_column_info(const evqgpu_table* tbl, uint32_t idx, evqgpu_column_info* out);
EVQGPU_API int evqgpu_table_find_column(const evqgpu_table* tbl, const char* name); /* index or -1 */

/* Make the named columns resident: host->device copy of their pages (async DMA when the file
 * image is pinned), then the device-side row-tile index and min/max statistics.  Idempotent.
 * names == NULL loads every flat, non-string column; a flat STRING_PLAIN column is loaded when it is named (stream + the
 * value index {start, length} + the record -> value map of optional columns). */
EVQGPU_API int evqgpu_table_load_columns(evqgpu_table* tbl, const char* const* names, uint32_t n);

/* Device -> host copy of one logical stream (tests, cstable export). nbytes_out may exceed cap:
 * nothing is copied then. */
EVQGPU_API int evqgpu_table_read_stream(evqgpu_table* tbl, const char* column, uint32_t kind, void* dst,
                                        uint64_t cap, uint64_t* nbytes_out, uint32_t* bitpack_max_out);

/* Decode a loaded flat column to one 9-byte packed SVector element per row ([8 B value][1 B tag],
 * 2 B for BOOL) on the device and copy rows [row0, row0+nrows) to `dst` (host).  This is
 * FastCSTableScan::fetchColumn* (sql/CSTableScan.cc:860-968) on its own, used by the decode parity tests. */
EVQGPU_API int evqgpu_table_decode_column(evqgpu_table* tbl, const char* column, uint64_t row0, uint64_t nrows,
                                          void* dst, uint64_t cap);

/* Decode rows [row0, row0 + nrows) of a flat STRING_PLAIN column (v0.2.0 `varuint length + bytes` values that may straddle
 * pages, v0.1.0 `u32 length + bytes`) on the device into the packed STRING SVector FastCSTableScan::fetchColumnString
 * builds (sql/CSTableScan.cc:970-995, sql/svalue.cc:533-549): present -> [u32 length][bytes][tag 0], NULL -> [u32 0][tag
 * EVQ_STAG_NULL].  *nbytes_out = bytes needed; nothing is copied when dst == NULL or cap is smaller. */
EVQGPU_API int evqgpu_table_decode_string_column(evqgpu_table* tbl, const char* column, uint64_t row0, uint64_t nrows,
                                                 void* dst, uint64_t cap, uint64_t* nbytes_out);

/* Read back the table's external row filter (1 bit per row, LSB first; (num_rows + 7) / 8 bytes).  *has_filter_out = 0
 * when the table has none (nothing is copied). */
EVQGPU_API int evqgpu_table_get_filter(evqgpu_table* tbl, void* bits, uint64_t cap_bytes, int* has_filter_out);

/* ------------------------------------------------------------------------------------------
 * LSM visibility: the row filters of a partition's segments, built on the device
 * ---------------------------------------------------------------------------------------- */

/* One table PartitionCursor::openNextTable visits (server/sql/partition_cursor.cc:82-155).  Segments are passed in the
 * cursor's order: head arena, compacting arena, then the partition's LSM tables newest first. */
#define EVQGPU_LSM_SKIP_COLUMN 1u   /* tbl->has_skiplist(): rows whose __lsm_skip column is true are dropped */
#define EVQGPU_LSM_NO_FILTER 2u     /* the cursor's needs_filter == false (partition_cursor.cc:149-155): every row stays
                                       visible, the segment's ids are not recorded, any filter of the table is removed */
#define EVQGPU_LSM_HAS_UPDATES 4u   /* LSMTableRef::has_updates of an on-disk table */
#define EVQGPU_LSM_OLDEST 8u        /* the partition's oldest on-disk table (tblidx == 0, visited last) */
#define EVQGPU_LSM_AUTO 16u         /* decide needs_filter by the cursor's own rule (partition_cursor.cc:149-155): a table
                                       without a skiplist is scanned unfiltered when no id has been recorded before it and it
                                       is the oldest table or has no updates */
typedef struct evqgpu_lsm_segment {
  evqgpu_table* table;      /* needs the reference's bookkeeping columns (db/partition_arena.cc:41-45): __lsm_id (string, 20
                               bytes per value), __lsm_is_update and, with EVQGPU_LSM_SKIP_COLUMN, __lsm_skip (booleans) */
  const void* skiplist;     /* arena skiplist (PartitionArena::SkiplistReader): host pointer, 1 bit per row, LSB first,
                               1 = skip; overrides the column; NULL = none */
  uint32_t flags;           /* EVQGPU_LSM_* */
  uint32_t filtered;        /* out: 1 = a filter was installed, 0 = the segment is scanned unfiltered */
  uint64_t visible_rows;    /* out: rows the segment's filter keeps (all rows when unfiltered) */
} evqgpu_lsm_segment;

/* Build and install (as with evqgpu_table_set_filter) the row filter of every segment: a row is dropped if it is skipped
 * or if an earlier row - of an earlier segment, or earlier in its own - was a visible update with the same __lsm_id
 * (partition_cursor.cc:157-194).  An id that is not 20 bytes long fails with EVQGPU_ERR_RUNTIME ("invalid SHA1Hash",
 * util/SHA1.cc:79-85).  All segments must belong to `ctx`. */
EVQGPU_API int evqgpu_lsm_build_filters(evqgpu_ctx* ctx, evqgpu_lsm_segment* segs, uint32_t nsegs);

/* Write the table as a v0.2.0 cstable file (the layout of io/cstable/cstable_writer.cc:267-293). */
EVQGPU_API int evqgpu_table_write_file(evqgpu_table* tbl, const char* path);

/* ------------------------------------------------------------------------------------------
 * synthetic tables, generated on the device (bench / tests; BASELINE.json's configs)
 * ---------------------------------------------------------------------------------------- */

typedef struct evqgpu_synth_column {
  const char* name;
  uint32_t logical_type;   /* EVQ_COL_UNSIGNED_INT | EVQ_COL_DATETIME | EVQ_COL_FLOAT | EVQ_COL_BOOLEAN */
  uint32_t encoding;       /* EVQ_ENC_* */
  uint32_t null_every;     /* 0 = required column; k>0 = optional, row i is NULL iff i % k == k-1 */
  uint32_t transform;      /* 0: v = lo + r % span;  1: v = splitmix64(lo + r % span)  (full-range keys);
                              2 (float): v = (double)(lo + r % span) / 100.0 */
  uint64_t seed;           /* r = splitmix64(seed + row) */
  uint64_t lo;
  uint64_t span;           /* >= 1 */
} evqgpu_synth_column;

EVQGPU_API int evqgpu_table_synthesize(evqgpu_ctx* ctx, uint64_t num_rows, uint64_t row_offset,
                                       const evqgpu_synth_column* cols, uint32_t ncols, evqgpu_table** out);

/* ------------------------------------------------------------------------------------------
 * expressions: postfix programs, the shape of csql::vm::Program (sql/runtime/vm.h:44-75)
 * ---------------------------------------------------------------------------------------- */

enum {
  EVQ_X_INPUT = 4,    /* push input column `arg` (index into evqgpu_query_desc.input_columns), keeps its NULL tag */
  EVQ_X_LITERAL = 3,  /* push literal: imm = raw 64-bit value (u64 / i64 / double bits / bool);
                         STRING: imm = (offset << 32 | length) into evqgpu_expr.strings */
  EVQ_X_CALL = 1,     /* pop nargs, push fn(args); arg = function id from evqgpu_function_lookup.
                         Pure functions drop NULL tags; aggregate functions mark the
                         split point between the accumulate and the get program (compiler.cc:67-100) */
  EVQ_X_IF = 6        /* pop else, then, cond (pushed in the order cond, then, else); lazily evaluated
                         like the X_CJUMP/X_JUMP form of compiler.cc:174-209 */
};

typedef struct evqgpu_insn {
  uint8_t op;      /* EVQ_X_* */
  uint8_t type;    /* result SType of this node */
  uint16_t nargs;  /* EVQ_X_CALL only */
  uint32_t arg;
  uint64_t imm;
} evqgpu_insn;

typedef struct evqgpu_expr {
  const evqgpu_insn* code;
  uint32_t len;            /* 0 = absent */
  const char* strings;     /* string literal pool, may be NULL */
  uint32_t strings_len;
} evqgpu_expr;

/* Resolve a reference symbol string ("lt#bool/uint64;uint64;", "sum#uint64/uint64;",
 * "count#uint64/nil;" ...) to a function id; -1 if the device path does not implement it.
 * Besides the reference's registry (sql/defaults.cc:38-171) the typed extension aggregates
 * min / max / mean / sum#float64 of oracle/ref_tools/ext_aggregates.cc are known. */
EVQGPU_API int evqgpu_function_lookup(const char* symbol);
EVQGPU_API const char* evqgpu_function_symbol(int function_id); /* NULL if out of range */
EVQGPU_API int evqgpu_function_is_aggregate(int function_id);

/* ------------------------------------------------------------------------------------------
 * queries: fused FastCSTableScan (+ GroupByExpression)
 * ---------------------------------------------------------------------------------------- */

#define EVQGPU_QUERY_GROUPBY 1u       /* aggregate plan: select items are GroupByNode select expressions */
#define EVQGPU_QUERY_PARTIAL 2u       /* results stay as mergeable partials until evqgpu_query_merge */
#define EVQGPU_QUERY_WIRE 4u          /* aggregate plan whose groups are fetched in the reference's partial-aggregation row format
                                         (evqgpu_query_fetch_partial): group key hashes and raw states are kept besides the rows */

#define EVQGPU_QUERY_COORDINATOR 8u   /* aggregate plan that scans nothing: the coordinator side of a cluster GROUP BY
                                         (GroupByMergeExpression), fed with the shards' partial rows - evqgpu_query_merge_rows */

typedef struct evqgpu_query_desc {
  uint32_t struct_size;                 /* sizeof(evqgpu_query_desc) */
  uint32_t flags;                       /* EVQGPU_QUERY_* */
  uint32_t num_input_columns;
  const char* const* input_columns;     /* SequentialScanNode::selectedColumns() */
  evqgpu_expr where;                    /* BOOL program or len == 0 */
  uint32_t num_group;                   /* GROUP BY expressions (0 with EVQGPU_QUERY_GROUPBY = one global group) */
  const evqgpu_expr* group;
  uint32_t num_select;
  const evqgpu_expr* select;            /* output columns; at most one aggregate call each */
  uint64_t expected_groups;             /* hint, 0 = unknown */
} evqgpu_query_desc;

EVQGPU_API int evqgpu_query_create(evqgpu_ctx* ctx, const evqgpu_query_desc* desc, evqgpu_query** out);
EVQGPU_API void evqgpu_query_destroy(evqgpu_query* q);

EVQGPU_API uint32_t evqgpu_query_num_columns(const evqgpu_query* q);
EVQGPU_API uint32_t evqgpu_query_column_type(const evqgpu_query* q, uint32_t idx); /* SType */

/* Scan the given tables (partitions of one logical table; each must have the query's input
 * columns loaded) and aggregate / project.  Resets previous results of q.  Work is issued on
 * the context stream; the call returns once the result row count is known. */
EVQGPU_API int evqgpu_query_execute(evqgpu_query* q, evqgpu_table* const* tables, uint32_t ntables);

/* Multi-rank jobs (evqgpu_comm_init with nranks > 1) and EVQGPU_QUERY_PARTIAL aggregate plans: a COLLECTIVE call.  Every
 * rank calls it - in the same order relative to its other collective calls (prepare, merge) - whenever the set of tables it
 * is about to scan changes.  In one all-gather that every rank enters no matter what happened locally, the ranks agree
 * on what must be identical everywhere before the scan: the aggregation strategy, the key -> slot assignment of the dense
 * tiers (from the key bounds of all ranks) and the aggregate state layout.  A rank whose local part failed
 * (a missing column, a key expression that divides by zero) still takes part, and the call then fails on EVERY rank.
 * evqgpu_query_execute calls it implicitly on every execution - so execute is itself collective in a multi-rank job -
 * while evqgpu_query_enqueue and evqgpu_query_merge never decide per rank whether to enter a collective: enqueue requires a
 * preceding prepare for exactly these tables and fails with EVQGPU_ERR_ARG otherwise.  Everywhere else: a no-op. */
EVQGPU_API int evqgpu_query_prepare(evqgpu_query* q, evqgpu_table* const* tables, uint32_t ntables);

/* Same, but only enqueues the device work (no host synchronisation); finish with
 * evqgpu_query_finish.  Lets a caller time the device portion with CUDA events on
 * evqgpu_ctx_stream(). */
EVQGPU_API int evqgpu_query_enqueue(evqgpu_query* q, evqgpu_table* const* tables, uint32_t ntables);
EVQGPU_API int evqgpu_query_finish(evqgpu_query* q);

EVQGPU_API int evqgpu_query_num_rows(evqgpu_query* q, uint64_t* out);

/* Copy result rows [row0, row0 + max_rows) to the host in the packed SVector encoding:
 * numeric columns 9 B per row ([8 B value][1 B STag]), BOOL 2 B per row.  columns[i] receives
 * column i and must hold max_rows * elem_size bytes.  *nrows_out = rows written (0 = EOF).
 * Group order is unspecified, like the reference's. */
EVQGPU_API int evqgpu_query_fetch(evqgpu_query* q, uint64_t row0, uint64_t max_rows, void* const* columns,
                                  uint64_t* nrows_out);

/* String result columns (evqgpu_query_column_type == EVQ_STRING: a bare string column as GROUP BY key or select item).  The
 * device result - and evqgpu_query_fetch - holds the column as [dictionary code u64][tag]; this call returns rows
 * [row0, row0 + max_rows) as packed STRING elements, [u32 length][bytes][tag] (sql/svalue.cc:533-549; NULL = length 0 + tag
 * EVQ_STAG_NULL).  *nbytes_out = bytes needed; when dst == NULL or cap is smaller nothing is written and *nrows_out = 0.
 *
 * String support of a plan (no CPU fallback for the rest - EVQGPU_ERR_UNSUPPORTED): eq / neq between string columns and
 * string literals (expressions/boolean.cc:235-257, 355-377: the NULL tag is dropped, a NULL compares as ""), bare string
 * columns as GROUP BY expressions and select items - these run on dictionary codes shared by all tables of a context;
 * lt / lte / gt / gte (strncmp over the shorter length, then the lengths) and startswith / endswith between a string
 * column and a literal - evaluated once per dictionary entry, applied as a 1-byte-per-row verdict column.  In a
 * multi-rank job evqgpu_query_prepare synchronises the ranks' dictionaries first, so string keys and predicates merge like
 * integers; in the partial-aggregation row format (evqgpu_query_fetch_partial) the key hash and a selected string key carry the
 * string bytes as the reference's do.  LIKE raises "not yet implemented" in the reference and is refused here. */
EVQGPU_API int evqgpu_query_fetch_strings(evqgpu_query* q, uint32_t column, uint64_t row0, uint64_t max_rows, void* dst,
                                          uint64_t cap, uint64_t* nrows_out, uint64_t* nbytes_out);

/* The groups of an executed EVQGPU_QUERY_WIRE plan as the rows csql::PartialGroupByExpression::nextBatch produces
 * (sql/statements/select/groupby.cc:411-445) - what a shard of a cluster query returns to GroupByMergeExpression
 * (groupby.cc:553-615) and stores in its query cache:
 *   keys:  20 bytes per group, SHA-1 of the group expressions' packed values and tags, last expression first
 *          (groupby.cc:112-135);
 *   data:  per group the select items in order - an aggregate item as its function's saved state (count / sum: varuint,
 *          aggregate.cc:52-58, 200-206; the extension aggregates min / max {value, seen}, mean {double sum, count},
 *          sum<float64> raw, oracle/ref_tools/ext_aggregates.cc), any other item as SValue::encode (svalue.cc:306-309:
 *          type byte, length, packed value);  data_offsets[i] .. data_offsets[i + 1] is group i's slice.
 * Groups [row0, row0 + max_rows); *nrows_out = groups written, *data_bytes_out = bytes needed (nothing is written when
 * data_cap is too small).  count_distinct travels as its value set (size, then the members in ascending order,
 * aggregate.cc:110-116). */
EVQGPU_API int evqgpu_query_fetch_partial(evqgpu_query* q, uint64_t row0, uint64_t max_rows, void* keys, void* data,
                                          uint64_t data_cap, uint64_t* data_offsets, uint64_t* nrows_out,
                                          uint64_t* data_bytes_out);

/* The query cache entry PartialGroupByExpression::execute stores for its groups (sql/statements/select/groupby.cc:411-432,
 * read back by :262-292; sql/runtime/query_cache.cc:58-75):  u8 0x01 | u64 number of groups | per group: 20-byte group key |
 * saved states of the select items - the keys / data slices evqgpu_query_fetch_partial returns.  Pure host function.
 * *nbytes_out = bytes needed; nothing is written when dst == NULL or cap is smaller. */
EVQGPU_API int evqgpu_partial_cache_encode(const void* keys, const void* data, const uint64_t* data_offsets, uint64_t ngroups,
                                           void* dst, uint64_t cap, uint64_t* nbytes_out);
/* File name of the entry inside the cache directory: hex(SHA1(hex(input cache key) + hex(expression fingerprint))) + ".qc"
 * (groupby.cc:474-483, query_cache.cc:45,64); both keys are 20 bytes, out receives 44 bytes incl. the NUL. */
EVQGPU_API int evqgpu_partial_cache_filename(const void* input_cache_key, const void* expression_fingerprint, char* out,
                                             uint64_t cap);
/* Write the groups of an executed EVQGPU_QUERY_WIRE plan as such an entry to `path` (temporary file + rename, like
 * QueryCache::storeEntry). */
EVQGPU_API int evqgpu_query_store_cache(evqgpu_query* q, const char* path);

/* The QUERY_PARTIALAGGR_RESULT frames a shard answers a coordinator with (transport/native/ops/query_partialaggr.cc:83-124):
 * per frame an 8-byte big-endian header {u16 opcode 0x0102, u16 flags (1 = end of request on the last), u32 payload length}
 * (transport/native/connection_tcp.cc:238-251) and the payload varuint 0 | varuint rows | per row 20-byte key + saved states
 * (frames/query_partialaggr_result.cc:56-60).  A frame closes once its body exceeds soft_max_body (0 = the reference's
 * 8 MiB).  Pure host function over evqgpu_query_fetch_partial's output; all frames are written back to back. */
EVQGPU_API int evqgpu_partial_frames_encode(const void* keys, const void* data, const uint64_t* data_offsets, uint64_t ngroups,
                                            uint64_t soft_max_body, void* dst, uint64_t cap, uint64_t* nbytes_out,
                                            uint64_t* nframes_out);

/* Reading partial rows back - the coordinator's side (GroupByMergeExpression, groupby.cc:553-615; the cache load of the
 * partial operator, groupby.cc:262-292).  A body of `20-byte key | saved states` rows can only be walked with the plan: the
 * functions parse `desc` (the partial GROUP BY plan the shard ran) and report where the rows are.  Pure host functions,
 * parsing only.
 *   evqgpu_partial_rows_split    body = rows back to back; row_offsets[0 .. n] receives the row starts and the end
 *                                (row i = [row_offsets[i], row_offsets[i + 1]): 20 bytes key, the rest saved states)
 *   evqgpu_partial_cache_decode  a .qc entry (header checked against the rows found); offsets are into `entry`
 *   evqgpu_partial_frames_decode QUERY_PARTIALAGGR_RESULT frames back to back (every frame's row count checked);
 *                                row_offsets[2 i], [2 i + 1] = start and end of row i in `frames` (rows of different
 *                                frames are not adjacent); *end_of_request_out = 1 when the last frame carried the flag
 * *nrows_out = rows found; nothing is written when row_offsets == NULL or cap_rows is smaller.  EVQGPU_ERR_FORMAT on
 * truncated or inconsistent input. */
EVQGPU_API int evqgpu_partial_rows_split(const evqgpu_query_desc* desc, const void* body, uint64_t nbytes, uint64_t* row_offsets,
                                         uint64_t cap_rows, uint64_t* nrows_out);
EVQGPU_API int evqgpu_partial_cache_decode(const evqgpu_query_desc* desc, const void* entry, uint64_t nbytes, uint64_t* row_offsets,
                                           uint64_t cap_rows, uint64_t* nrows_out);
EVQGPU_API int evqgpu_partial_frames_decode(const evqgpu_query_desc* desc, const void* frames, uint64_t nbytes,
                                            uint64_t* row_offsets, uint64_t cap_rows, uint64_t* nrows_out, uint64_t* nframes_out,
                                            int* end_of_request_out);

/* The coordinator's side on the device: csql::GroupByMergeExpression (sql/statements/select/groupby.cc:528-637).  `q` is created
 * from the partial GROUP BY plan the shards ran, with EVQGPU_QUERY_GROUPBY | EVQGPU_QUERY_COORDINATOR; it scans no table.
 * evqgpu_query_merge_rows hands it rows a shard - the reference's CPU PartialGroupByExpression or a GPU shard's
 * evqgpu_query_fetch_partial - returned: row i is bytes [row_starts[i], row_ends[i]) of `base`, 20 bytes of SHA-1 group key +
 * the saved states of the select items (what evqgpu_partial_rows_split / _cache_decode / _frames_decode locate; for the first
 * two pass row_offsets and row_offsets + 1).  The states are parsed (loadInstanceState, groupby.cc:577-612; SValue::decode for
 * non-aggregate items) and kept.  evqgpu_query_merge_finish merges them on the device - a hash table keyed by the 20-byte
 * group key, the aggregates' merge functions as atomics (count / sum: +, min / max over the seen ones, mean: sums and counts
 * add; a non-aggregate item: the first row's value - all rows of a group carry the same one when it is a function of the
 * key) - and evaluates every select item's `get` side per group; the rows are then fetched like any result
 * (evqgpu_query_num_rows / _fetch / _order_by / _limit).  A string item's value enters the context's dictionary and is fetched with evqgpu_query_fetch_strings like a
 * string column's.  count_distinct states (value sets) are united per group and counted. */
EVQGPU_API int evqgpu_query_merge_rows(evqgpu_query* q, const void* base, const uint64_t* row_starts, const uint64_t* row_ends,
                                       uint64_t nrows);
EVQGPU_API int evqgpu_query_merge_finish(evqgpu_query* q);

/* ORDER BY over the result rows of an executed (and, for multi-rank jobs, merged) query, on the device:
 * csql::OrderByExpression (sql/statements/select/orderby.cc:58-160) with sort expressions that are columns of the
 * result (what the planner hands the operator: it appends hidden select items for anything else).  Values compare like
 * the reference's typed `cmp` functions (unsigned / signed / double / bool; NULL tags are ignored, a NULL sorts as its
 * value bits 0).  Rows with equal sort keys keep their previous order (the reference's std::sort leaves it unspecified).
 * Later fetches see the new order. */
typedef struct evqgpu_sort_spec {
  uint32_t column;       /* result column */
  uint32_t descending;   /* 0 = ASC */
} evqgpu_sort_spec;
EVQGPU_API int evqgpu_query_order_by(evqgpu_query* q, const evqgpu_sort_spec* specs, uint32_t nspecs);

/* LIMIT / OFFSET: csql::LimitExpression (sql/statements/select/limit.cc:43-112): keep result rows
 * [offset, offset + limit). */
EVQGPU_API int evqgpu_query_limit(evqgpu_query* q, uint64_t limit, uint64_t offset);

/* Statistics of the last execute. */
typedef struct evqgpu_query_stats {
  uint64_t rows_scanned;
  uint64_t rows_passed;          /* rows that satisfied WHERE */
  uint64_t algorithmic_bytes;    /* payload bytes of the referenced streams */
  uint64_t num_groups;
  uint32_t kernel_launches;      /* device kernels launched by the last execute */
  uint32_t strategy;             /* 0 scan-only, 1 register/shared-memory low-cardinality, 2 global hash table,
                                    3 direct-addressed group array in global memory (key tuples spanning a small box),
                                    4 global hash table filled by partitioned aggregation (records partitioned by home slot over
                                    two levels, every table slice aggregated in shared memory) */
  float jit_ms;                  /* time spent making the specialised kernel loadable in the last execute: NVRTC, or reading the
                                    cubin from the on-disk cache ($EVQGPU_CACHE_DIR, default ~/.cache/evqgpu); 0 = already loaded */
  float scan_ms;                 /* summed device time of the scan kernel launches since the last finish
                                    (only with evqgpu_ctx_set_profiling) */
  uint32_t scan_launches;        /* number of scan kernel launches scan_ms covers */
  uint32_t jit_disk_hits;        /* kernels of the last execute that came from the on-disk cubin cache instead of NVRTC */
} evqgpu_query_stats;
EVQGPU_API int evqgpu_query_get_stats(evqgpu_query* q, evqgpu_query_stats* out);

/* The generated CUDA source of the query's scan kernel (debugging, profiles/). */
EVQGPU_API const char* evqgpu_query_kernel_source(evqgpu_query* q);

/* ------------------------------------------------------------------------------------------
 * multi-GPU: one process per GPU, partial aggregate tables merged over NVLink with NCCL
 * ---------------------------------------------------------------------------------------- */

#define EVQGPU_COMM_ID_BYTES 128
/* rank 0 creates the id; the launcher broadcasts it (torch.distributed / MPI / file). */
EVQGPU_API int evqgpu_comm_unique_id(void* id_out /* EVQGPU_COMM_ID_BYTES */);
EVQGPU_API int evqgpu_comm_init(evqgpu_ctx* ctx, const void* id, int rank, int nranks);
EVQGPU_API int evqgpu_comm_destroy(evqgpu_ctx* ctx);

/* GroupByMergeExpression: combine the partial aggregates of all ranks (count/sum: +, min/max:
 * min/max over non-empty, mean: (sum, n) pairs).  Low-cardinality results are all-reduced and
 * every rank ends with the full result; high-cardinality results are repartitioned by key hash
 * (all-to-all) and every rank ends with its share of the groups. */
EVQGPU_API int evqgpu_query_merge(evqgpu_query* q);

/* ------------------------------------------------------------------------------------------
 * build-time check (no device needed): generate the scan kernel text for a plan against a
 * described column layout and compile it with NVRTC for sm_100a.
 * ---------------------------------------------------------------------------------------- */
typedef struct evqgpu_debug_column {
  uint32_t sql_type;   /* EVQ_* SType */
  uint32_t encoding;   /* EVQ_ENC_* */
  uint32_t dlevel_max; /* 0 = required */
  uint32_t value_bits;   /* column statistic: every value < 2^value_bits; 0 = unknown (64) */
  uint32_t leb_max_len;  /* column statistic: longest LEB128 value in bytes; 0 = unknown (10) */
  uint32_t reserved;
  uint64_t value_min;    /* column statistics: value range; value_max 0 = unknown (2^value_bits - 1) */
  uint64_t value_max;
} evqgpu_debug_column;

/* tier: 1 = dense / single group (dense_slots groups), 2 = global hash table, 4 = global hash table filled by partitioned
 * aggregation; ignored for scan-only plans.
 * src_out (may be NULL) receives the NUL-terminated kernel text when src_cap suffices; *src_len_out its length.
 * compile != 0 runs NVRTC; *cubin_bytes_out receives the cubin size. */
EVQGPU_API int evqgpu_debug_generate(const evqgpu_query_desc* desc, const evqgpu_debug_column* columns, uint32_t tier,
                                     uint32_t dense_slots, char* src_out, uint64_t src_cap, uint64_t* src_len_out,
                                     int compile, uint64_t* cubin_bytes_out);

#ifdef __cplusplus
}
#endif
#endif /* EVQGPU_H */
