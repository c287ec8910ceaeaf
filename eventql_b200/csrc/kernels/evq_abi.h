// evq_abi.h - types, constants and kernel parameter blocks shared by the host library (g++) and the
// device code (nvcc / NVRTC).  Plain C structs only; this text is pasted in front of every JIT kernel.
#ifndef EVQ_ABI_H
#define EVQ_ABI_H

typedef unsigned char u8;
typedef unsigned short u16;
typedef unsigned int u32;
typedef int i32;
typedef unsigned long long u64;
typedef long long i64;
typedef double f64;

#ifndef EVQ_TILE_ROWS
#define EVQ_TILE_ROWS 1024
#endif

#define EVQ_KIND_PLAIN64 0
#define EVQ_KIND_PLAIN32 1
#define EVQ_KIND_BITPACK 2
#define EVQ_KIND_LEB128 3
#define EVQ_KIND_LEVEL 4
#define EVQ_KIND_SUBIDX 5   // per-tile table of decode entry points (u16 per EVQ_SUB_GRAN values), not column data
#define EVQ_KIND_FILTER 6   // external row filter of the table: 1 bit per row (128 bytes per row tile)
#define EVQ_SUB_GRAN 4     // the sub-index of a variable-length LEB128 column records the start of every 4th value
#define EVQ_SUB_ENTRIES (EVQ_TILE_ROWS / EVQ_SUB_GRAN)

#define EVQ_ERR_DIV_ZERO 1u
#define EVQ_ERR_MOD_ZERO 2u
#define EVQ_ERR_TABLE_FULL 4u
#define EVQ_ERR_SLOT_RANGE 8u
#define EVQ_ERR_STAGE_OVERFLOW 16u
#define EVQ_ERR_PART_FULL 64u      // partitioned aggregation: a record partition overflowed (heavily skewed keys)
#define EVQ_ERR_PEER_TIMEOUT 32u   // the fused merge tail waited in vain for a peer rank's state

#define EVQ_MAX_STREAMS 32
#define EVQ_MAX_KEYS 8
#define EVQ_MAX_DISTINCT 4   // count_distinct aggregates per query

// ---- kernel parameter blocks ----------------------------------------------------------------------------------

struct EvqStream {
  const u8* base;        // logical stream (pages concatenated), 16-byte aligned, padded
  const u64* off_index;  // LEB128: byte offset of the first value of every row tile [ntiles + 1]
  const u64* val_index;  // optional column: non-NULL values before every row tile [ntiles + 1]
  u64 nbytes;            // payload bytes
  u32 kind;              // EVQ_KIND_*
  u32 bits;              // bit width (BITPACK / LEVEL)
  u32 smem_off;          // byte offset of this stream inside a pipeline stage
  u32 smem_cap;          // bytes reserved in a stage
};

// Global group table (tier 2): open addressing, linear probing, array of slots.  One slot = `stride` 64-bit words:
//   [0]                fingerprint word: bit 0 = occupied, bit 1 = the claiming thread still writes the keys,
//                      bits 2..9 = NULL tags of keys 0..7, bits 10..63 = hash bits
//   [1 .. nkeys]       raw 64-bit key values
//   [1 + nkeys ..]     aggregate state words
// so that probing a group, comparing its key and updating its aggregates touch ONE 64-byte DRAM atom for the usual
// <= 7 payload words (the table is far larger than L2: every distinct line touched per row is HBM traffic).
struct EvqHashTable {
  u64* slots;
  u64 cap;       // slots, power of two
  u32 stride;    // words per slot, multiple of 4
  u32 nkeys;
};

struct EvqScanParams {
  u64 num_rows;
  u32 num_tiles;
  u32 num_streams;
  EvqStream streams[EVQ_MAX_STREAMS];
  // aggregation state
  u64* dense_state;                 // tier 1: [G1][NSTATE] merged across CTAs with atomics
  EvqHashTable ht;                  // tier 2
  EvqHashTable dt[EVQ_MAX_DISTINCT]; // count_distinct: sets of (group, value) pairs, one table per distinct argument
  u64 key_min[EVQ_MAX_KEYS];        // tier 1 dense slot = sum((key - min) * stride), NULL -> null_idx * stride
  u64 key_stride[EVQ_MAX_KEYS];
  u64 key_null_idx[EVQ_MAX_KEYS];
  u64 key_span[EVQ_MAX_KEYS];       // largest non-NULL index of key i (direct-addressed group array: checked at run time)
  u64 dense_slots;                  // number of valid dense slots (<= EVQ_G1)
  // scan-only output
  u64* tile_counts;                 // pass 1: rows passing WHERE per tile
  const u64* tile_out_base;         // pass 2: exclusive prefix of tile_counts (+ table base)
  u8* out_cols[EVQ_MAX_STREAMS];    // pass 2: packed SVector output per select item
  // bookkeeping
  u32* status;                      // [0] error bits, [1] unused
  u64* counters;                    // [0] rows passed
  u64 tile_row_base;                // first tile index of this table inside tile_counts
  u64 ord_base;                     // added to a row's ordinal (tile index * 1024 + row in tile): rank-major order of a multi-rank job
  // partitioned aggregation (tier 2, groups far beyond L2): pass 1 writes the rows that pass WHERE as records into
  // 2^part_bits flat partition arrays by the top bits of their home slot (a tile's records are gathered per partition in
  // shared memory, one run per partition and tile is claimed from the cursors); the passes behind it (evq_repart +
  // evq_agg_smem, or evq_agg_part) aggregate one table slice at a time in shared memory / L2
  u64* part_buf;                    // [2^part_bits][part_cap][EVQ_NREC] record words
  u32* part_cursor;                 // [2^part_bits] records appended to every partition (runs are claimed with atomicAdd)
  u64 part_cap;                     // records per partition
  u32 part_shift;                   // partition = home slot >> part_shift
  u32 part_bits;
};


#endif  // EVQ_ABI_H
