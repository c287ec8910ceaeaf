// evq_scan_kernel.cuh - the fused decode + WHERE + GROUP BY kernel (fixed part).
//
// Replaces, in one pass over the encoded column pages (reference paths under src/eventql/):
//   FastCSTableScan::nextBatch        sql/CSTableScan.cc:757-858   (fetchColumn*, WHERE, projection)
//   VM::evaluatePredicateVector       sql/runtime/vm.cc:231-272
//   GroupByExpression::execute        sql/statements/select/groupby.cc:69-185
//
// Structure (one persistent CTA per SM slot, EVQ_NCONS consumer threads + 1 producer warp):
//   producer warp : for every row tile (EVQ_TILE_ROWS rows) of this CTA, lane i issues ONE TMA bulk copy
//                   (cp.async.bulk, UBLKCP) of stream i's byte range of the tile into the next free
//                   pipeline stage and publishes it through an mbarrier (EVQ_NSTAGES-deep ring).
//   consumer warps: wait for the stage, cooperatively resolve what row alignment needs (presence ranks of
//                   optional columns via warp ballots; LEB128 value boundaries via a popcount scan, with
//                   fast paths when every value of the tile has the same width), then each thread decodes
//                   its rows from shared memory, evaluates the query's WHERE program and aggregates:
//                     tier 1: thread-private accumulators (registers for a single group, conflict-free
//                             shared memory for up to EVQ_G1 dense groups), merged once per CTA
//                     tier 2: global open-addressing table, atomics in L2
//                     tier 0/3: scan-only plans: count rows per tile / write the compacted projection
//
// The generated part in front of this file (csrc/codegen.cc) defines:
//   EVQ_NCONS EVQ_NSTAGES EVQ_NSTREAMS EVQ_TIER EVQ_G1 EVQ_NSTATE EVQ_NKEYS EVQ_NLEB EVQ_NNULL EVQ_HAS_PREP EVQ_MIN_CTAS
//   struct EvqRow; struct EvqPrep;
//   evq_prep_a / evq_prep_b / evq_prep_c   cooperative per-tile phases
//   evq_load_row                              decode one row from the staged tile
//   evq_where                                 WHERE program
//   evq_keys                                  GROUP BY expressions -> raw key tuple
//   evq_accumulate                            aggregate updates through EVQ_UPD(state, op, value)
//   evq_state_init / evq_state_flush          identities / merge of one accumulator set
//   evq_project                               scan-only: select list -> packed output

#define EVQ_NWARPS (EVQ_NCONS / 32)
#define EVQ_RPT (EVQ_TILE_ROWS / EVQ_NCONS)
#define EVQ_NTHREADS (EVQ_NCONS + 32)

struct EvqScratch {
  u32 wtot[EVQ_NLEB > 0 ? EVQ_NLEB : 1][EVQ_NWARPS];   // per-warp value counts for the boundary scan
  u32 pres[EVQ_NNULL > 0 ? EVQ_NNULL : 1][EVQ_TILE_ROWS / 32];   // presence bitmap of optional columns, row order
  u16 endpos[EVQ_NLEB > 0 ? EVQ_NLEB : 1][EVQ_TILE_ROWS];        // offset of every value's last byte
  u32 scan[EVQ_NWARPS];                                // scan-only: pass counts per warp
};

// ---- LEB128 boundary resolution -------------------------------------------------------------------------------------

// phase A: classify the tile (all 1-byte / uniform width / general) and count this thread's terminators
template <int S, int L>
__device__ __forceinline__ void evq_leb_phase_a(const EvqTile& T, const EvqScanParams& P, EvqScratch* scr,
                                                u32* flagword, EvqLebState& st, u32& count) {
  const EvqStreamDesc d = T.desc[S];
  st.base = P.streams[S].smem_off + d.delta;
  count = 0;
  if (d.nbytes == d.nvals) {        // every value is exactly one byte: value i = byte i
    st.mode = 1;
    return;
  }
  const u32 w = d.nvals ? d.nbytes / d.nvals : 0u;
  const bool cand = d.nvals && w * d.nvals == d.nbytes && w <= 10u;
  const u8* region = T.stage + P.streams[S].smem_off;
  const u32 tb = d.delta + d.nbytes;
  const u32 nchunks = (tb + 15u) >> 4;
  const u32 per = (nchunks + EVQ_NCONS - 1) / EVQ_NCONS;
  const u32 c0 = T.ctid * per;
  const u32 c1 = c0 + per < nchunks ? c0 + per : nchunks;
  const u64 pat = evq_uniform_pattern(w ? w : 1u);
  bool ok = true;
  for (u32 c = c0; c < c1; ++c) {
    const u32 m = evq_leb_chunk_mask(region, c, d.delta, tb);
    count += __popc(m);
    if (cand) {
      // bit k is expected iff (16c + k - delta) % w == w-1
      const u32 q0 = (16u * c + w * 16u - d.delta % w) % w;
      u32 e = (u32) (pat >> q0) & 0xffffu;
      const u32 pos = 16u * c;
      if (pos < d.delta) e &= ~((1u << (d.delta - pos)) - 1u);
      if (tb - pos < 16u) e &= (1u << (tb - pos)) - 1u;
      ok = ok && (e == m);
    }
  }
  st.mode = w;
  const bool bad = !cand || !ok;
  if (__any_sync(0xffffffffu, bad) && evq_lane() == 0) atomicOr(flagword, 1u << L);
}

// phase B (only when the tile is not uniform): exclusive scan of the counts, then every thread writes the end
// offset of the values terminating in its chunks.  Contains two consumer barriers.
template <int S, int L>
__device__ __forceinline__ void evq_leb_phase_b(const EvqTile& T, const EvqScanParams& P, EvqScratch* scr,
                                                EvqLebState& st, u32 count) {
  const u32 lane = evq_lane(), warp = T.ctid >> 5;
  u32 incl = count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (u32) o) incl += n;
  }
  if (lane == 31) scr->wtot[L][warp] = incl;
  evq_cons_sync();
  u32 base = incl - count;
  for (u32 w = 0; w < warp; ++w) base += scr->wtot[L][w];
  const EvqStreamDesc d = T.desc[S];
  const u8* region = T.stage + P.streams[S].smem_off;
  const u32 tb = d.delta + d.nbytes;
  const u32 nchunks = (tb + 15u) >> 4;
  const u32 per = (nchunks + EVQ_NCONS - 1) / EVQ_NCONS;
  const u32 c0 = T.ctid * per;
  const u32 c1 = c0 + per < nchunks ? c0 + per : nchunks;
  for (u32 c = c0; c < c1; ++c) {
    u32 m = evq_leb_chunk_mask(region, c, d.delta, tb);
    while (m) {
      const u32 k = __ffs(m) - 1;
      m &= m - 1;
      if (base < EVQ_TILE_ROWS) scr->endpos[L][base] = (u16) (16u * c + k - d.delta);
      ++base;
    }
  }
  st.mode = EVQ_LEB_GENERAL;
  evq_cons_sync();
}

// value `idx` (index among the tile's values of this stream) of a LEB128 column
template <int L>
__device__ __forceinline__ u64 evq_ld_leb(const EvqTile& T, const EvqScratch* scr, const EvqLebState& st, u32 idx) {
  const u8* pay = T.stage + st.base;
  if (st.mode == 1) return (u64) pay[idx];
  if (st.mode != EVQ_LEB_GENERAL) return evq_leb_decode(pay + st.mode * idx, st.mode);
  const u32 e = scr->endpos[L][idx];
  const u32 s = idx ? (u32) scr->endpos[L][idx - 1] + 1u : 0u;
  return evq_leb_decode(pay + s, e - s + 1u);
}

// ---- optional columns: presence bitmap via warp ballots, ranks via popcount ---------------------------------------------

template <int S, int N>
__device__ __forceinline__ void evq_null_phase_a(const EvqTile& T, const EvqScanParams& P, EvqScratch* scr, u32 dmax) {
  const u32 off = P.streams[S].smem_off;
  const u32 bits = P.streams[S].bits;
#pragma unroll
  for (int k = 0; k < EVQ_RPT; ++k) {
    const u32 r = k * EVQ_NCONS + T.ctid;
    bool present = false;
    if (r < T.rows) present = evq_ld_level(T, S, off, r, bits) == dmax;
    const u32 word = __ballot_sync(0xffffffffu, present);
    if (evq_lane() == 0) scr->pres[N][r >> 5] = word;
  }
}

// after the barrier: lane i of every warp keeps the number of present rows before 32-row word i
template <int N>
__device__ __forceinline__ u32 evq_null_prefix(const EvqScratch* scr) {
  const u32 lane = evq_lane();
  const u32 cnt = __popc(scr->pres[N][lane]);
  u32 incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (u32) o) incl += n;
  }
  return incl - cnt;
}

// presence + rank (index into the tile's values) of row r; must be called by all lanes of the warp
template <int N>
__device__ __forceinline__ bool evq_null_rank(const EvqScratch* scr, u32 prefix, u32 r, u32& rank) {
  const u32 lane = evq_lane();
  const u32 w = scr->pres[N][r >> 5];
  rank = __shfl_sync(0xffffffffu, prefix, (r >> 5) & 31u) + __popc(w & ((1u << lane) - 1u));
  return (w >> lane) & 1u;
}

// ---- the kernel ------------------------------------------------------------------------------------------------------

struct EvqSmemHeader {
  u64 full[4];
  u64 empty[4];
  u32 flags[4];   // [it & 3] bit l: LEB column l needs the general path in tile iteration `it`
  EvqStreamDesc desc[4][EVQ_NSTREAMS > 0 ? EVQ_NSTREAMS : 1];
};

#define EVQ_HDR_BYTES ((sizeof(EvqSmemHeader) + 127) & ~127)

// The generated row functions (struct EvqRow, struct EvqPrep, evq_prep_*, evq_load_row, evq_where, evq_keys,
// evq_accumulate_*, evq_state_*, evq_project) are pasted at the next line by csrc/query.cc.
//@@EVQ_GENERATED@@

// count_distinct (generated only when the query has such aggregates): per passing row, after its group is known
#ifdef EVQ_NDISTINCT
#define EVQ_DISTINCT_ROW(row, gid, state) evq_accumulate_distinct(row, gid, state, P, err)
#else
#define EVQ_DISTINCT_ROW(row, gid, state) ((void) 0)
#endif

extern "C" __global__ void __launch_bounds__(EVQ_NTHREADS, EVQ_MIN_CTAS)
evq_scan(const __grid_constant__ EvqScanParams P, const u32 stage_bytes) {
  extern __shared__ __align__(128) u8 evq_smem[];
  EvqSmemHeader* hdr = (EvqSmemHeader*) evq_smem;
  u8* stages = evq_smem + EVQ_HDR_BYTES;
  EvqScratch* scratch = (EvqScratch*) (stages + (size_t) EVQ_NSTAGES * stage_bytes);
#if EVQ_TIER == 1 && EVQ_G1 > 1
  u64* sacc = (u64*) ((u8*) scratch + ((2 * sizeof(EvqScratch) + 127) & ~127));
#endif

  const u32 tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < EVQ_NSTAGES; ++s) {
      evq_mbar_init(&hdr->full[s], 1);
      evq_mbar_init(&hdr->empty[s], EVQ_NWARPS);
    }
    evq_mbar_fence_init();
  }
  if (tid < 4) hdr->flags[tid] = 0;
  __syncthreads();

  const u32 first_tile = blockIdx.x;
  const u32 tile_step = gridDim.x;

  if (tid >= EVQ_NCONS) {
    // ===================== producer warp =====================
    u32 it = 0;
    for (u32 tile = first_tile; tile < P.num_tiles; tile += tile_step, ++it) {
      const u32 s = it % EVQ_NSTAGES;
      const u32 round = it / EVQ_NSTAGES;
      if (round > 0) evq_mbar_wait(&hdr->empty[s], (round - 1) & 1u);
      evq_producer_issue(P, tile, stages + (size_t) s * stage_bytes, hdr->desc[s], &hdr->full[s]);
    }
    return;
  }

  // ===================== consumer warps =====================
  u32 err = 0;
  u64 passed = 0;
  EvqTile T;
  T.ctid = tid;

#if EVQ_TIER == 1
#if EVQ_G1 > 1
#define EVQ_ACC(slot, st) sacc[((slot) * EVQ_NSTATE + (st)) * EVQ_NCONS + tid]
  for (u32 g = 0; g < EVQ_G1; ++g) evq_state_init_slot(sacc, g, tid);
#else
  u64 racc[EVQ_NSTATE];
  evq_state_init_regs(racc);
#define EVQ_ACC(slot, st) racc[st]
#endif
#endif

  u32 it = 0;
  for (u32 tile = first_tile; tile < P.num_tiles; tile += tile_step, ++it) {
    const u32 s = it % EVQ_NSTAGES;
    evq_mbar_wait(&hdr->full[s], (it / EVQ_NSTAGES) & 1u);
    T.stage = stages + (size_t) s * stage_bytes;
    T.desc = hdr->desc[s];
    T.row0 = (u64) tile * EVQ_TILE_ROWS;
    {
      const u64 rem = P.num_rows - T.row0;
      T.rows = rem < EVQ_TILE_ROWS ? (u32) rem : EVQ_TILE_ROWS;
    }
    T.parity = it & 1u;
    EvqScratch* scr = scratch + T.parity;

    EvqPrep prep;
#if EVQ_HAS_PREP
    evq_prep_a(T, P, scr, &hdr->flags[it & 3u], prep);
    evq_cons_sync();
    {
      const u32 flags = hdr->flags[it & 3u];
      // slot (it+2)&3 was last read two tiles ago by threads that have all passed this barrier since; its next
      // writers (phase A of tile it+2) run after the next barrier, which this thread joins only after the reset
      if (tid == 0) hdr->flags[(it + 2u) & 3u] = 0;
      evq_prep_b(T, P, scr, prep, flags);     // block-uniform: LEB boundary scans where needed
      evq_prep_c(T, P, scr, prep);            // presence prefixes
    }
#endif

#if EVQ_TIER == 0 || EVQ_TIER == 3
    // scan-only plans: ordered compaction (CSTableScan.cc:826-857 keeps table order)
    u32 tile_pass = 0;
    u64 out_base = 0;
#if EVQ_TIER == 3
    out_base = P.tile_out_base[P.tile_row_base + tile];
#endif
#endif

#pragma unroll
    for (int k = 0; k < EVQ_RPT; ++k) {
      const u32 r = k * EVQ_NCONS + tid;
      const bool valid = r < T.rows;
      EvqRow row;
      evq_load_row(T, P, scr, prep, valid ? r : 0u, row);
      bool pass = false;
      if (valid) pass = evq_where(row, err);
#ifdef EVQ_FILTER_STREAM   // the table's external row filter, ANDed with WHERE (CSTableScan.cc:826-833)
      pass = pass && ((evq_filter_byte(T, P, r >> 3) >> (r & 7u)) & 1u) != 0u;
#endif
#if EVQ_TIER == 1
      if (pass) {
        ++passed;
#if EVQ_G1 > 1
        u64 key[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
        u32 ktag[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
        evq_keys(row, key, ktag, err);
        const u32 g = evq_dense_slot(key, ktag, err);
        if (g != ~0u) {
          evq_accumulate_smem(row, sacc, g, tid, P.dense_state, err);
          EVQ_DISTINCT_ROW(row, (u64) g, P.dense_state + (u64) g * EVQ_NSTATE_ALL);
        }
#else
        evq_accumulate_regs(row, racc, P.dense_state, err);
        EVQ_DISTINCT_ROW(row, 0ull, P.dense_state);
#endif
      }
#elif EVQ_TIER == 2 && defined(EVQ_DENSE_GLOBAL)
      if (pass) {   // direct-addressed group array (see evq_scan_fast.cuh)
        u64 key[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
        u32 ktag[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
        evq_keys(row, key, ktag, err);
        const u64 g = evq_dense_slot_rt(key, ktag, P, err);
        if (g != ~0ull) {
          ++passed;
          u64* st = P.dense_state + g * EVQ_NSTATE_ALL;
          evq_accumulate_global(row, st, err);
          EVQ_DISTINCT_ROW(row, g, st);
        }
      }
#elif EVQ_TIER == 2
      if (pass) {
        ++passed;
        u64 key[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
        u32 ktag[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
        evq_keys(row, key, ktag, err);
        u64* sp = evq_ht_upsert<EVQ_NKEYS>(P.ht, key, ktag, (u64*) 0);
        if (!sp) {
          err |= EVQ_ERR_TABLE_FULL;
        } else {
          evq_accumulate_global(row, sp + 1 + EVQ_NKEYS, err);
          EVQ_DISTINCT_ROW(row, (u64) sp, sp + 1 + EVQ_NKEYS);
        }
      }
#else
      {
        // rank of this row among the passing rows of the tile: rows are visited in k-major order, i.e. rows
        // [k*NCONS, (k+1)*NCONS) precede those of k+1; inside one k the order is warp, then lane
        const u32 ballot = __ballot_sync(0xffffffffu, pass);
        const u32 lane = evq_lane(), warp = tid >> 5;
        if (lane == 0) scr->scan[warp] = __popc(ballot);
        evq_cons_sync();
        u32 before = tile_pass;
        u32 total = 0;
#pragma unroll
        for (int w = 0; w < EVQ_NWARPS; ++w) {
          const u32 c = scr->scan[w];
          if ((u32) w < warp) before += c;
          total += c;
        }
        evq_cons_sync();
#if EVQ_TIER == 3
        if (pass) evq_project(row, P, out_base + before + __popc(ballot & ((1u << lane) - 1u)), err);
#endif
        tile_pass += total;
        if (pass) ++passed;
      }
#endif
    }

#if EVQ_TIER == 0
    if (tid == 0) P.tile_counts[P.tile_row_base + tile] = tile_pass;
#endif

    __syncwarp();
    if (evq_lane() == 0) evq_mbar_arrive(&hdr->empty[s]);
  }

  // ===================== epilogue: merge this CTA's partial state =====================
#if EVQ_TIER == 1
#if EVQ_G1 > 1
  for (u32 g = 0; g < EVQ_G1; ++g) evq_state_flush_smem(sacc, g, tid, P.dense_state);
#else
  evq_state_flush_regs(racc, P.dense_state);
#endif
#endif
  // statistics + errors: one atomic per warp
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      passed += __shfl_xor_sync(0xffffffffu, passed, o);
      err |= __shfl_xor_sync(0xffffffffu, err, o);
    }
    if (evq_lane() == 0) {
      if (passed) atomicAdd(P.counters, passed);
      if (err) atomicOr(P.status, err);
    }
  }
}
