// evq_scan_fast.cuh - the fused decode + WHERE + GROUP BY kernel for tables whose referenced columns are all
// required (no definition levels): the layout production cstables have.  Optional columns take the
// general kernel of evq_scan_kernel.cuh.
//
// Same operator fusion as the general kernel (FastCSTableScan::nextBatch, sql/CSTableScan.cc:757-858 +
// VM::evaluatePredicateVector, sql/runtime/vm.cc:231-272 + GroupByExpression::execute,
// sql/statements/select/groupby.cc:69-185) and the same producer warp / TMA bulk-copy pipeline; what differs is the
// work assignment of the consumer threads, chosen so that the decode is a few straight-line instructions per value:
//
//   * thread t owns the EVQ_RPT = 4 CONSECUTIVE rows 4t..4t+3 of the 1024-row tile.  A 1-byte LEB128 column is then one
//     32-bit shared-memory load per thread, a plain column one vector load, and a variable-length LEB128 column a
//     sequential decode of 4 values from one known byte offset.
//   * the longest value of every LEB128 column (Column::leb_max_len, a statistic computed when the column is loaded)
//     is a compile-time constant L of the specialised kernel: L == 1 needs no boundary search at all; L <= 4 decodes
//     from a 32-bit window; a tile with nbytes == L * nvals holds only L-byte values (uniform fast path, decided from the
//     tile descriptor alone).  Otherwise the consumers locate the start of every 4th value with one popcount scan
//     (2 consumer barriers per tile, shared by all such columns).
//   * columns whose values are known to fit 32 bits are carried as u32, so the compiler narrows the query's
//     comparisons and multiplications (64-bit results are formed where the expression needs them).
//
// The generated part (csrc/codegen.cc) defines:
//   EVQ_NCONS EVQ_NSTAGES EVQ_NSTREAMS EVQ_TIER EVQ_G1 EVQ_NSTATE EVQ_NKEYS EVQ_NGEN EVQ_MIN_CTAS
//   struct EvqRow; struct EvqCols; struct EvqFastPrep;
//   evq_fast_prep      cooperative boundary search of the variable-length columns of the tile
//   evq_fast_decode    decode the thread's 4 rows of every referenced column into registers
//   evq_fast_row       pick row k out of EvqCols
//   evq_where / evq_keys / evq_accumulate_* / evq_state_* / evq_project   as in the general kernel

#define EVQ_NWARPS (EVQ_NCONS / 32)
#define EVQ_RPT (EVQ_TILE_ROWS / EVQ_NCONS)   // consecutive rows per thread: 4 (256 consumers) or 8 (128)
#define EVQ_NTHREADS (EVQ_NCONS + 32)

// EVQ_GEN_CHUNKS: 16-byte chunks a tile of the longest variable-length column can span (generated)
struct EvqFastScratch {
  u32 wtot[EVQ_NGEN > 0 ? EVQ_NGEN : 1][EVQ_NWARPS];          // terminators per consumer warp (boundary search)
  u32 chunk[EVQ_NGEN > 0 ? EVQ_NGEN : 1][EVQ_GEN_CHUNKS];     // per chunk: terminators before it in its warp << 16 | terminator mask
  u32 scan[EVQ_NWARPS];                                       // scan-only plans: pass counts per warp
  u32 nullw[2][EVQ_NNULL > 0 ? EVQ_NNULL : 1][EVQ_NWARPS];    // optional columns: present values per consumer warp (two tiles in flight)
#if defined(EVQ_NNV) && EVQ_NNV > 0
  // optional variable-length columns: the tile's present values, decoded by value ordinal (two tiles in flight)
  __align__(16) u32 nval[2][EVQ_NNV][EVQ_TILE_ROWS + 8];
#endif
#ifdef EVQ_PARTITION
  __align__(16) u64 prec[EVQ_MAX_PARTS * EVQ_PART_BIN * EVQ_NREC];   // the tile's records: one bin of EVQ_PART_BIN records per partition
  u32 phist[EVQ_MAX_PARTS];                                   // records of the current tile per partition
#endif
};

// ---- boundary search of a variable-length LEB128 column ------------------------------------------------------------------
// Thread t must find the byte at which value RPT*t starts, i.e. the byte behind terminator number RPT*t - 1 of the tile.
// Phase A (chunk-parallel, before the one consumer barrier): every thread takes `per` consecutive 16-byte chunks, builds
// their terminator masks, and a warp scan turns the counts into "terminators before this chunk inside my warp"; mask and
// count go to the chunk table, the warp total to wtot.  Phase B (row-parallel, after the barrier): pick the warp from the
// <= 8 warp totals, binary-search the chunk table of that warp, then drop set bits of the chunk's mask.

// A tile is uniform when all its values have the column's maximal length L (decided from the tile descriptor alone).
template <int S, int G, int L>
__device__ __forceinline__ void evq_fast_prep_a(const EvqTile& T, const EvqScanParams& P, EvqFastScratch* scr, bool& general) {
  const EvqStreamDesc d = T.desc[S];
  general = d.nbytes != (u32) L * d.nvals;   // the same for every thread of the CTA
  if (!general) return;
  // chunks are counted from the 16-byte block the tile's payload starts in (with several tiles per stage the payload of
  // a tile starts anywhere inside the stream's region)
  const u8* region = T.stage + P.streams[S].smem_off + (d.delta & ~15u);
  const u32 dl = d.delta & 15u;
  const u32 tb = dl + d.nbytes;
  const u32 nchunks = (tb + 15u) >> 4;
  const u32 per = (nchunks + EVQ_NCONS - 1) / EVQ_NCONS;
  const u32 c0 = T.ctid * per;
  const u32 c1 = c0 + per < nchunks ? c0 + per : nchunks;
  // the first two masks stay in registers (L <= 4: at most 2 chunks per thread with 128 consumers, 1 with 256)
  const u32 m0 = c0 < c1 ? evq_leb_chunk_mask(region, c0, dl, tb) : 0u;
  const u32 m1 = c0 + 1u < c1 ? evq_leb_chunk_mask(region, c0 + 1u, dl, tb) : 0u;
  u32 count = __popc(m0) + __popc(m1);
  for (u32 c = c0 + 2u; c < c1; ++c) count += __popc(evq_leb_chunk_mask(region, c, dl, tb));
  const u32 lane = evq_lane();
  u32 incl = count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (u32) o) incl += n;
  }
  if (lane == 31) scr->wtot[G][T.ctid >> 5] = incl;
  u32 before = incl - count;
  if (c0 < c1) scr->chunk[G][c0] = (before << 16) | m0;
  before += __popc(m0);
  if (c0 + 1u < c1) scr->chunk[G][c0 + 1u] = (before << 16) | m1;
  before += __popc(m1);
  for (u32 c = c0 + 2u; c < c1; ++c) {
    const u32 m = evq_leb_chunk_mask(region, c, dl, tb);
    scr->chunk[G][c] = (before << 16) | m;
    before += __popc(m);
  }
}

// byte offset (from the payload start) of the thread's first value
template <int S, int G>
__device__ __forceinline__ u32 evq_fast_prep_b(const EvqTile& T, const EvqScanParams& P, const EvqFastScratch* scr) {
  const EvqStreamDesc d = T.desc[S];
  const u32 first = EVQ_RPT * T.ctid;
  if (first == 0u || first >= d.nvals) return 0u;
  const u32 r = first - 1u;                        // number (in the tile) of the terminator in front of the value
  u32 wt[EVQ_NWARPS];
#pragma unroll
  for (int i = 0; i < EVQ_NWARPS - 1; ++i) wt[i] = scr->wtot[G][i];
  const u32 dl = d.delta & 15u;
  const u32 tb = dl + d.nbytes;
  const u32 nchunks = (tb + 15u) >> 4;
  const u32 per = (nchunks + EVQ_NCONS - 1) / EVQ_NCONS;
  const u32 span = 32u * per;                      // chunks per warp
  // terminators in front of chunk c: the warp-local count of the chunk table + the totals of the warps before it
  auto before = [&](u32 c, u32& entry) -> u32 {
    entry = scr->chunk[G][c];
    u32 base = entry >> 16;
#pragma unroll
    for (int i = 1; i < EVQ_NWARPS; ++i) base += c >= (u32) i * span ? wt[i - 1] : 0u;
    return base;
  };
  // values have nearly uniform length inside a tile: start at the interpolated chunk and walk (typically 0-1 steps)
  u32 lo = (u32) ((float) r * __fdividef((float) nchunks, (float) d.nvals));
  lo = lo < nchunks ? lo : nchunks - 1u;
  u32 e;
  u32 b = before(lo, e);
  while (b > r) { --lo; b = before(lo, e); }
  while (lo + 1u < nchunks && b + __popc(e & 0xffffu) <= r) { ++lo; b = before(lo, e); }
  u32 m = e & 0xffffu;
  for (u32 k = r - b; k > 0u; --k) m &= m - 1u;
  return 16u * lo + (u32) __ffs(m) - dl;           // (__ffs - 1) is the terminator's byte; the value starts behind it
}

// with a sub-index (Column::sub_index, staged with the tile as its own stream) nothing has to be searched
template <int S, int L>
__device__ __forceinline__ bool evq_fast_general(const EvqTile& T) {
  return T.desc[S].nbytes != (u32) L * T.desc[S].nvals;
}

// EVQ_RPT == 8, EVQ_SUB_GRAN == 4: two entries per thread (low half: value 8t, high half: value 8t + 4) in one word
template <int X>
__device__ __forceinline__ u32 evq_fast_substart(const EvqTile& T, const EvqScanParams& P) {
  return ((const u32*) (T.stage + P.streams[X].smem_off + T.desc[X].delta))[T.ctid];
}

// ---- per-thread decode of 4 consecutive values -------------------------------------------------------------------------

// index of the thread's first value inside the tile; threads past the end of a short tile decode (and discard) the
// bytes behind the payload, which the stage regions are padded for
__device__ __forceinline__ u32 evq_fast_first(const EvqTile& T, int s) {
  (void) s;   // required columns: every stream holds one value per row of the tile
  const u32 i = EVQ_RPT * T.ctid;
  return i < T.rows ? i : (T.rows & ~(EVQ_RPT - 1u));   // threads past the end re-read the last group (keeps word alignment)
}

// 4 bytes at byte offset `off` of a 128-byte aligned stage: two aligned ld.shared (32-bit address arithmetic) + funnel shift
__device__ __forceinline__ u32 evq_stage_u32(const EvqTile& T, u32 off) {
  const u32 a = T.stage_sa + (off & ~3u);
  u32 w0, w1;
  asm volatile("ld.shared.u32 %0, [%2];\n\tld.shared.u32 %1, [%2+4];" : "=r"(w0), "=r"(w1) : "r"(a));
  return __funnelshift_r(w0, w1, off << 3);   // the shift count is taken modulo 32
}

// vector loads: a thread's 8 consecutive values are 8 / 16 / 32 / 64 contiguous bytes; one wide load per thread keeps the
// warp's access conflict-free where 32-bit loads at an 8- or 16-byte lane stride would collide 2- or 4-way on the banks
__device__ __forceinline__ void evq_stage_v2(const EvqTile& T, u32 off, u32& a, u32& b) {   // off is a multiple of 8
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(T.stage_sa + off));
}
__device__ __forceinline__ void evq_stage_v4(const EvqTile& T, u32 off, u32& a, u32& b, u32& c, u32& d) {   // multiple of 16
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(T.stage_sa + off));
}

__device__ __forceinline__ u32 evq_stage_word(const EvqTile& T, u32 off) {   // off is a multiple of 4
  u32 w;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(T.stage_sa + off));
  return w;
}

// 7-bit groups -> value with integer dot products (the continuation bits of x must be cleared):
// byte0 + 128 * byte1 is one dp2a against the 16-bit weights (1, 128); a 4-byte value is two of them + one multiply-add
__device__ __forceinline__ u32 evq_leb_pack2(u32 x) { return __dp2a_lo(0x00800001u, x, 0u); }
__device__ __forceinline__ u32 evq_fast_pack4(u32 x) { return __dp2a_hi(0x00800001u, x, 0u) * 16384u + __dp2a_lo(0x00800001u, x, 0u); }

__device__ __forceinline__ u32 evq_fixed_mask(u32 len) { return len >= 4u ? 0x7f7f7f7fu : ((1u << (8u * len)) - 1u) & 0x7f7f7f7fu; }

// L == 1: value i is byte i
template <int S>
__device__ __forceinline__ void evq_fast_ld_leb1(const EvqTile& T, const EvqScanParams& P, u32 (&v)[EVQ_RPT]) {
  // every value is one byte: byte offset == row index, so a tile starts at a multiple of 1024 (delta is word aligned)
  const u32 off = P.streams[S].smem_off + T.desc[S].delta + evq_fast_first(T, S);
  u32 x[EVQ_RPT / 4];
  if (EVQ_RPT == 8) evq_stage_v2(T, off, x[0], x[EVQ_RPT / 4 - 1]);
  else x[0] = evq_stage_word(T, off);
#pragma unroll
  for (int j = 0; j < EVQ_RPT / 4; ++j) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[4 * j + i] = __byte_perm(x[j], 0u, 0x4440u + i);   // byte i, zero extended (one PRMT)
  }
}

// same, also keeping the raw bytes (4 rows per word) for the dp4a aggregates
template <int S>
__device__ __forceinline__ void evq_fast_ld_leb1p(const EvqTile& T, const EvqScanParams& P, u32 (&v)[EVQ_RPT], u32 (&packed)[EVQ_RPT / 4]) {
  const u32 off = P.streams[S].smem_off + T.desc[S].delta + evq_fast_first(T, S);
  if (EVQ_RPT == 8) evq_stage_v2(T, off, packed[0], packed[EVQ_RPT / 4 - 1]);
  else packed[0] = evq_stage_word(T, off);
#pragma unroll
  for (int j = 0; j < EVQ_RPT / 4; ++j) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[4 * j + i] = __byte_perm(packed[j], 0u, 0x4440u + i);
  }
}

// 4 bytes at the shared-window byte address a (the stages are 128-byte aligned: the shift is the same as for an offset)
__device__ __forceinline__ u32 evq_sa_u32(u32 a) {
  u32 w0, w1;
  asm volatile("ld.shared.u32 %0, [%2];\n\tld.shared.u32 %1, [%2+4];" : "=r"(w0), "=r"(w1) : "r"(a & ~3u));
  return __funnelshift_r(w0, w1, a << 3);
}

// 2 <= L <= 4: 32-bit window.  LMIN: no value of the column is shorter than LMIN bytes (from the column's minimum).
template <int S, int G, int L, int LMIN>
__device__ __forceinline__ void evq_fast_ld_leb32(const EvqTile& T, const EvqScanParams& P, bool general, u32 start,
                                                  u32 (&v)[EVQ_RPT]) {
  const u32 pay = P.streams[S].smem_off + T.desc[S].delta;
  if (!general) {
    const u32 p = pay + (u32) L * evq_fast_first(T, S);
    if ((L == 2 || L == 4) && EVQ_RPT == 8 && (p & 15u) == 0u) {
      // the thread's 8 values are 16 (L == 2) or 32 (L == 4) aligned bytes: 128-bit loads
      u32 w[L == 2 ? 4 : 8];
      evq_stage_v4(T, p, w[0], w[1], w[2], w[3]);
      if (L == 4) evq_stage_v4(T, p + 16u, w[L == 2 ? 0 : 4], w[L == 2 ? 1 : 5], w[L == 2 ? 2 : 6], w[L == 2 ? 3 : 7]);
#pragma unroll
      for (int i = 0; i < EVQ_RPT; ++i) {
        // two 2-byte values per word: dp2a_lo / dp2a_hi pick their byte pair; one 4-byte value per word
        if (L == 2) {
          const u32 ww = w[i >> 1] & 0x7f7f7f7fu;
          v[i] = (i & 1) ? __dp2a_hi(0x00800001u, ww, 0u) : __dp2a_lo(0x00800001u, ww, 0u);
        } else {
          v[i] = evq_fast_pack4(w[L == 2 ? 0 : i] & 0x7f7f7f7fu);
        }
      }
    } else if ((L == 2 || L == 4) && (p & 3u) == 0u) {   // whole words: two 2-byte values or one 4-byte value each
#pragma unroll
      for (int i = 0; i < EVQ_RPT; ++i) {
        if (L == 2) {
          const u32 w = evq_stage_word(T, p + 4 * (i >> 1)) & 0x7f7f7f7fu;
          v[i] = (i & 1) ? __dp2a_hi(0x00800001u, w, 0u) : __dp2a_lo(0x00800001u, w, 0u);
        } else {
          v[i] = evq_fast_pack4(evq_stage_word(T, p + 4 * i) & 0x7f7f7f7fu);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < EVQ_RPT; ++i) {
        const u32 x = evq_stage_u32(T, p + L * i) & evq_fixed_mask(L);
        v[i] = L == 2 ? evq_leb_pack2(x) : evq_fast_pack4(x);
      }
    }
  } else {
    // G == 2: `start` packs two entry points from the sub-index (value 0 and value EVQ_RPT / 2 of the thread), decoded as
    // two independent chains; G == 1: one entry point (searched by evq_fast_prep), one chain over all values
    u32 p[2] = {T.stage_sa + pay + (G == 2 ? (start & 0xffffu) : start), T.stage_sa + pay + (start >> 16)};
#pragma unroll
    for (int i = 0; i < EVQ_RPT / G; ++i) {
#pragma unroll
      for (int h = 0; h < G; ++h) {
        const u32 x = evq_sa_u32(p[h]);
        u32 y;
        if (LMIN == L - 1) {
          // every value has L - 1 or L bytes: the continuation bit of byte L - 2 says which
          const u32 cont = (x >> (8 * (L - 1) - 1)) & 1u;
          y = x & (evq_fixed_mask(L - 1) + cont * (0x7fu << (8 * (L - 1))));
          p[h] += (u32) (L - 1) + cont;
        } else {
          const u32 tm = ~x & 0x80808080u;           // terminator bits of the window
          const u32 msk = tm ^ (tm - 1u);            // every bit up to and including the first of them
          y = x & msk & 0x7f7f7f7fu;
          p[h] += __popc(msk) >> 3;
        }
        v[h * (EVQ_RPT / 2) + i] = L == 2 ? evq_leb_pack2(y) : evq_fast_pack4(y);
      }
    }
  }
}

// The thread's 64 contiguous, 16-byte aligned bytes (8 values of 8 bytes; EVQ_RPT == 8) without bank conflicts: at a
// 64-byte lane stride four 128-bit loads at the same offset collide 4-way inside every quarter warp (lanes t and t + 2
// start at the same bank), per-value 64-bit loads 8- to 16-way.  Lane t therefore reads its four 16-byte pieces in the
// rotated order (j + t / 2) & 3 - the 8 lanes of a quarter warp then cover all 32 banks - and rotates them back in
// registers (two rounds of selects).
__device__ __forceinline__ void evq_stage_64bytes(const EvqTile& T, u32 off, u32 (&w)[16]) {
  const u32 r = (T.ctid >> 1) & 3u;
  u32 c[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) evq_stage_v4(T, off + 16u * ((j + r) & 3u), c[j][0], c[j][1], c[j][2], c[j][3]);
  // c[j] = piece (j + r) & 3; piece k = c[(k - r) & 3]
  const bool r1 = (r & 1u) != 0u, r2 = (r & 2u) != 0u;
  u32 e[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) e[k][i] = r1 ? c[(k + 3) & 3][i] : c[k][i];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) w[4 * k + i] = r2 ? e[(k + 2) & 3][i] : e[k][i];
}

// 5 <= L <= 10: 64-bit window (+ 2 bytes for 9- and 10-byte values)
template <int S, int G, int L>
__device__ __forceinline__ void evq_fast_ld_leb64(const EvqTile& T, const EvqScanParams& P, bool general, u32 start,
                                                  u64 (&v)[EVQ_RPT]) {
  const u8* pay = T.stage + P.streams[S].smem_off + T.desc[S].delta;
  if (L == 8 && EVQ_RPT == 8 && !general) {
    // every value has 8 bytes: the thread's 8 values are 64 contiguous bytes
    const u32 off = P.streams[S].smem_off + T.desc[S].delta + 8u * evq_fast_first(T, S);
    if ((off & 15u) == 0u) {
      u32 w[16];
      evq_stage_64bytes(T, off, w);
#pragma unroll
      for (int i = 0; i < EVQ_RPT; ++i)
        v[i] = (u64) evq_fast_pack4(w[2 * i] & 0x7f7f7f7fu) | ((u64) evq_fast_pack4(w[2 * i + 1] & 0x7f7f7f7fu) << 28);
      return;
    }
  }
  const u8* p;
  if (!general) p = pay + (u32) L * evq_fast_first(T, S);
  else p = pay + (G == 2 ? (start & 0xffffu) : start);   // (one chain: the second sub-index entry is not needed)
#pragma unroll
  for (int i = 0; i < EVQ_RPT; ++i) {
    u32 lo, hi;
    evq_lds_unaligned64(p, lo, hi);
    u32 len;
    if (!general) {
      len = L;
    } else {
      const u32 tl = ~lo & 0x80808080u, th = ~hi & 0x80808080u;
      if (tl) len = (32u - __clz(tl & (0u - tl))) >> 3;
      else if (th) len = 4u + ((32u - __clz(th & (0u - th))) >> 3);
      else len = (L > 9 && (p[8] & 0x80u)) ? 10u : 9u;
    }
    u64 val;
    if (len <= 4u) {
      val = evq_leb_pack4(lo & evq_fixed_mask(len));
    } else {
      val = (u64) evq_leb_pack4(lo & 0x7f7f7f7fu) | ((u64) evq_leb_pack4(hi & evq_fixed_mask(len - 4u)) << 28);
      if (L > 8 && len > 8u) {
        val |= ((u64) (p[8] & 0x7fu)) << 56;
        if (L > 9 && len > 9u) val |= ((u64) (p[9] & 0x7fu)) << 63;
      }
    }
    v[i] = val;
    p += len;
  }
}

template <int S>
__device__ __forceinline__ void evq_fast_ld_plain64(const EvqTile& T, const EvqScanParams& P, u64 (&v)[EVQ_RPT]) {
  const u32 off = P.streams[S].smem_off + T.desc[S].delta + 8u * evq_fast_first(T, S);
  if (EVQ_RPT == 8 && (off & 15u) == 0u) {   // 64 contiguous bytes per thread, bank-conflict free
    u32 w[16];
    evq_stage_64bytes(T, off, w);
#pragma unroll
    for (int i = 0; i < EVQ_RPT; ++i) v[i] = (u64) w[2 * i] | ((u64) w[2 * i + 1] << 32);
  } else if ((off & 15u) == 0u) {   // 16-byte vector loads (two values each)
#pragma unroll
    for (int i = 0; i < EVQ_RPT; i += 2) {
      u32 a, b, c, d;
      evq_stage_v4(T, off + 8u * i, a, b, c, d);
      v[i] = (u64) a | ((u64) b << 32);
      v[i + 1] = (u64) c | ((u64) d << 32);
    }
  } else {
    const u64* p = (const u64*) (T.stage + off);
#pragma unroll
    for (int i = 0; i < EVQ_RPT; ++i) v[i] = p[i];
  }
}

// the low halves only: for PLAIN64 columns whose values are known to fit 32 bits
template <int S>
__device__ __forceinline__ void evq_fast_ld_plain64_lo(const EvqTile& T, const EvqScanParams& P, u32 (&v)[EVQ_RPT]) {
  const u32 off = P.streams[S].smem_off + T.desc[S].delta + 8u * evq_fast_first(T, S);
  if (EVQ_RPT == 8 && (off & 15u) == 0u) {
    u32 w[16];
    evq_stage_64bytes(T, off, w);
#pragma unroll
    for (int i = 0; i < EVQ_RPT; ++i) v[i] = w[2 * i];
  } else if ((off & 15u) == 0u) {
#pragma unroll
    for (int i = 0; i < EVQ_RPT; i += 2) {
      u32 b, d;
      evq_stage_v4(T, off + 8u * i, v[i], b, v[i + 1], d);
    }
  } else {
    const u32* p = (const u32*) (T.stage + off);
#pragma unroll
    for (int i = 0; i < EVQ_RPT; ++i) v[i] = p[2 * i];
  }
}

template <int S>
__device__ __forceinline__ void evq_fast_ld_plain32(const EvqTile& T, const EvqScanParams& P, u32 (&v)[EVQ_RPT]) {
  const u32 off = P.streams[S].smem_off + T.desc[S].delta + 4u * evq_fast_first(T, S);
  if ((off & 15u) == 0u) {
#pragma unroll
    for (int i = 0; i < EVQ_RPT; i += 4) evq_stage_v4(T, off + 4u * i, v[i], v[i + 1], v[i + 2], v[i + 3]);
  } else {
    const u32* p = (const u32*) (T.stage + off);
#pragma unroll
    for (int i = 0; i < EVQ_RPT; ++i) v[i] = p[i];
  }
}

template <int S>
__device__ __forceinline__ void evq_fast_ld_bitpack(const EvqTile& T, const EvqScanParams& P, u32 (&v)[EVQ_RPT]) {
  const EvqStreamDesc& d = T.desc[S];
  const u32* words = (const u32*) (T.stage + P.streams[S].smem_off + d.delta);
  const u32 bits = P.streams[S].bits;
#pragma unroll
  const u32 i0 = evq_fast_first(T, S);
#pragma unroll
  for (int i = 0; i < EVQ_RPT; ++i) v[i] = evq_unpack_vertical(words, d.skew + i0 + i, bits);
}

// ---- optional columns (one definition-level bit per row) ------------------------------------------------------------------
// Only the present values are in the data stream (ColumnWriter.cc:92-100): thread t needs (a) which of its 8 rows are
// present and (b) the ordinal, inside the tile, of its first present value = present rows of the threads before it.

// presence bits of rows 8t .. 8t+7 (bit k = row 8t + k) from the tile's 1-bit level blocks (libsimdcomp vertical layout:
// value i of a 128-block is bit i / 4 of the block's word i % 4; 8t is a multiple of 8, so the thread's rows are bits
// b0 and b0 + 1 of the four words of one block)
template <int LS>
__device__ __forceinline__ u32 evq_fast_presence(const EvqTile& T, const EvqScanParams& P) {
  const u32 i0 = EVQ_RPT * T.ctid;
  const u32 off = P.streams[LS].smem_off + T.desc[LS].delta + (i0 >> 7) * 16u;
  u32 w0, w1, w2, w3;
  evq_stage_v4(T, off, w0, w1, w2, w3);
  const u32 b0 = (i0 & 127u) >> 2;
  const u32 x0 = w0 >> b0, x1 = w1 >> b0, x2 = w2 >> b0, x3 = w3 >> b0;
  u32 pb = (x0 & 1u) | ((x1 & 1u) << 1) | ((x2 & 1u) << 2) | ((x3 & 1u) << 3);
  if (EVQ_RPT == 8) pb |= ((x0 & 2u) << 3) | ((x1 & 2u) << 4) | ((x2 & 2u) << 5) | ((x3 & 2u) << 6);
  return pb;
}

// ordinal of the thread's first present value among the tile's values of the column: warp scan of the per-thread counts,
// warp totals through shared memory (`buf` alternates between consecutive tiles, so one barrier per tile suffices)
// (`cnt`: the thread's count, or the counts of up to three columns in 10-bit fields - the sums never leave their fields)
__device__ __forceinline__ u32 evq_fast_null_scan(u32 cnt, EvqFastScratch* scr, u32 buf, int slot, u32 tid) {
  u32 incl = cnt;
  const u32 lane = evq_lane();
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (u32) o) incl += n;
  }
  if (lane == 31) scr->nullw[buf][slot][tid >> 5] = incl;
  return incl - cnt;
}
__device__ __forceinline__ u32 evq_fast_null_rank(u32 excl, const EvqFastScratch* scr, u32 buf, int slot, u32 tid) {
  u32 r = excl;
#pragma unroll
  for (int w = 0; w < EVQ_NWARPS; ++w) r += (u32) w < (tid >> 5) ? scr->nullw[buf][slot][w] : 0u;
  return r;
}

// index, among the thread's present values, of row k's value
__device__ __forceinline__ u32 evq_pidx(u32 pb, int k) { return __popc(pb & ((1u << k) - 1u)); }

// prmt.b32 with its full 4-bit selector nibbles: bit 3 replicates the sign bit of the selected byte (the __byte_perm
// intrinsic only uses 3 bits per nibble)
__device__ __forceinline__ u32 evq_prmt(u32 a, u32 b, u32 sel) {
  u32 d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// presence bit k -> bit 4k (8 rows -> 8 nibbles)
__device__ __forceinline__ u32 evq_spread_nibbles(u32 pb) {
  u32 x = pb & 0xffu;
  x = (x | (x << 12)) & 0x000f000fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x;
}

// presence bytes (1 = not NULL) of the thread's 8 rows, 4 rows per word: nibble k of the spread bits selects byte 0 (0x00)
// or byte 1 (0x01) of a constant - one PRMT per 4 rows
__device__ __forceinline__ void evq_presence_bytes(u32 pb, u32 (&q)[EVQ_RPT / 4]) {
  const u32 x = evq_spread_nibbles(pb);
  q[0] = __byte_perm(0x00000100u, 0u, x & 0xffffu);
  if (EVQ_RPT == 8) q[EVQ_RPT / 4 - 1] = __byte_perm(0x00000100u, 0u, x >> 16);
}

// 1-byte LEB128, optional: the value with ordinal r is byte r of the tile's payload; the thread's (up to 8) present values
// are the 8 bytes at `rank`.  They are scattered to the rows that are present with ONE PRMT per 4 rows: selector nibble k =
// number of present rows below row k (exclusive prefix of the presence bits, a multiply on the nibble-spread bits); an absent
// row sets bit 3 of its nibble, which makes PRMT replicate the sign bit of the selected byte - 0, the bytes are masked to 7
// bits (a 1-byte LEB128 value is < 128) - so NULL rows read 0 (CSTableScan.cc:877-890).  `packed`: 4 rows per word.
template <int S>
__device__ __forceinline__ void evq_fast_ldn_leb1p(const EvqTile& T, const EvqScanParams& P, u32 pb, u32 rank, u32 (&v)[EVQ_RPT],
                                                   u32 (&packed)[EVQ_RPT / 4]) {
  u32 lo, hi;
  evq_lds_unaligned64(T.stage + P.streams[S].smem_off + T.desc[S].delta + rank, lo, hi);
  lo &= 0x7f7f7f7fu;
  hi &= 0x7f7f7f7fu;
  const u32 x = evq_spread_nibbles(pb);
  const u32 sel = (x * 0x11111110u) | ((x ^ 0x11111111u) << 3);
  packed[0] = evq_prmt(lo, hi, sel & 0xffffu);
  if (EVQ_RPT == 8) packed[EVQ_RPT / 4 - 1] = evq_prmt(lo, hi, sel >> 16);
#pragma unroll
  for (int j = 0; j < EVQ_RPT / 4; ++j) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[4 * j + i] = __byte_perm(packed[j], 0u, 0x4440u + i);
  }
}
template <int S>
__device__ __forceinline__ void evq_fast_ldn_leb1(const EvqTile& T, const EvqScanParams& P, u32 pb, u32 rank, u32 (&v)[EVQ_RPT]) {
  u32 packed[EVQ_RPT / 4];
  evq_fast_ldn_leb1p<S>(T, P, pb, rank, v, packed);
}

template <int S>
__device__ __forceinline__ void evq_fast_ldn_plain64(const EvqTile& T, const EvqScanParams& P, u32 pb, u32 rank, u64 (&v)[EVQ_RPT]) {
  const u64* p = (const u64*) (T.stage + P.streams[S].smem_off + T.desc[S].delta) + rank;
#pragma unroll
  for (int k = 0; k < EVQ_RPT; ++k) v[k] = ((pb >> k) & 1u) ? p[evq_pidx(pb, k)] : 0ull;
}

template <int S>
__device__ __forceinline__ void evq_fast_ldn_plain32(const EvqTile& T, const EvqScanParams& P, u32 pb, u32 rank, u32 (&v)[EVQ_RPT]) {
  const u32* p = (const u32*) (T.stage + P.streams[S].smem_off + T.desc[S].delta) + rank;
#pragma unroll
  for (int k = 0; k < EVQ_RPT; ++k) v[k] = ((pb >> k) & 1u) ? p[evq_pidx(pb, k)] : 0u;
}

template <int S>
__device__ __forceinline__ void evq_fast_ldn_bitpack(const EvqTile& T, const EvqScanParams& P, u32 pb, u32 rank, u32 (&v)[EVQ_RPT]) {
  const EvqStreamDesc& d = T.desc[S];
  const u32* words = (const u32*) (T.stage + P.streams[S].smem_off + d.delta);
  const u32 bits = P.streams[S].bits;
#pragma unroll
  for (int k = 0; k < EVQ_RPT; ++k) v[k] = ((pb >> k) & 1u) ? evq_unpack_vertical(words, d.skew + rank + evq_pidx(pb, k), bits) : 0u;
}

#if defined(EVQ_NNV) && EVQ_NNV > 0
// Optional variable-length LEB128 column of <= 4 bytes.  Only the present values are in the data stream, so a thread's
// rows start at a value ordinal that depends on the presence bits of all rows before them - but the VALUES can be decoded
// without knowing that: thread t decodes the tile's value ordinals 8t .. 8t+7 exactly like a required column (two chains
// from the column's sub-index entries) into the staging array, before the one barrier of the tile that the presence scan
// needs anyway; behind it every row gathers its value at ordinal rank + (present rows of the thread below it).
template <int S, int X, int L, int LMIN, int NV>
__device__ __forceinline__ void evq_fast_stage_vals(const EvqTile& T, const EvqScanParams& P, EvqFastScratch* scr, u32 buf) {
  if (EVQ_RPT * T.ctid >= T.desc[S].nvals) return;
  u32 raw[EVQ_RPT];
  evq_fast_ld_leb32<S, 2, L, LMIN>(T, P, evq_fast_general<S, L>(T), evq_fast_substart<X>(T, P), raw);
  const u32 a = evq_smem_u32(&scr->nval[buf][NV][EVQ_RPT * T.ctid]);
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(a), "r"(raw[0]), "r"(raw[1]), "r"(raw[2]), "r"(raw[3]) : "memory");
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(a + 16u), "r"(raw[4]), "r"(raw[5]), "r"(raw[6]), "r"(raw[7]) : "memory");
}
template <int NV>
__device__ __forceinline__ void evq_fast_gather_vals(const EvqFastScratch* scr, u32 buf, u32 pb, u32 rank, u32 (&v)[EVQ_RPT]) {
  const u32 x = evq_spread_nibbles(pb);
  const u32 excl = x * 0x11111110u;
  const u32 a = evq_smem_u32(&scr->nval[buf][NV][rank]);
#pragma unroll
  for (int k = 0; k < EVQ_RPT; ++k) {
    u32 w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(a + 4u * ((excl >> (4 * k)) & 7u)));
    v[k] = ((pb >> k) & 1u) ? w : 0u;
  }
}
#endif

// LEB128 of 2..10 bytes, optional.  X = stream of the column's sub-index (u16 per EVQ_SUB_GRAN values of the tile), or -1
// when every value has L bytes.  The thread enters at the sub-index entry in front of its first value, skips the
// values in between and then decodes one value per present row.
template <int S, int X, int L>
__device__ __forceinline__ void evq_fast_ldn_leb(const EvqTile& T, const EvqScanParams& P, u32 pb, u32 rank, u64 (&v)[EVQ_RPT]) {
  const u8* pay = T.stage + P.streams[S].smem_off + T.desc[S].delta;
  const bool uniform = X < 0 || T.desc[S].nbytes == (u32) L * T.desc[S].nvals;
  const u8* p;
  u32 skip = 0;
  if (uniform) {
    p = pay + (u32) L * rank;
  } else {
    const u32 e = rank / EVQ_SUB_GRAN < EVQ_SUB_ENTRIES ? rank / EVQ_SUB_GRAN : EVQ_SUB_ENTRIES - 1u;   // (rank == nvals == 1024: nothing to decode)
    p = pay + ((const u16*) (T.stage + P.streams[X < 0 ? 0 : X].smem_off + T.desc[X < 0 ? 0 : X].delta))[e];
    skip = rank - e * EVQ_SUB_GRAN;
  }
#pragma unroll
  for (int k = -(int) (EVQ_SUB_GRAN - 1); k < EVQ_RPT; ++k) {
    // k < 0: one of the up to EVQ_SUB_GRAN - 1 values between the entry point and the thread's first value
    const bool take = k < 0 ? (u32) (-k) <= skip : ((pb >> k) & 1u) != 0u;
    u64 val = 0;
    if (take) {
      u32 lo, hi;
      evq_lds_unaligned64(p, lo, hi);
      u32 len;
      if (uniform) {
        len = L;
      } else {
        const u32 tl = ~lo & 0x80808080u, th = ~hi & 0x80808080u;
        if (tl) len = (32u - __clz(tl & (0u - tl))) >> 3;
        else if (L > 4 && th) len = 4u + ((32u - __clz(th & (0u - th))) >> 3);
        else len = L <= 4 ? 4u : (L > 9 && (p[8] & 0x80u)) ? 10u : 9u;
      }
      if (k >= 0) {
        if (len <= 4u) {
          val = evq_leb_pack4(lo & evq_fixed_mask(len));
        } else {
          val = (u64) evq_leb_pack4(lo & 0x7f7f7f7fu) | ((u64) evq_leb_pack4(hi & evq_fixed_mask(len - 4u)) << 28);
          if (L > 8 && len > 8u) {
            val |= ((u64) (p[8] & 0x7fu)) << 56;
            if (L > 9 && len > 9u) val |= ((u64) (p[9] & 0x7fu)) << 63;
          }
        }
      }
      p += len;
    }
    if (k >= 0) v[k] = val;
  }
}

// ---- the kernel ------------------------------------------------------------------------------------------------------

struct EvqSmemHeader {
  u64 full[4];
  u64 empty[4];
  EvqStreamDesc desc[4][EVQ_KT][EVQ_NSTREAMS > 0 ? EVQ_NSTREAMS : 1];
};

#define EVQ_HDR_BYTES ((sizeof(EvqSmemHeader) + 127) & ~127)

// The generated row functions are pasted at the next line by csrc/codegen.cc.
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
  EvqFastScratch* scr = (EvqFastScratch*) (stages + (size_t) EVQ_NSTAGES * stage_bytes);
#if EVQ_TIER == 1 && EVQ_G1 > 1
  u64* sacc = (u64*) ((u8*) scr + ((sizeof(EvqFastScratch) + 127) & ~127));
#endif

  const u32 tid = threadIdx.x;
#ifdef EVQ_PARTITION
  for (u32 p = tid; p < EVQ_MAX_PARTS; p += EVQ_NTHREADS) scr->phist[p] = 0u;
#endif
  if (tid == 0) {
    for (int s = 0; s < EVQ_NSTAGES; ++s) {
      evq_mbar_init(&hdr->full[s], 1);
      evq_mbar_init(&hdr->empty[s], EVQ_NWARPS);
    }
    evq_mbar_fence_init();
  }
  __syncthreads();

  // a CTA walks groups of EVQ_KT consecutive row tiles (one pipeline stage each)
  const u32 first_group = blockIdx.x;
  const u32 group_step = gridDim.x;
  const u32 num_groups = (P.num_tiles + EVQ_KT - 1u) / EVQ_KT;

  if (tid >= EVQ_NCONS) {
    // ===================== producer warp: one TMA bulk copy per column stream and tile group =====================
    // lane i owns stream i; its descriptor lives in registers, and the row-tile index of the next group is read while
    // the warp still waits for that group's stage to be released
    const u32 lane = evq_lane();
    const bool active = lane < P.num_streams;
    EvqStream S;
    S.base = 0; S.off_index = 0; S.val_index = 0; S.nbytes = 0; S.kind = 0; S.bits = 0; S.smem_off = 0; S.smem_cap = 0;
    if (active) S = P.streams[lane];
    EvqCopyPlanK cp;
    evq_producer_plan_k(P, S, active, first_group, num_groups, cp);
    u32 it = 0;
    for (u32 group = first_group; group < num_groups; group += group_step, ++it) {
      const u32 s = it % EVQ_NSTAGES;
      const u32 round = it / EVQ_NSTAGES;
      if (round > 0) evq_mbar_wait(&hdr->empty[s], (round - 1) & 1u);
      evq_producer_commit_k(S, active, cp, stages + (size_t) s * stage_bytes, &hdr->desc[s][0][0], EVQ_NSTREAMS > 0 ? EVQ_NSTREAMS : 1,
                            &hdr->full[s]);
      evq_producer_plan_k(P, S, active, group + group_step, num_groups, cp);
    }
    return;
  }

  // ===================== consumer warps =====================
  u32 err = 0;
  u64 passed = 0;
  EvqTile T;
  T.ctid = tid;
  T.parity = 0;

#if EVQ_TIER == 1
#if EVQ_G1 > 1
  for (u32 g = 0; g < EVQ_G1; ++g) evq_state_init_slot(sacc, g, tid);
#if EVQ_NNARROW > 0
  u32 nacc[EVQ_NNARROW * EVQ_NG];
#pragma unroll
  for (int i = 0; i < EVQ_NNARROW * EVQ_NG; ++i) nacc[i] = 0u;
#endif
#else
  u64 racc[EVQ_NSTATE];
  evq_state_init_regs(racc);
#endif
#endif

#if EVQ_NNULL > 0
  u32 nullbuf = 0;
#endif
#ifdef EVQ_PARTITION
  const u32 phist_sa = evq_smem_u32(&scr->phist[0]);
#endif
  u32 it = 0;
  for (u32 group = first_group; group < num_groups; group += group_step, ++it) {
    const u32 s = it % EVQ_NSTAGES;
    evq_mbar_wait(&hdr->full[s], (it / EVQ_NSTAGES) & 1u);
    T.stage = stages + (size_t) s * stage_bytes;
    T.stage_sa = evq_smem_u32(T.stage);
#ifndef EVQ_DRYRUN   // (EVQGPU_DRYRUN=1, a measurement aid: the copy pipeline alone, the consumers release every stage untouched)
#pragma unroll 1
   for (u32 sub = 0; sub < EVQ_KT; ++sub) {
    const u32 tile = group * EVQ_KT + sub;
    if (tile >= P.num_tiles) break;
    T.desc = hdr->desc[s][sub];
    T.row0 = (u64) tile * EVQ_TILE_ROWS;
    {
      const u64 rem = P.num_rows - T.row0;
      T.rows = rem < EVQ_TILE_ROWS ? (u32) rem : EVQ_TILE_ROWS;
    }

    // rows of this thread inside the tile: all EVQ_RPT except in the table's last tile (one compare against a constant per row)
    const u32 nvalid = T.rows >= EVQ_RPT * (tid + 1u) ? (u32) EVQ_RPT : (T.rows > EVQ_RPT * tid ? T.rows - EVQ_RPT * tid : 0u);

#ifdef EVQ_FILTER_STREAM
    // the table's external row filter, ANDed with WHERE (which still runs on every row: CSTableScan.cc:826-833)
    const u32 keep = EVQ_RPT == 8 ? evq_filter_byte(T, P, tid) : (evq_filter_byte(T, P, tid >> 1) >> (4u * (tid & 1u))) & 15u;
#define EVQ_ROW_KEPT(k) (((keep >> (k)) & 1u) != 0u)
#else
#define EVQ_ROW_KEPT(k) true
#endif

    EvqFastPrep prep;
    evq_fast_prep(T, P, scr, prep);
    EvqCols cols;
    cols.ord0 = P.ord_base + (P.tile_row_base + tile) * (u64) EVQ_TILE_ROWS + EVQ_RPT * tid;   // (dead unless a first-row item reads row.ord)
#if EVQ_NNULL > 0
    evq_fast_nulls(T, P, scr, nullbuf, nvalid, cols);
    nullbuf ^= 1u;
#endif
    evq_fast_decode(T, P, scr, prep, cols);

#if EVQ_TIER == 0 || EVQ_TIER == 3
    // scan-only plans: ordered compaction (CSTableScan.cc:826-857 keeps table order).  Rows are owned in table order
    // (thread t: rows 4t..4t+3), so the output position is an exclusive scan of the per-thread pass counts.
    bool pass_k[EVQ_RPT];
    u32 mine = 0;
#pragma unroll
    for (int k = 0; k < EVQ_RPT; ++k) {
      EvqRow row;
      evq_fast_row(cols, k, row);
      pass_k[k] = ((u32) k < nvalid) && evq_where(row, err) && EVQ_ROW_KEPT(k);
      mine += pass_k[k] ? 1u : 0u;
    }
    u32 incl = mine;
    {
      const u32 lane = evq_lane();
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32) o) incl += n;
      }
      if (lane == 31) scr->scan[tid >> 5] = incl;
    }
    evq_cons_sync();
    u32 before = incl - mine, total = 0;
#pragma unroll
    for (int w = 0; w < EVQ_NWARPS; ++w) {
      const u32 c = scr->scan[w];
      if ((u32) w < (tid >> 5)) before += c;
      total += c;
    }
    evq_cons_sync();
    passed += mine;
#if EVQ_TIER == 0
    if (tid == 0) P.tile_counts[P.tile_row_base + tile] = total;
#else
    {
      u64 out_row = P.tile_out_base[P.tile_row_base + tile] + before;
#pragma unroll
      for (int k = 0; k < EVQ_RPT; ++k) {
        if (pass_k[k]) {
          EvqRow row;
          evq_fast_row(cols, k, row);
          evq_project(row, P, out_row, err);
          ++out_row;
        }
      }
    }
#endif
#elif EVQ_TIER == 2 && defined(EVQ_DENSE_GLOBAL)
    // direct-addressed group array: the slot follows from the key bounds, the aggregates are atomics at L2 (the array of a
    // few million groups stays L2-resident: no DRAM traffic besides the column streams)
#pragma unroll
    for (int k = 0; k < EVQ_RPT; ++k) {
      EvqRow row;
      evq_fast_row(cols, k, row);
      const bool pass = ((u32) k < nvalid) && evq_where(row, err) && EVQ_ROW_KEPT(k);
      if (pass) {
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
    }
#elif EVQ_TIER == 2 && defined(EVQ_PARTITION)
    // partitioned aggregation, pass 1: the rows that pass WHERE become records (the columns the keys and the aggregate
    // arguments read), appended to the partition their group's home slot falls into.  Every partition is ONE flat array.
    // Scattered 16-byte stores cost ~3x the kernel's other work (profiles/r02_c4_pass1_experiments.txt), so the tile's
    // records are gathered per partition in shared memory first:
    //   (1) a row takes the next place of its partition's bin (a shared-memory atomic on the bin's counter) and is stored
    //       there at once - no scan, no second look at the row; a bin holds twice the expected records, the rare row
    //       that finds it full goes to the partition directly (one global atomic, one scattered store)
    //   (2) every warp takes its share of the partitions: ONE global atomic per partition claims the run for the bin,
    //       then the bin is copied out as one contiguous run (8 bytes per lane: full-width coalesced stores)
#pragma unroll
    for (int k = 0; k < EVQ_RPT; ++k) {
      EvqRow row;
      evq_fast_row(cols, k, row);
      const bool pass = ((u32) k < nvalid) && evq_where(row, err) && EVQ_ROW_KEPT(k);
      if (pass) {
        ++passed;
        u64 key[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
        u32 ktag[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
        u64 fpv, slot;
        evq_keys(row, key, ktag, err);
        evq_ht_hash<EVQ_NKEYS>(P.ht, key, ktag, fpv, slot);
        const u32 part = (u32) (slot >> P.part_shift);
        u32 rank;   // (an explicit shared-memory atomic: through the generic pointer it went down the global-memory path)
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(rank) : "r"(phist_sa + 4u * part) : "memory");
        if (rank < EVQ_PART_BIN) {
          evq_row_store(row, scr->prec + ((size_t) part * EVQ_PART_BIN + rank) * EVQ_NREC);
        } else {
          const u64 pos = atomicAdd(P.part_cursor + part, 1u);
          if (pos < P.part_cap) evq_row_store(row, P.part_buf + ((u64) part * P.part_cap + pos) * EVQ_NREC);
          else err |= EVQ_ERR_PART_FULL;
        }
      }
    }
    evq_cons_sync();
    {
      constexpr u32 PER = (EVQ_MAX_PARTS + EVQ_NWARPS - 1) / EVQ_NWARPS;   // partitions per warp
      const u32 lane = evq_lane(), warp = tid >> 5;
#pragma unroll 1
      for (u32 p0 = 0; p0 < PER; p0 += 32u) {
        // lane l: count and claim of partition warp * PER + p0 + l (the atomics of up to 32 partitions are in flight together)
        const u32 mine = warp * PER + p0 + lane;
        u32 n_l = 0, base_l = 0;
        if (p0 + lane < PER && mine < EVQ_MAX_PARTS) {
          const u32 c = scr->phist[mine];
          scr->phist[mine] = 0u;
          n_l = c < EVQ_PART_BIN ? c : (u32) EVQ_PART_BIN;
          if (n_l) base_l = atomicAdd(P.part_cursor + mine, n_l);
        }
        const u32 cnt = PER - p0 < 32u ? PER - p0 : 32u;
        for (u32 l = 0; l < cnt; ++l) {
          const u32 n = __shfl_sync(0xffffffffu, n_l, l), base = __shfl_sync(0xffffffffu, base_l, l);
          if (n == 0u) continue;
          const u32 part = warp * PER + p0 + l;
          if ((u64) base + n > P.part_cap) { err |= EVQ_ERR_PART_FULL; continue; }
          const u64* src = scr->prec + (size_t) part * EVQ_PART_BIN * EVQ_NREC;
          u64* dst = P.part_buf + ((u64) part * P.part_cap + base) * EVQ_NREC;
          for (u32 t = lane; t < n * EVQ_NREC; t += 32u) dst[t] = src[t];
        }
      }
    }
    evq_cons_sync();
#elif EVQ_TIER == 2
    // hash tier: the group table lives in HBM, every probe is a DRAM round trip.  Rows are handled in quads: first the
    // home slots of all 4 rows are computed and their first probes issued, then the rows are resolved and their
    // aggregates updated with fire-and-forget atomics.
#pragma unroll
    for (int j = 0; j < EVQ_RPT / 4; ++j) {
      u64 key[4][EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
      u64 fpv[4], slot[4], w0[4], w1[4];
      bool pass[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = 4 * j + kk;
        EvqRow row;
        evq_fast_row(cols, k, row);
        pass[kk] = ((u32) k < nvalid) && evq_where(row, err) && EVQ_ROW_KEPT(k);
        fpv[kk] = slot[kk] = w0[kk] = w1[kk] = 0;
        if (pass[kk]) {
          u32 ktag[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
          evq_keys(row, key[kk], ktag, err);
          evq_ht_hash<EVQ_NKEYS>(P.ht, key[kk], ktag, fpv[kk], slot[kk]);
          evq_ht_prefetch(P.ht, slot[kk], w0[kk], w1[kk]);
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (pass[kk]) {
          ++passed;
          u64* sp = evq_ht_upsert_from<EVQ_NKEYS>(P.ht, key[kk], fpv[kk], slot[kk], w0[kk], w1[kk], (u64*) 0);
          if (!sp) {
            err |= EVQ_ERR_TABLE_FULL;
          } else {
            EvqRow row;
            evq_fast_row(cols, 4 * j + kk, row);
            evq_accumulate_global(row, sp + 1 + EVQ_NKEYS, err);
            EVQ_DISTINCT_ROW(row, (u64) sp, sp + 1 + EVQ_NKEYS);
          }
        }
      }
    }
#elif EVQ_TIER == 1 && EVQ_G1 > 1 && EVQ_NNARROW > 0
    // dense tier with byte-plane sums: rows are handled in quads; the quad's selector (one nibble per row: dense slot, or
    // >= 4 = did not pass) drives the dp4a accumulators, the remaining words (if any) take the shared-memory path per row
#pragma unroll
    for (int j = 0; j < EVQ_RPT / 4; ++j) {
      // one nibble per row: its dense slot, or EVQ_SEL_NONE = "did not pass" (a nibble no group mask answers to)
#define EVQ_SEL_NONE (EVQ_NG > 4 ? 7u : 4u)
#ifdef EVQ_SWAR_SLOTS
      u32 selector = evq_quad_slots(cols, j);   // slots of the 4 rows from the packed key bytes
#else
      u32 selector = EVQ_SEL_NONE * 0x1111u;
#endif
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = 4 * j + kk;
        EvqRow row;
        evq_fast_row(cols, k, row);
#ifdef EVQ_WHERE_PURE
        const bool pass = evq_where(row, err) & ((u32) k < nvalid) & EVQ_ROW_KEPT(k);
#else
        const bool pass = ((u32) k < nvalid) && evq_where(row, err) && EVQ_ROW_KEPT(k);
#endif
#if defined(EVQ_SWAR_SLOTS) && EVQ_NSTATE_SMEM == 0
        selector |= pass ? 0u : (EVQ_SEL_NONE << (4 * kk));   // nothing else to do per row: no branch
#elif defined(EVQ_SWAR_SLOTS)
        if (pass) evq_accumulate_smem(row, sacc, (selector >> (4 * kk)) & 3u, tid, P.dense_state, err);
        else selector |= EVQ_SEL_NONE << (4 * kk);
#else
        if (pass) {   // (rows passed are counted from the rows accumulators at the end)
          u64 key[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
          u32 ktag[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
          evq_keys(row, key, ktag, err);
          const u32 g = evq_dense_slot(key, ktag, err);
#ifdef EVQ_SLOT_ALWAYS_VALID   // the column statistics bound every key inside the slot range
          {
#else
          if (g != ~0u) {
#endif
            selector ^= (g ^ EVQ_SEL_NONE) << (4 * kk);
#if EVQ_NSTATE_SMEM > 0
            evq_accumulate_smem(row, sacc, g, tid, P.dense_state, err);
#endif
            EVQ_DISTINCT_ROW(row, (u64) g, P.dense_state + (u64) g * EVQ_NSTATE_ALL);
          }
        }
#endif
      }
      evq_accumulate_narrow(cols, j, selector, nacc);
    }
#else
#pragma unroll
    for (int k = 0; k < EVQ_RPT; ++k) {
      EvqRow row;
      evq_fast_row(cols, k, row);
      const bool pass = ((u32) k < nvalid) && evq_where(row, err) && EVQ_ROW_KEPT(k);
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
#else   // EVQ_TIER == 2
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
#endif
    }
#endif

   }
#endif
    __syncwarp();
    if (evq_lane() == 0) evq_mbar_arrive(&hdr->empty[s]);
  }

  // ===================== epilogue: merge this CTA's partial state =====================
#if EVQ_TIER == 1
#if EVQ_G1 > 1
  for (u32 g = 0; g < EVQ_G1; ++g) evq_state_flush_smem(sacc, g, tid, P.dense_state);
#if EVQ_NNARROW > 0
  passed += evq_narrow_rows(nacc);   // plane 0 is the rows counter
  evq_narrow_flush(nacc, P.dense_state);
#endif
#else
  evq_state_flush_regs(racc, P.dense_state);
#endif
#endif
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
