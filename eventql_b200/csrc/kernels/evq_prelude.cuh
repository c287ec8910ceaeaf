// evq_prelude.cuh - fixed device-side prelude of every query kernel (sm_100a).
//
// This text is compiled twice: by nvcc (static checks in build.py, -lineinfo / -Xptxas -v) and by
// NVRTC at query time, specialised by the #defines + generated row functions that
// csrc/codegen.cc puts in front of / behind it.  It must therefore not include any header.
//
// What is in here
//   * PTX wrappers: mbarrier, cp.async.bulk (TMA 1-D bulk copy global -> shared), named barriers
//   * the tile pipeline structures shared by producer warp and consumer warps
//   * cstable page decoders reading from the staged shared-memory tile:
//       plain u64/u32/f64      io/cstable/columns/page_reader_uint64.cc:50-70, _uint32.cc:50-71, _ieee754.cc:38-59
//       libsimdcomp vertical   deps/3rdparty/libsimdcomp/simdbitpacking.c:13793 (simdunpack), SURVEY A.4
//       unsigned LEB128        io/cstable/columns/page_reader_leb128.cc:50-72
//       definition levels      io/cstable/columns/column_reader_uint.cc:92-115 (value only if d == dmax)
//   * the global open-addressing group table (GroupByExpression's unordered_map,
//     sql/statements/select/groupby.cc:129-149) with atomics for the aggregate states

// ---- PTX wrappers ------------------------------------------------------------------------------------------------

__device__ __forceinline__ u32 evq_smem_u32(const void* p) {
  return (u32) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void evq_mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(evq_smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void evq_mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void evq_mbar_arrive(u64* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(evq_smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void evq_mbar_arrive_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(evq_smem_u32(bar)), "r"(bytes) : "memory");
}

// potentially-blocking wait: the hardware suspends the warp up to the time hint instead of spinning on the barrier
// (a spinning producer warp would otherwise burn issue slots the consumer warps need)
__device__ __forceinline__ void evq_mbar_wait(u64* bar, u32 parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "EVQ_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra EVQ_WAIT_DONE;\n"
      "bra EVQ_WAIT_LOOP;\n"
      "EVQ_WAIT_DONE:\n"
      "}\n" :: "r"(evq_smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}

// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void evq_bulk_g2s(void* dst_smem, const void* src_gmem, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(evq_smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(evq_smem_u32(bar)) : "memory");
}

// named barrier over the consumer warps only (the producer warp never joins it)
__device__ __forceinline__ void evq_cons_sync() {
  asm volatile("bar.sync 1, %0;" :: "n"(EVQ_NCONS) : "memory");
}

__device__ __forceinline__ u32 evq_lane() { return threadIdx.x & 31u; }

// ---- pipeline stage bookkeeping ------------------------------------------------------------------------------------

struct EvqStreamDesc {   // written by the producer lane that issued the copy, read by all consumers
  u32 delta;   // offset of the first payload byte inside the stream's stage region
  u32 nbytes;  // payload bytes of this tile
  u32 nvals;   // values of this tile in the stream
  u32 skew;    // BITPACK: index of the first value inside its first 128-block
};

struct EvqTile {         // per-consumer-thread view of the tile being processed
  const u8* stage;             // shared memory base of the stage
  u32 stage_sa;                // the same as a 32-bit shared-window address (for ld.shared with 32-bit address arithmetic)
  const EvqStreamDesc* desc;   // [EVQ_NSTREAMS]
  u32 rows;                    // rows in this tile
  u32 ctid;                    // consumer thread id 0..EVQ_NCONS-1
  u32 parity;                  // scratch double-buffer selector
  u64 row0;                    // table row of the tile's first row
};

// Issue the bulk copies of one row tile. Called by all 32 lanes of the producer warp; lane i owns stream i.
__device__ __forceinline__ void evq_producer_issue(const EvqScanParams& P, u32 tile, u8* stage, EvqStreamDesc* desc,
                                                   u64* full_bar) {
  const u32 lane = evq_lane();
  u32 bytes = 0;
  const u8* src = 0;
  u32 dst_off = 0;
  if (lane < P.num_streams) {
    const EvqStream& S = P.streams[lane];
    const u64 row0 = (u64) tile * EVQ_TILE_ROWS;
    const u64 rem = P.num_rows - row0;
    const u32 rows = rem < EVQ_TILE_ROWS ? (u32) rem : EVQ_TILE_ROWS;
    u64 start, end;
    u32 nvals = rows, skew = 0;
    if (S.kind == EVQ_KIND_LEVEL) {
      const u64 blk0 = row0 >> 7, blk1 = (row0 + rows + 127) >> 7;
      start = blk0 * 16 * S.bits;
      end = blk1 * 16 * S.bits;
    } else if (S.kind == EVQ_KIND_FILTER) {
      start = (u64) tile * (EVQ_TILE_ROWS / 8);
      end = start + (EVQ_TILE_ROWS / 8);
    } else {
      u64 v0 = row0, v1 = row0 + rows;
      if (S.val_index) {
        v0 = S.val_index[tile];
        v1 = S.val_index[tile + 1];
      }
      nvals = (u32) (v1 - v0);
      if (S.kind == EVQ_KIND_PLAIN64) {
        start = v0 * 8; end = v1 * 8;
      } else if (S.kind == EVQ_KIND_PLAIN32) {
        start = v0 * 4; end = v1 * 4;
      } else if (S.kind == EVQ_KIND_BITPACK) {
        const u64 blk0 = v0 >> 7, blk1 = (v1 + 127) >> 7;
        start = blk0 * 16 * S.bits;
        end = blk1 * 16 * S.bits;
        skew = (u32) (v0 - (blk0 << 7));
      } else {
        start = S.off_index[tile];
        end = S.off_index[tile + 1];
      }
    }
    const u64 al = start & ~15ull;
    bytes = (u32) (((end - al) + 15) & ~15ull);
    if (end == start) bytes = 0;
    if (bytes > S.smem_cap) {   // the host sized the stage from the tile index: cannot happen unless that is wrong
      atomicOr(P.status, EVQ_ERR_STAGE_OVERFLOW);
      bytes = S.smem_cap & ~15u;
    }
    src = S.base + al;
    dst_off = S.smem_off;
    EvqStreamDesc d;
    d.delta = (u32) (start - al);
    d.nbytes = (u32) (end - start);
    d.nvals = nvals;
    d.skew = skew;
    desc[lane] = d;
  }
  u32 total = bytes;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  __syncwarp();
  if (lane == 0) evq_mbar_arrive_expect_tx(full_bar, total);
  __syncwarp();
  if (bytes) evq_bulk_g2s(stage + dst_off, src, bytes, full_bar);
}

// The same in two steps, so that the index loads of the NEXT tile are in flight while the producer still waits for a free
// stage: plan (global loads of the row-tile index, no shared-memory side effects), then commit (descriptor + copy).
struct EvqCopyPlan {
  const u8* src;
  u32 bytes;
  EvqStreamDesc desc;
};

__device__ __forceinline__ void evq_producer_plan(const EvqScanParams& P, const EvqStream& S, bool active, u32 tile, EvqCopyPlan& cp) {
  cp.src = 0;
  cp.bytes = 0;
  if (!active || tile >= P.num_tiles) return;
  const u64 row0 = (u64) tile * EVQ_TILE_ROWS;
  const u64 rem = P.num_rows - row0;
  const u32 rows = rem < EVQ_TILE_ROWS ? (u32) rem : EVQ_TILE_ROWS;
  u64 start, end;
  u32 nvals = rows, skew = 0;
  if (S.kind == EVQ_KIND_LEVEL) {
    const u64 blk0 = row0 >> 7, blk1 = (row0 + rows + 127) >> 7;
    start = blk0 * 16 * S.bits;
    end = blk1 * 16 * S.bits;
  } else if (S.kind == EVQ_KIND_SUBIDX) {
    start = (u64) tile * EVQ_SUB_ENTRIES * 2;
    end = start + EVQ_SUB_ENTRIES * 2;
  } else if (S.kind == EVQ_KIND_FILTER) {
    start = (u64) tile * (EVQ_TILE_ROWS / 8);
    end = start + (EVQ_TILE_ROWS / 8);
  } else {
    u64 v0 = row0, v1 = row0 + rows;
    if (S.val_index) {
      v0 = S.val_index[tile];
      v1 = S.val_index[tile + 1];
    }
    nvals = (u32) (v1 - v0);
    if (S.kind == EVQ_KIND_PLAIN64) {
      start = v0 * 8; end = v1 * 8;
    } else if (S.kind == EVQ_KIND_PLAIN32) {
      start = v0 * 4; end = v1 * 4;
    } else if (S.kind == EVQ_KIND_BITPACK) {
      const u64 blk0 = v0 >> 7, blk1 = (v1 + 127) >> 7;
      start = blk0 * 16 * S.bits;
      end = blk1 * 16 * S.bits;
      skew = (u32) (v0 - (blk0 << 7));
    } else {
      start = S.off_index[tile];
      end = S.off_index[tile + 1];
    }
  }
  const u64 al = start & ~15ull;
  u32 bytes = (u32) (((end - al) + 15) & ~15ull);
  if (end == start) bytes = 0;
  if (bytes > S.smem_cap) {   // the host sized the stage from the tile index: cannot happen unless that is wrong
    atomicOr(P.status, EVQ_ERR_STAGE_OVERFLOW);
    bytes = S.smem_cap & ~15u;
  }
  cp.src = S.base + al;
  cp.bytes = bytes;
  cp.desc.delta = (u32) (start - al);
  cp.desc.nbytes = (u32) (end - start);
  cp.desc.nvals = nvals;
  cp.desc.skew = skew;
}

__device__ __forceinline__ void evq_producer_commit(const EvqStream& S, bool active, const EvqCopyPlan& cp, u8* stage, EvqStreamDesc* desc,
                                                    u64* full_bar) {
  const u32 lane = evq_lane();
  if (active) desc[lane] = cp.desc;
  u32 total = cp.bytes;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  __syncwarp();
  if (lane == 0) evq_mbar_arrive_expect_tx(full_bar, total);
  __syncwarp();
  if (cp.bytes) evq_bulk_g2s(stage + S.smem_off, cp.src, cp.bytes, full_bar);
}

// ---- the fast kernel's producer: EVQ_KT consecutive row tiles per pipeline stage ------------------------------------------
// The streams of required columns are contiguous across row tiles, so the byte ranges of EVQ_KT consecutive tiles are ONE
// bulk copy per stream.  Copies of a few KB use the copy engine far better than 1 KB ones (measured with the consumers
// idle: 9 copies of ~1.2 KB per tile reach 4.1 TB/s, 8 KB copies 7.4 TB/s), the consumers then walk the stage tile by tile.
#ifndef EVQ_KT
#define EVQ_KT 1
#endif

struct EvqCopyPlanK {
  const u8* src;
  u32 bytes;
  EvqStreamDesc desc[EVQ_KT];
};

// ordinal of the first VALUE of row tile `tile` (<= num_tiles: the end) in a data stream: the row for required columns,
// the number of non-NULL values before the tile for optional ones
__device__ __forceinline__ u64 evq_tile_first_value(const EvqScanParams& P, const EvqStream& S, u32 tile) {
  if (S.val_index) return S.val_index[tile];
  const u64 rows = (u64) tile * EVQ_TILE_ROWS;
  return rows < P.num_rows ? rows : P.num_rows;
}

// byte offset of value ordinal v in the stream (BITPACK: of the 128-value block that holds it; `end`: of the block behind)
__device__ __forceinline__ u64 evq_value_offset(const EvqStream& S, u32 tile, u64 v, bool end) {
  switch (S.kind) {
    case EVQ_KIND_PLAIN64: return v * 8;
    case EVQ_KIND_PLAIN32: return v * 4;
    case EVQ_KIND_BITPACK:
    case EVQ_KIND_LEVEL: return ((end ? v + 127 : v) >> 7) * 16 * S.bits;
    case EVQ_KIND_SUBIDX: return (u64) tile * EVQ_SUB_ENTRIES * 2;
    case EVQ_KIND_FILTER: return (u64) tile * (EVQ_TILE_ROWS / 8);
    default: return S.off_index[tile];
  }
}

__device__ __forceinline__ void evq_producer_plan_k(const EvqScanParams& P, const EvqStream& S, bool active, u32 group, u32 num_groups,
                                                    EvqCopyPlanK& cp) {
  cp.src = 0;
  cp.bytes = 0;
  if (!active || group >= num_groups) return;
  const u32 t0 = group * EVQ_KT;
  u64 val[EVQ_KT + 1], start[EVQ_KT + 1];
#pragma unroll
  for (int i = 0; i <= EVQ_KT; ++i) {
    const u32 t = t0 + i < P.num_tiles ? t0 + i : P.num_tiles;
    if (S.kind == EVQ_KIND_LEVEL) {   // level streams are row-indexed
      const u64 rows = (u64) t * EVQ_TILE_ROWS;
      val[i] = rows < P.num_rows ? rows : P.num_rows;
    } else {
      val[i] = evq_tile_first_value(P, S, t);
    }
    start[i] = evq_value_offset(S, t, val[i], false);
  }
  const u64 end = evq_value_offset(S, t0 + EVQ_KT < P.num_tiles ? t0 + EVQ_KT : P.num_tiles, val[EVQ_KT], true);
  const u64 al = start[0] & ~15ull;
  u32 bytes = end > start[0] ? (u32) (((end - al) + 15) & ~15ull) : 0u;
  if (bytes > S.smem_cap) {   // the host sized the stage from the tile index: cannot happen unless that is wrong
    atomicOr(P.status, EVQ_ERR_STAGE_OVERFLOW);
    bytes = S.smem_cap & ~15u;
  }
  cp.src = S.base + al;
  cp.bytes = bytes;
#pragma unroll
  for (int i = 0; i < EVQ_KT; ++i) {
    cp.desc[i].delta = (u32) (start[i] - al);
    cp.desc[i].nbytes = (u32) (start[i + 1] - start[i]);
    cp.desc[i].nvals = (u32) (val[i + 1] - val[i]);
    cp.desc[i].skew = S.kind == EVQ_KIND_BITPACK ? (u32) (val[i] & 127ull) : 0u;
  }
}

// desc: [EVQ_KT][EVQ_NSTREAMS] of the stage
__device__ __forceinline__ void evq_producer_commit_k(const EvqStream& S, bool active, const EvqCopyPlanK& cp, u8* stage, EvqStreamDesc* desc,
                                                      u32 nstreams, u64* full_bar) {
  const u32 lane = evq_lane();
  if (active) {
#pragma unroll
    for (int i = 0; i < EVQ_KT; ++i) desc[i * nstreams + lane] = cp.desc[i];
  }
  u32 total = cp.bytes;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  __syncwarp();
  if (lane == 0) evq_mbar_arrive_expect_tx(full_bar, total);
  __syncwarp();
  if (cp.bytes) evq_bulk_g2s(stage + S.smem_off, cp.src, cp.bytes, full_bar);
}

// external row filter (FastCSTableScan::setFilter, CSTableScan.cc:826-833): bit r of the tile's 128 filter bytes
#ifdef EVQ_FILTER_STREAM
__device__ __forceinline__ u32 evq_filter_byte(const EvqTile& T, const EvqScanParams& P, u32 byte_index) {
  return (u32) (T.stage + P.streams[EVQ_FILTER_STREAM].smem_off + T.desc[EVQ_FILTER_STREAM].delta)[byte_index];
}
#endif

// ---- decoders over the staged tile ------------------------------------------------------------------------------------

__device__ __forceinline__ u64 evq_ld_plain64(const EvqTile& T, u32 s, u32 off, u32 idx) {
  return *(const u64*) (T.stage + off + T.desc[s].delta + 8u * idx);
}

__device__ __forceinline__ u64 evq_ld_plain32(const EvqTile& T, u32 s, u32 off, u32 idx) {
  return (u64) * (const u32*) (T.stage + off + T.desc[s].delta + 4u * idx);
}

// libsimdcomp vertical layout: value i of a 128-block lives in lane i%4 at bit (i/4)*b of that lane's stream
__device__ __forceinline__ u32 evq_unpack_vertical(const u32* words, u32 v, u32 b) {
  const u32 blk = v >> 7, i = v & 127u;
  const u32 o = (i >> 2) * b;
  const u32* w = words + blk * 4u * b + 4u * (o >> 5) + (i & 3u);
  const u32 sh = o & 31u;
  u32 x = w[0] >> sh;
  if (sh + b > 32u) x |= w[4] << (32u - sh);
  return b >= 32u ? x : (x & ((1u << b) - 1u));
}

__device__ __forceinline__ u64 evq_ld_bitpack(const EvqTile& T, u32 s, u32 off, u32 idx, u32 bits) {
  const EvqStreamDesc& d = T.desc[s];
  return (u64) evq_unpack_vertical((const u32*) (T.stage + off + d.delta), d.skew + idx, bits);
}

// definition level of row r of the tile (LEVEL streams are row-indexed; tiles start on a 128-block boundary)
__device__ __forceinline__ u32 evq_ld_level(const EvqTile& T, u32 s, u32 off, u32 r, u32 bits) {
  return evq_unpack_vertical((const u32*) (T.stage + off + T.desc[s].delta), r, bits);
}

// 8 bytes at an arbitrary shared-memory byte offset (regions are padded, over-read is safe)
__device__ __forceinline__ void evq_lds_unaligned64(const u8* p, u32& lo, u32& hi) {
  const u32 a = evq_smem_u32(p);
  const u32* w = (const u32*) (p - (a & 3u));
  const u32 sh = (a & 3u) * 8u;
  const u32 w0 = w[0], w1 = w[1], w2 = w[2];
  lo = __funnelshift_r(w0, w1, sh);
  hi = __funnelshift_r(w1, w2, sh);
}

__device__ __forceinline__ u32 evq_leb_pack4(u32 x) {
  return (x & 0x7fu) | ((x & 0x7f00u) >> 1) | ((x & 0x7f0000u) >> 2) | ((x & 0x7f000000u) >> 3);
}

// decode one unsigned LEB128 value of `len` bytes (1..10) starting at p
__device__ __forceinline__ u64 evq_leb_decode(const u8* p, u32 len) {
  u32 lo, hi;
  evq_lds_unaligned64(p, lo, hi);
  u64 v = evq_leb_pack4(lo);
  if (len < 4u) {
    v &= (1ull << (7u * len)) - 1ull;
  } else if (len > 4u) {
    u64 h = evq_leb_pack4(hi);
    if (len < 8u) h &= (1ull << (7u * (len - 4u))) - 1ull;
    v |= h << 28;
    if (len > 8u) {
      v |= ((u64) (p[8] & 0x7fu)) << 56;
      if (len > 9u) v |= ((u64) (p[9] & 0x7fu)) << 63;
    }
  }
  return v;
}

// bit i of the result = byte i of the 16-byte chunk terminates a value (msb clear)
__device__ __forceinline__ u32 evq_term_mask16(uint4 c) {
  u32 m = 0;
  u32 w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const u32 t = ~w[k] & 0x80808080u;                 // 0x80 in every terminating byte
    const u32 nib = (t * 0x00204081u) >> 28;           // gather bits 7,15,23,31 -> 4 bits
    m |= (nib & 0xfu) << (4 * k);
  }
  return m;
}

// terminator mask of 16-byte chunk c of a stage region whose payload is bytes [delta, tb)
__device__ __forceinline__ u32 evq_leb_chunk_mask(const u8* region, u32 c, u32 delta, u32 tb) {
  const uint4 q = *(const uint4*) (region + 16u * c);
  u32 m = evq_term_mask16(q);
  const u32 pos = 16u * c;
  if (pos < delta) m &= ~((1u << (delta - pos)) - 1u);
  if (tb - pos < 16u) m &= (1u << (tb - pos)) - 1u;
  return m;
}

// 4 bytes at an arbitrary shared-memory byte offset (regions are padded, over-read is safe)
__device__ __forceinline__ u32 evq_lds_unaligned32(const u8* p) {
  const u32 a = evq_smem_u32(p);
  const u32* w = (const u32*) (p - (a & 3u));
  return __funnelshift_r(w[0], w[1], (a & 3u) * 8u);
}

// terminator bit pattern of a run of w-byte values: bit k set iff k % w == w-1
__device__ __forceinline__ u64 evq_uniform_pattern(u32 w) {
  switch (w) {
    case 1: return 0xffffffffffffffffull;
    case 2: return 0xaaaaaaaaaaaaaaaaull;
    case 3: return 0x4924924924924924ull;
    case 4: return 0x8888888888888888ull;
    case 5: return 0x0842108421084210ull;
    case 6: return 0x0820820820820820ull;
    case 7: return 0x4081020408102040ull;
    case 8: return 0x8080808080808080ull;
    case 9: return 0x4020100804020100ull;
    default: return 0x0802008020080200ull;  // 10
  }
}

// Per-tile state of one LEB128 column
struct EvqLebState {
  u32 mode;     // 0 = every value is 1 byte, >0 = every value is `mode` bytes, 0xffffffff = general (use endpos)
  u32 base;     // smem byte offset of the first payload byte
};

#define EVQ_LEB_GENERAL 0xffffffffu

// ---- group table (tier 2) ------------------------------------------------------------------------------------------

__device__ __forceinline__ u64 evq_mix64(u64 x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

// Group-table words are read and written at L2 (relaxed, gpu scope): no L1 line can go stale and no probe pays
// for an L1 invalidation.  The claimant orders "key written" before "fingerprint published" with one fence.
__device__ __forceinline__ u64 evq_ld_l2(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ u32 evq_ld_l2_u8(const u8* p) {
  u32 v;
  asm volatile("ld.relaxed.gpu.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void evq_st_l2(u64* p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void evq_st_l2_u8(u8* p, u32 v) {
  asm volatile("st.relaxed.gpu.global.u8 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// hash of a key tuple -> fingerprint word and home slot
template <int NK>
__device__ __forceinline__ void evq_ht_hash(const EvqHashTable& H, const u64* key, const u32* tag, u64& fpv, u64& slot) {
  u64 h = 0x9e3779b97f4a7c15ull;
  u64 tagbits = 0;
#pragma unroll
  for (int i = 0; i < NK; ++i) {
    h = evq_mix64(h ^ key[i] ^ ((u64) tag[i] << 57)) + 0x632be59bd9b4e019ull * (u64) (i + 1);
    tagbits |= (u64) (tag[i] & 1u) << (2 + i);
  }
  fpv = (h & ~0x3ffull) | tagbits | 1ull;
  slot = (h >> 10) & (H.cap - 1);
}

// the first two words of a slot (fingerprint, key 0) in one 16-byte load: issued for several rows before any of them is
// resolved, so that the DRAM round trips of a thread's rows overlap instead of queueing behind each other
__device__ __forceinline__ void evq_ht_prefetch(const EvqHashTable& H, u64 slot, u64& w0, u64& w1) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(H.slots + slot * H.stride) : "memory");
}

// find or claim the slot of a key tuple, starting from a prefetched first probe (w0 = fingerprint word, w1 = key 0 of the
// home slot).  Returns the slot's first word, or 0 when the table is full.
template <int NK>
__device__ __forceinline__ u64* evq_ht_upsert_from(const EvqHashTable& H, const u64* key, u64 fpv, u64 slot, u64 w0, u64 w1,
                                                  u64* claimed_counter) {
  const u64 mask = H.cap - 1;
  bool first = true;
  // a probe sequence this long means the table is (locally) full: report it, the host grows the table and runs again
  // (walking a full table to its end would take cap probes for every remaining row)
  const u64 limit = mask < 4096ull ? mask : 4096ull;
  for (u64 probes = 0; probes <= limit; ++probes) {
    u64* s = H.slots + slot * H.stride;
    u64 cur = first ? w0 : evq_ld_l2(s);
    if (cur == 0) {
      cur = atomicCAS(s, 0ull, fpv | 2ull);
      if (cur == 0) {
        if (NK == 1) {
          // one key: the final fingerprint word and the key are the slot's first 16 aligned bytes - ONE vector store
          // publishes both (a reader that saw the claiming bit spins on the fingerprint word and then finds the key of the
          // same store), so no memory fence is needed: with ~every warp inserting some group while the table fills, a
          // membar per insert stalled whole warps (4.4 stalled warps per issue in pass 2 of the partitioned aggregation)
          asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" :: "l"(s), "l"(fpv), "l"(key[0]) : "memory");
        } else {
#pragma unroll
          for (int i = 0; i < NK; ++i) evq_st_l2(s + 1 + i, key[i]);
          __threadfence();
          evq_st_l2(s, fpv);
        }
        // (count_distinct: the inserting thread counts the new member of its group's set.  The GROUP table passes no
        // counter: one atomic on ONE address per new group serialises at L2 - 10 M new groups took ~10 ms for it alone)
        if (claimed_counter) atomicAdd(claimed_counter, 1ull);
        return s;
      }
      first = false;   // somebody else claimed it meanwhile: its keys must be (re)read
    }
    if ((cur | 2ull) == (fpv | 2ull)) {
      const bool stale = !first || (cur & 2ull);
      while (cur & 2ull) cur = evq_ld_l2(s);   // the claimant is still writing the keys
      bool same = true;
#pragma unroll
      for (int i = 0; i < NK; ++i) {
        u64 k = (i == 0 && !stale) ? w1 : evq_ld_l2(s + 1 + i);
        // a prefetched key that differs under an equal fingerprint is re-read (the 16-byte load is not formally atomic)
        if (i == 0 && !stale && k != key[0]) k = evq_ld_l2(s + 1);
        same = same && k == key[i];
      }
      if (same) return s;
    }
    first = false;
    slot = (slot + 1) & mask;
  }
  return (u64*) 0;
}

// find or claim the slot of a key tuple. Returns the slot's first word, or 0 when the table is full.
template <int NK>
__device__ __forceinline__ u64* evq_ht_upsert(const EvqHashTable& H, const u64* key, const u32* tag, u64* claimed_counter) {
  u64 fpv, slot, w0, w1;
  evq_ht_hash<NK>(H, key, tag, fpv, slot);
  evq_ht_prefetch(H, slot, w0, w1);
  return evq_ht_upsert_from<NK>(H, key, fpv, slot, w0, w1, claimed_counter);
}

// unsigned bytes of a times SIGNED bytes of b (a 0xff mask byte counts as -1), added to c
__device__ __forceinline__ u32 evq_dp4a_us(u32 a, u32 b, u32 c) {
  u32 d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// ---- aggregate state updates ------------------------------------------------------------------------------------------
// state identities: sum/count 0; min = all ones (u64) / INT64_MAX / +inf; max = 0 / INT64_MIN / -inf; "seen" counters 0

__device__ __forceinline__ void evq_atomic_min_f64(u64* addr, f64 v) {
  u64 old = *addr;
  while (v < __longlong_as_double((i64) old)) {
    const u64 prev = atomicCAS(addr, old, (u64) __double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}

__device__ __forceinline__ void evq_atomic_max_f64(u64* addr, f64 v) {
  u64 old = *addr;
  while (v > __longlong_as_double((i64) old)) {
    const u64 prev = atomicCAS(addr, old, (u64) __double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}

// op codes of one state word (mirrored in csrc/codegen.cc)
#define EVQ_OP_ADD_U64 0
#define EVQ_OP_ADD_F64 1
#define EVQ_OP_MIN_U64 2
#define EVQ_OP_MAX_U64 3
#define EVQ_OP_MIN_I64 4
#define EVQ_OP_MAX_I64 5
#define EVQ_OP_MIN_F64 6
#define EVQ_OP_MAX_F64 7
// "first row wins" for a non-aggregate select item that is not a function of the GROUP BY key (groupby.cc:161-172: the
// reference evaluates and boxes it for the group's FIRST row only).  A pair of adjacent, 16-byte aligned state words:
// [row ordinal << 1 | NULL tag] (the smallest wins) and [value bits of that row], updated with ONE 128-bit CAS.
#define EVQ_OP_FIRST_ORD 8
#define EVQ_OP_FIRST_VAL 9

template <int OP>
__device__ __forceinline__ u64 evq_state_identity() {
  switch (OP) {
    case EVQ_OP_MIN_U64: return ~0ull;
    case EVQ_OP_FIRST_ORD: return ~0ull;
    case EVQ_OP_MIN_I64: return 0x7fffffffffffffffull;
    case EVQ_OP_MAX_I64: return 0x8000000000000000ull;
    case EVQ_OP_MIN_F64: return 0x7ff0000000000000ull;
    case EVQ_OP_MAX_F64: return 0xfff0000000000000ull;
    default: return 0ull;
  }
}

// combine two partial states (thread-private -> CTA -> global; also the cross-GPU merge)
template <int OP>
__device__ __forceinline__ u64 evq_state_combine(u64 a, u64 b) {
  switch (OP) {
    case EVQ_OP_ADD_U64: return a + b;
    case EVQ_OP_ADD_F64: return (u64) __double_as_longlong(__longlong_as_double((i64) a) + __longlong_as_double((i64) b));
    case EVQ_OP_MIN_U64: return a < b ? a : b;
    case EVQ_OP_MAX_U64: return a > b ? a : b;
    case EVQ_OP_MIN_I64: return (i64) a < (i64) b ? a : b;
    case EVQ_OP_MAX_I64: return (i64) a > (i64) b ? a : b;
    case EVQ_OP_MIN_F64: return __longlong_as_double((i64) a) < __longlong_as_double((i64) b) ? a : b;
    default: return __longlong_as_double((i64) a) > __longlong_as_double((i64) b) ? a : b;
  }
}

template <int OP>
__device__ __forceinline__ void evq_state_atomic(u64* addr, u64 v) {
  switch (OP) {
    case EVQ_OP_ADD_U64: atomicAdd(addr, v); break;
    case EVQ_OP_ADD_F64: atomicAdd((f64*) addr, __longlong_as_double((i64) v)); break;
    case EVQ_OP_MIN_U64: atomicMin(addr, v); break;
    case EVQ_OP_MAX_U64: atomicMax(addr, v); break;
    case EVQ_OP_MIN_I64: atomicMin((i64*) addr, (i64) v); break;
    case EVQ_OP_MAX_I64: atomicMax((i64*) addr, (i64) v); break;
    case EVQ_OP_MIN_F64: evq_atomic_min_f64(addr, __longlong_as_double((i64) v)); break;
    default: evq_atomic_max_f64(addr, __longlong_as_double((i64) v)); break;
  }
}

// the same on a state word in shared memory (partitioned aggregation: the table slice of a partition lives in shared
// memory while the partition's records are aggregated).  64-bit shared atomics are compare-and-swap loops in SASS.
__device__ __forceinline__ u64 evq_lds64(u32 sa) {
  u64 v;
  asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(sa) : "memory");
  return v;
}
__device__ __forceinline__ void evq_sts64(u32 sa, u64 v) {
  asm volatile("st.volatile.shared.u64 [%0], %1;" :: "r"(sa), "l"(v) : "memory");
}
__device__ __forceinline__ u64 evq_cas_smem(u32 sa, u64 cmp, u64 val) {
  u64 o;
  asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(o) : "r"(sa), "l"(cmp), "l"(val) : "memory");
  return o;
}
__device__ __forceinline__ u64 evq_add_ret_smem(u32 sa, u64 v) {
  u64 o;
  asm volatile("atom.shared.add.u64 %0, [%1], %2;" : "=l"(o) : "r"(sa), "l"(v) : "memory");
  return o;
}
template <int OP>
__device__ __forceinline__ void evq_state_atomic_smem(u32 sa, u64 v) {
  switch (OP) {
    case EVQ_OP_ADD_U64: {
      // two NATIVE 32-bit atomics instead of a 64-bit compare-and-swap loop (whose lanes retry one after the other): the
      // low words add up modulo 2^32, and the atomic that wraps the low word carries into the high word - every carry
      // is seen by exactly one thread, so the final 64-bit sum is exact (nobody reads the word before the CTA's barrier)
      const u32 lo = (u32) v;
      u32 hi = (u32) (v >> 32);
      u32 old;
      asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(sa), "r"(lo) : "memory");
      hi += (old + lo < old) ? 1u : 0u;
      if (hi) asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(sa + 4u), "r"(hi) : "memory");
      break;
    }
    case EVQ_OP_ADD_F64: asm volatile("red.shared.add.f64 [%0], %1;" :: "r"(sa), "d"(__longlong_as_double((i64) v)) : "memory"); break;
    // minima / maxima change rarely once a group has seen a few rows: look first, update only when the value improves
    case EVQ_OP_MIN_U64: if (v < evq_lds64(sa)) asm volatile("red.shared.min.u64 [%0], %1;" :: "r"(sa), "l"(v) : "memory"); break;
    case EVQ_OP_MAX_U64: if (v > evq_lds64(sa)) asm volatile("red.shared.max.u64 [%0], %1;" :: "r"(sa), "l"(v) : "memory"); break;
    case EVQ_OP_MIN_I64: if ((i64) v < (i64) evq_lds64(sa)) asm volatile("red.shared.min.s64 [%0], %1;" :: "r"(sa), "l"(v) : "memory"); break;
    case EVQ_OP_MAX_I64: if ((i64) v > (i64) evq_lds64(sa)) asm volatile("red.shared.max.s64 [%0], %1;" :: "r"(sa), "l"(v) : "memory"); break;
    case EVQ_OP_MIN_F64: {
      const f64 x = __longlong_as_double((i64) v);
      u64 old = evq_lds64(sa);
      while (x < __longlong_as_double((i64) old)) {
        const u64 prev = evq_cas_smem(sa, old, v);
        if (prev == old) break;
        old = prev;
      }
      break;
    }
    default: {
      const f64 x = __longlong_as_double((i64) v);
      u64 old = evq_lds64(sa);
      while (x > __longlong_as_double((i64) old)) {
        const u64 prev = evq_cas_smem(sa, old, v);
        if (prev == old) break;
        old = prev;
      }
      break;
    }
  }
}

// keep the (ordinal|tag, value) pair with the smallest first word (global memory, 16-byte aligned)
__device__ __forceinline__ void evq_first_update(u64* pair, u64 ordtag, u64 value) {
  u64 cur = *(volatile u64*) pair;
  if (ordtag >= cur) return;   // the common case after the group's first rows: one (L2-resident) load
  u64 curv = *(volatile u64*) (pair + 1);
  for (;;) {
    u64 olo, ohi;
    asm volatile("{\n\t.reg .b128 c, n, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 n, {%4, %5};\n\t"
                 "atom.global.relaxed.gpu.cas.b128 o, [%6], c, n;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(olo), "=l"(ohi) : "l"(cur), "l"(curv), "l"(ordtag), "l"(value), "l"(pair) : "memory");
    if ((olo == cur && ohi == curv) || ordtag >= olo) return;
    cur = olo;
    curv = ohi;
  }
}

// SHA-1 (FIPS 180-4) of a short message (< 120 bytes: at most two blocks): the group key of the partial-aggregation wire
// format (util/SHA1.cc is what the reference calls)
__device__ __forceinline__ u32 evq_rotl(u32 x, int n) { return (x << n) | (x >> (32 - n)); }
__device__ void evq_sha1(const u8* msg, u32 len, u8* out) {
  u32 h0 = 0x67452301u, h1 = 0xefcdab89u, h2 = 0x98badcfeu, h3 = 0x10325476u, h4 = 0xc3d2e1f0u;
  const u32 nblocks = len < 56u ? 1u : 2u;
  for (u32 blk = 0; blk < nblocks; ++blk) {
    u32 w[80];
    for (u32 i = 0; i < 16; ++i) {
      u32 word = 0;
      for (u32 b = 0; b < 4; ++b) {
        const u32 pos = blk * 64u + i * 4u + b;
        u32 byte = 0;
        if (pos < len) byte = msg[pos];
        else if (pos == len) byte = 0x80u;
        word = (word << 8) | byte;
      }
      w[i] = word;
    }
    if (blk == nblocks - 1u) { w[14] = 0u; w[15] = len * 8u; }
    for (u32 i = 16; i < 80; ++i) w[i] = evq_rotl(w[i - 3] ^ w[i - 8] ^ w[i - 14] ^ w[i - 16], 1);
    u32 a = h0, b = h1, c = h2, d = h3, e = h4;
    for (u32 i = 0; i < 80; ++i) {
      u32 f, k;
      if (i < 20) { f = (b & c) | (~b & d); k = 0x5a827999u; }
      else if (i < 40) { f = b ^ c ^ d; k = 0x6ed9eba1u; }
      else if (i < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8f1bbcdcu; }
      else { f = b ^ c ^ d; k = 0xca62c1d6u; }
      const u32 t = evq_rotl(a, 5) + f + e + k + w[i];
      e = d; d = c; c = evq_rotl(b, 30); b = a; a = t;
    }
    h0 += a; h1 += b; h2 += c; h3 += d; h4 += e;
  }
  const u32 h[5] = {h0, h1, h2, h3, h4};
  for (u32 i = 0; i < 5; ++i)
    for (u32 b = 0; b < 4; ++b) out[4 * i + b] = (u8) (h[i] >> (24 - 8 * b));
}

// ---- expression helpers (semantics of sql/expressions/math.cc, boolean.cc, conversion.cc) ------------------------------

__device__ __forceinline__ u64 evq_div_u64(u64 a, u64 b, u32& err) {
  if (b == 0) { err |= EVQ_ERR_DIV_ZERO; return 0; }
  return a / b;
}
__device__ __forceinline__ u64 evq_mod_u64(u64 a, u64 b, u32& err) {
  if (b == 0) { err |= EVQ_ERR_MOD_ZERO; return 0; }
  return a % b;
}
__device__ __forceinline__ i64 evq_div_i64(i64 a, i64 b, u32& err) {
  if (b == 0) { err |= EVQ_ERR_DIV_ZERO; return 0; }
  if (b == -1) return (i64) (0ull - (u64) a);
  return a / b;
}
__device__ __forceinline__ i64 evq_mod_i64(i64 a, i64 b, u32& err) {
  if (b == 0) { err |= EVQ_ERR_MOD_ZERO; return 0; }
  if (b == -1) return 0;
  return a % b;
}
__device__ __forceinline__ f64 evq_f64(u64 bits) { return __longlong_as_double((i64) bits); }
__device__ __forceinline__ u64 evq_bits(f64 v) { return (u64) __double_as_longlong(v); }

// write one packed SVector element (sql/svalue.cc:533-549): [8 B value][1 B tag], unaligned
__device__ __forceinline__ void evq_store_packed9(u8* dst, u64 v, u32 tag) {
#pragma unroll
  for (int i = 0; i < 8; ++i) dst[i] = (u8) (v >> (8 * i));
  dst[8] = (u8) tag;
}
__device__ __forceinline__ void evq_store_packed2(u8* dst, u64 v, u32 tag) {
  dst[0] = (u8) (v != 0);
  dst[1] = (u8) tag;
}
