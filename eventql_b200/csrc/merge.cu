// merge.cu - GroupByMergeExpression (sql/statements/select/groupby.cc:528-637) across GPUs.
//
// The reference merges partial aggregates on a coordinator: every shard ships rows of (SHA-1 of the group key bytes,
// saved aggregate states) over TCP (transport/native/ops/query_partialaggr.cc:41-126) and the coordinator calls
// SFunction.vtable.merge per group (groupby.cc:577-612; count/sum: += , aggregate.cc:48-50,196-198).  Here every GPU
// is a shard and the exchange runs over NVLink with NCCL; group identity is byte equality of the evaluated key tuple
// incl. NULL tags (an exact substitute for the SHA-1 of those bytes).
//
//   dense tier (<= 64 slots, canonical slot assignment agreed on before the scan, query.cu:compute_dense_map):
//       all-gather of the [slots][nstate] state arrays (a few KB), combined in rank order by one small kernel ->
//       every rank ends with the full result; fully stream-ordered (no host synchronisation)
//   hash tier: owner(group) = bits of the key hash mod nranks; each rank packs its groups per owner, the packed
//       records cross NVLink in one grouped ncclSend/ncclRecv (all-to-all) and the owner re-inserts them into a
//       fresh open-addressing table with the same upsert + atomics as the scan kernel -> results stay distributed
#include <string.h>
#include <algorithm>
#include <vector>
#include "query.h"

#define EVQ_NCONS 256
#include "kernels/evq_prelude.cuh"

namespace evq {

struct MergeOps {
  int nstate;
  int nkeys;
  int ops[72];        // EVQ_OP_* per state word
  int carry_of[72];   // for a carry word: the sum word whose 64-bit wraps it counts (exact 128-bit sums of mean()), else -1
  int carry_at[72];   // for a sum word: its carry word, else -1
  unsigned char zero[72];   // hash merge: the word travels as 0 (count_distinct: the owner counts the union of the sets instead)
};

__device__ __forceinline__ u64 merge_identity(int op) {
  switch (op) {
    case EVQ_OP_MIN_U64: return evq_state_identity<EVQ_OP_MIN_U64>();
    case EVQ_OP_FIRST_ORD: return ~0ull;
    case EVQ_OP_MIN_I64: return evq_state_identity<EVQ_OP_MIN_I64>();
    case EVQ_OP_MAX_I64: return evq_state_identity<EVQ_OP_MAX_I64>();
    case EVQ_OP_MIN_F64: return evq_state_identity<EVQ_OP_MIN_F64>();
    case EVQ_OP_MAX_F64: return evq_state_identity<EVQ_OP_MAX_F64>();
    default: return 0ull;
  }
}

__device__ __forceinline__ u64 merge_combine(int op, u64 a, u64 b) {
  switch (op) {
    case EVQ_OP_ADD_U64: return evq_state_combine<EVQ_OP_ADD_U64>(a, b);
    case EVQ_OP_ADD_F64: return evq_state_combine<EVQ_OP_ADD_F64>(a, b);
    case EVQ_OP_MIN_U64: return evq_state_combine<EVQ_OP_MIN_U64>(a, b);
    case EVQ_OP_MAX_U64: return evq_state_combine<EVQ_OP_MAX_U64>(a, b);
    case EVQ_OP_MIN_I64: return evq_state_combine<EVQ_OP_MIN_I64>(a, b);
    case EVQ_OP_MAX_I64: return evq_state_combine<EVQ_OP_MAX_I64>(a, b);
    case EVQ_OP_MIN_F64: return evq_state_combine<EVQ_OP_MIN_F64>(a, b);
    default: return evq_state_combine<EVQ_OP_MAX_F64>(a, b);
  }
}

__device__ __forceinline__ void merge_atomic(int op, u64* addr, u64 v) {
  switch (op) {
    case EVQ_OP_ADD_U64: evq_state_atomic<EVQ_OP_ADD_U64>(addr, v); break;
    case EVQ_OP_ADD_F64: evq_state_atomic<EVQ_OP_ADD_F64>(addr, v); break;
    case EVQ_OP_MIN_U64: evq_state_atomic<EVQ_OP_MIN_U64>(addr, v); break;
    case EVQ_OP_MAX_U64: evq_state_atomic<EVQ_OP_MAX_U64>(addr, v); break;
    case EVQ_OP_MIN_I64: evq_state_atomic<EVQ_OP_MIN_I64>(addr, v); break;
    case EVQ_OP_MAX_I64: evq_state_atomic<EVQ_OP_MAX_I64>(addr, v); break;
    case EVQ_OP_MIN_F64: evq_state_atomic<EVQ_OP_MIN_F64>(addr, v); break;
    default: evq_state_atomic<EVQ_OP_MAX_F64>(addr, v); break;
  }
}

// ---- dense tier ------------------------------------------------------------------------------------------------------
// gathered: [nranks][nwords]; state[i] = combine over ranks in rank order (deterministic, also for double sums)
__global__ void k_merge_dense(u64* __restrict__ state, const u64* __restrict__ gathered, int nranks, u64 nwords, MergeOps mo) {
  const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nwords) return;
  const int w = (int) (i % mo.nstate);
  const int op = mo.ops[w];
  if (op == EVQ_OP_FIRST_ORD || op == EVQ_OP_FIRST_VAL) {
    // first-row pair: the (ordinal | tag, value) of the rank whose ordinal word is smallest (ordinals are rank-major)
    const u64 oi = op == EVQ_OP_FIRST_ORD ? i : i - 1;
    u64 best = gathered[oi], val = gathered[oi + 1];
    for (int r = 1; r < nranks; ++r) {
      const u64 o = gathered[(u64) r * nwords + oi];
      if (o < best) { best = o; val = gathered[(u64) r * nwords + oi + 1]; }
    }
    state[i] = op == EVQ_OP_FIRST_ORD ? best : val;
    return;
  }
  u64 acc = gathered[i];
  for (int r = 1; r < nranks; ++r) acc = merge_combine(op, acc, gathered[(u64) r * nwords + i]);
  if (mo.carry_of[w] >= 0) {
    // add the wraps of re-summing the partner word in the same rank order
    const u64 j = i - w + mo.carry_of[w];
    u64 lo = gathered[j];
    for (int r = 1; r < nranks; ++r) {
      const u64 n = gathered[(u64) r * nwords + j];
      lo += n;
      acc += lo < n ? 1ull : 0ull;
    }
  }
  state[i] = acc;
}

// ---- hash tier -------------------------------------------------------------------------------------------------------
// packed record: [keys nk][tag word][state nstate]   (u64 words)
__device__ __forceinline__ u32 owner_of(u64 fp, int nranks) { return (u32) ((fp >> 40) % (u64) nranks); }

__global__ void k_merge_count(EvqHashTable H, int nranks, u64* __restrict__ counts) {
  __shared__ u32 local[16];
  if (threadIdx.x < 16) local[threadIdx.x] = 0;
  __syncthreads();
  for (u64 slot = (u64) blockIdx.x * blockDim.x + threadIdx.x; slot < H.cap; slot += (u64) gridDim.x * blockDim.x) {
    const u64 fp = H.slots[slot * H.stride];
    if (fp) atomicAdd(&local[owner_of(fp, nranks)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < nranks && local[threadIdx.x]) atomicAdd(counts + threadIdx.x, (u64) local[threadIdx.x]);
}

__global__ void k_merge_pack(EvqHashTable H, int nranks, MergeOps mo, const u64* __restrict__ offsets, u64* __restrict__ cursors,
                             u64* __restrict__ out) {
  const int rec = mo.nkeys + 1 + mo.nstate;
  for (u64 slot = (u64) blockIdx.x * blockDim.x + threadIdx.x; slot < H.cap; slot += (u64) gridDim.x * blockDim.x) {
    const u64* sp = H.slots + slot * H.stride;
    const u64 fp = sp[0];
    if (!fp) continue;
    const u32 o = owner_of(fp, nranks);
    const u64 pos = offsets[o] + atomicAdd(cursors + o, 1ull);
    u64* dst = out + pos * rec;
    u64 tags = 0;
    for (int k = 0; k < mo.nkeys; ++k) {
      dst[k] = sp[1 + k];
      tags |= ((fp >> (2 + k)) & 1ull) << (8 * k);
    }
    dst[mo.nkeys] = tags;
    for (int s = 0; s < mo.nstate; ++s) dst[mo.nkeys + 1 + s] = mo.zero[s] ? 0ull : sp[1 + mo.nkeys + s];
  }
}

// count_distinct sets of the hash tier: a member (group, value) names its group by the address of the group's slot in this
// rank's table.  It travels to the GROUP's owner as [group keys][tag word][value]; there the group is looked up in the merged
// table and (merged slot, value) goes into a fresh set whose inserting threads count into the group's word.
__global__ void k_dpair_count(EvqHashTable D, int nranks, u64* __restrict__ counts) {
  __shared__ u32 local[16];
  if (threadIdx.x < 16) local[threadIdx.x] = 0;
  __syncthreads();
  for (u64 slot = (u64) blockIdx.x * blockDim.x + threadIdx.x; slot < D.cap; slot += (u64) gridDim.x * blockDim.x) {
    const u64* sp = D.slots + slot * D.stride;
    if (!sp[0]) continue;
    const u64* g = (const u64*) sp[1];
    atomicAdd(&local[owner_of(g[0], nranks)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < nranks && local[threadIdx.x]) atomicAdd(counts + threadIdx.x, (u64) local[threadIdx.x]);
}

__global__ void k_dpair_pack(EvqHashTable D, int gnk, int nranks, const u64* __restrict__ offsets, u64* __restrict__ cursors,
                             u64* __restrict__ out) {
  const int rec = gnk + 2;
  for (u64 slot = (u64) blockIdx.x * blockDim.x + threadIdx.x; slot < D.cap; slot += (u64) gridDim.x * blockDim.x) {
    const u64* sp = D.slots + slot * D.stride;
    if (!sp[0]) continue;
    const u64* g = (const u64*) sp[1];
    const u64 gfp = g[0];
    const u32 o = owner_of(gfp, nranks);
    const u64 pos = offsets[o] + atomicAdd(cursors + o, 1ull);
    u64* dst = out + pos * rec;
    u64 tags = 0;
    for (int k = 0; k < gnk; ++k) {
      dst[k] = g[1 + k];
      tags |= ((gfp >> (2 + k)) & 1ull) << (8 * k);
    }
    dst[gnk] = tags;
    dst[gnk + 1] = sp[2];
  }
}

template <int NK>
__global__ void k_dpair_insert(EvqHashTable M, EvqHashTable S, const u64* __restrict__ recs, u64 nrecs, int word, u32* __restrict__ status) {
  const int rec = NK + 2;
  for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < nrecs; i += (u64) gridDim.x * blockDim.x) {
    const u64* src = recs + i * rec;
    u64 key[NK > 0 ? NK : 1];
    u32 tag[NK > 0 ? NK : 1];
    const u64 tags = src[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      key[k] = src[k];
      tag[k] = (u32) ((tags >> (8 * k)) & 0xffu);
    }
    u64* sp = evq_ht_upsert<NK>(M, key, tag, (u64*) 0);   // (the group is there: its record came from the same rank)
    if (!sp) { atomicOr(status, EVQ_ERR_TABLE_FULL); continue; }
    u64 k2[2] = {(u64) sp, src[NK + 1]};
    const u32 t2[2] = {0u, 0u};
    if (!evq_ht_upsert<2>(S, k2, t2, sp + 1 + NK + word)) atomicOr(status, EVQ_ERR_TABLE_FULL);
  }
}

__global__ void k_merge_init(EvqHashTable H, MergeOps mo) {
  const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H.cap) return;
  u64* sp = H.slots + i * H.stride;
  sp[0] = 0ull;
  for (int s = 0; s < mo.nstate; ++s) sp[1 + mo.nkeys + s] = merge_identity(mo.ops[s]);
}

template <int NK>
__global__ void k_merge_insert(EvqHashTable H, MergeOps mo, const u64* __restrict__ recs, u64 nrecs, u64* __restrict__ counters,
                               u32* __restrict__ status) {
  const int rec = NK + 1 + mo.nstate;
  for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < nrecs; i += (u64) gridDim.x * blockDim.x) {
    const u64* src = recs + i * rec;
    u64 key[NK > 0 ? NK : 1];
    u32 tag[NK > 0 ? NK : 1];
    const u64 tags = src[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      key[k] = src[k];
      tag[k] = (u32) ((tags >> (8 * k)) & 0xffu);
    }
    u64* sp = evq_ht_upsert<NK>(H, key, tag, (u64*) 0);
    if (!sp) {
      atomicOr(status, EVQ_ERR_TABLE_FULL);
      continue;
    }
    u64* state = sp + 1 + NK;
    for (int s = 0; s < mo.nstate; ++s) {
      const u64 v = src[NK + 1 + s];
      if (mo.ops[s] == EVQ_OP_FIRST_VAL) continue;
      if (mo.ops[s] == EVQ_OP_FIRST_ORD) {
        if (v != ~0ull) evq_first_update(state + s, v, src[NK + 1 + s + 1]);
        continue;
      }
      if (mo.carry_at[s] >= 0) {
        if (v) {
          const u64 old = atomicAdd(state + s, v);
          if (old + v < v) atomicAdd(state + mo.carry_at[s], 1ull);
        }
      } else if (v != merge_identity(mo.ops[s])) {
        merge_atomic(mo.ops[s], state + s, v);
      }
    }
  }
}

static uint64_t next_pow2_(uint64_t v) {
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

static MergeOps merge_ops_of(const evqgpu_query& q) {
  MergeOps mo;
  memset(&mo, 0, sizeof(mo));
  mo.nstate = (int) q.state_ops.size();
  mo.nkeys = (int) q.group.size();
  if (mo.nstate > 72) fail(EVQGPU_ERR_UNSUPPORTED, "merge: more than 72 aggregate state words");
  for (int i = 0; i < 72; ++i) mo.carry_of[i] = mo.carry_at[i] = -1;
  for (int i = 0; i < mo.nstate; ++i) {
    mo.ops[i] = q.state_ops[i];
    mo.carry_of[i] = q.state_carry_of[i];
    if (q.state_carry_of[i] >= 0) mo.carry_at[q.state_carry_of[i]] = i;
  }
  for (int w : q.distinct_word)
    if (w >= 0 && w < 72) mo.zero[w] = 1;
  return mo;
}

static uint64_t exchange_by_owner(evqgpu_query& q, const EvqHashTable& H, const MergeOps& mo, DevBuf& sendbuf, DevBuf& recvbuf,
                                  int pairs_of_groups_with_keys = -1);

// count_distinct across ranks, dense tier (count_distinct_uint64_merge, aggregate.cc:102-108: the union of the sets).  A rank's
// set of one distinct argument is a table of (dense slot, value) pairs; the slot assignment is the same on every rank, so
// the pairs are exchanged by owner like groups, every owner inserts what it received into a fresh set and counts the NEW
// pairs per slot - its share of the union - into the slot's (zeroed) state word; the dense merge then sums the shares.
__global__ void k_distinct_zero(u64* __restrict__ state, u64 slots, int nstate, int word) {
  const u64 g = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (g < slots) state[g * nstate + word] = 0ull;
}

__global__ void k_distinct_insert(EvqHashTable M, const u64* __restrict__ recs, u64 nrecs, u64* __restrict__ state, u64 slots, int nstate,
                                  int word, u32* __restrict__ status) {
  for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < nrecs; i += (u64) gridDim.x * blockDim.x) {
    const u64* src = recs + i * 3;   // [slot][value][tags]
    u64 key[2] = {src[0], src[1]};
    const u32 tag[2] = {0u, 0u};
    if (key[0] >= slots) { atomicOr(status, EVQ_ERR_SLOT_RANGE); continue; }
    if (!evq_ht_upsert<2>(M, key, tag, state + key[0] * nstate + word)) atomicOr(status, EVQ_ERR_TABLE_FULL);
  }
}

static void merge_distinct_dense(evqgpu_query& q) {
  evqgpu_ctx* ctx = q.ctx;
  if (ctx->nranks > 16) fail(EVQGPU_ERR_UNSUPPORTED, "merge: at most 16 ranks");
  if (q.pending) finish_query(q);   // the local sets must be complete (a full set is grown and the scan re-run there)
  const int nstate = (int) q.state_ops.size();
  const uint64_t slots = q.shape.g1 > 1 ? (uint64_t) q.shape.g1 : 1;
  MergeOps mo;
  memset(&mo, 0, sizeof(mo));
  mo.nkeys = 2;
  mo.nstate = 0;
  if (q.merge_status.bytes < 16) q.merge_status.alloc(16);
  for (size_t d = 0; d < q.distinct_args.size(); ++d) {
    EvqHashTable H;
    H.slots = q.dt_slots[d].as<u64>();
    H.cap = q.dt_cap;
    H.stride = 4;
    H.nkeys = 2;
    const uint64_t recv_total = exchange_by_owner(q, H, mo, q.merge_send, q.merge_recv);
    k_distinct_zero<<<(unsigned) ((slots + 127) / 128), 128, 0, ctx->stream>>>(q.dense_base, slots, nstate, q.distinct_word[d]);
    EVQ_CUDA(cudaGetLastError());
    ctx->kernel_launches += 3;
    q.stats.kernel_launches += 3;
    if (!recv_total) continue;
    EvqHashTable M = H;
    M.cap = next_pow2_(std::max<uint64_t>(1024, recv_total * 2));
    if (q.merge_slots.bytes < M.cap * 8 * M.stride) q.merge_slots.alloc(M.cap * 8 * M.stride);
    M.slots = q.merge_slots.as<u64>();
    EVQ_CUDA(cudaMemsetAsync(M.slots, 0, M.cap * 8 * M.stride, ctx->stream));
    EVQ_CUDA(cudaMemsetAsync(q.merge_status.p, 0, 16, ctx->stream));
    const unsigned grid = (unsigned) std::min<uint64_t>((recv_total + 255) / 256, (uint64_t) ctx->sm_count * 8);
    k_distinct_insert<<<grid, 256, 0, ctx->stream>>>(M, q.merge_recv.as<u64>(), recv_total, q.dense_base, slots, nstate, q.distinct_word[d],
                                                     q.merge_status.as<u32>());
    EVQ_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
    q.stats.kernel_launches++;
    u32 mst = 0;
    EVQ_CUDA(cudaMemcpyAsync(&mst, q.merge_status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (mst) fail(EVQGPU_ERR_RUNTIME, "count_distinct merge: the union set overflowed or a slot was out of range (status %u)", mst);
  }
}

static void merge_dense(evqgpu_query& q) {
  evqgpu_ctx* ctx = q.ctx;
  if (!q.distinct_args.empty()) merge_distinct_dense(q);
  // the fused path: ONE kernel pushes this rank's few KB of state into the peers' buffers over NVLink, waits for theirs,
  // combines in rank order, emits and re-arms (query.cu launch_tail) - no NCCL call, no separate merge / emit kernels
  if (q.use_tail && ctx->p2p_ok && !getenv("EVQGPU_NO_P2P")) {
    launch_tail(q, true);
    q.pending = true;   // row count is read by the next finish
    return;
  }
  const MergeOps mo = merge_ops_of(q);
  const uint64_t slots = q.shape.g1 > 1 ? (uint64_t) q.shape.g1 : 1;
  const uint64_t nwords = slots * mo.nstate;
  if (q.merge_recv.bytes < nwords * 8 * ctx->nranks) q.merge_recv.alloc(nwords * 8 * ctx->nranks);
  comm_all_gather(ctx, q.dense_base, q.merge_recv.p, nwords * 8);
  k_merge_dense<<<(unsigned) ((nwords + 127) / 128), 128, 0, ctx->stream>>>(q.dense_base, q.merge_recv.as<u64>(),
                                                                           ctx->nranks, nwords, mo);
  EVQ_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  q.stats.kernel_launches++;
  if (q.use_tail) launch_tail(q, false);   // (the merged words are in place: emit + re-arm)
  else emit_results(q);
  q.pending = true;   // row count is read by the next finish
}

template <int NK>
static void launch_insert(evqgpu_ctx* ctx, EvqHashTable H, const MergeOps& mo, const u64* recs, u64 nrecs, u64* counters, u32* status) {
  const unsigned grid = (unsigned) std::min<uint64_t>((nrecs + 255) / 256, (uint64_t) ctx->sm_count * 8);
  k_merge_insert<NK><<<grid, 256, 0, ctx->stream>>>(H, mo, recs, nrecs, counters, status);
}

// The entries of an open-addressing table cross NVLink to their owners (owner = bits of the fingerprint mod nranks): packed
// as [keys][tag word][state words] into sendbuf, one grouped ncclSend/ncclRecv, received into recvbuf.  Returns the number
// of records this rank received.  Collective: every rank calls it.
// (pairs_of_groups_with_keys >= 0: H is a count_distinct set of the hash tier, its members travel to their GROUP's owner as
// [that many group keys][tag word][value], k_dpair_*)
static uint64_t exchange_by_owner(evqgpu_query& q, const EvqHashTable& H, const MergeOps& mo, DevBuf& sendbuf, DevBuf& recvbuf,
                                  int pairs_of_groups_with_keys) {
  evqgpu_ctx* ctx = q.ctx;
  const int n = ctx->nranks;
  const int gnk = pairs_of_groups_with_keys;
  const int rec = gnk >= 0 ? gnk + 2 : mo.nkeys + 1 + mo.nstate;
  // 1. groups per owner
  DevBuf& counts = q.merge_counts;
  if (counts.bytes < 3 * 16 * 8) counts.alloc(3 * 16 * 8);   // [0..16) counts, [16..32) offsets, [32..48) cursors
  EVQ_CUDA(cudaMemsetAsync(counts.p, 0, counts.bytes, ctx->stream));
  const unsigned grid = (unsigned) std::min<uint64_t>((H.cap + 255) / 256, (uint64_t) ctx->sm_count * 8);
  if (gnk >= 0) k_dpair_count<<<grid, 256, 0, ctx->stream>>>(H, n, counts.as<u64>());
  else k_merge_count<<<grid, 256, 0, ctx->stream>>>(H, n, counts.as<u64>());
  EVQ_CUDA(cudaGetLastError());
  std::vector<uint64_t> mine(16, 0);
  EVQ_CUDA(cudaMemcpyAsync(mine.data(), counts.p, 16 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
  // 2. everybody learns everybody's counts
  std::vector<uint64_t> all = comm_all_gather_host(ctx, mine);   // [rank][16]
  std::vector<uint64_t> send_off(n), send_bytes(n), recv_off(n), recv_bytes(n), offs(16, 0);
  uint64_t send_total = 0, recv_total = 0;
  for (int r = 0; r < n; ++r) {
    offs[r] = send_total;
    send_off[r] = send_total * rec * 8;
    send_bytes[r] = mine[r] * rec * 8;
    send_total += mine[r];
    const uint64_t from_r = all[(size_t) r * 16 + ctx->rank];
    recv_off[r] = recv_total * rec * 8;
    recv_bytes[r] = from_r * rec * 8;
    recv_total += from_r;
  }
  // 3. pack per owner
  // exchange buffers and the merged table are kept across executions (cudaMalloc of hundreds of MB per step would dominate)
  if (sendbuf.bytes < std::max<uint64_t>(send_total, 1) * rec * 8) sendbuf.alloc(std::max<uint64_t>(send_total, 1) * rec * 8 * 5 / 4);
  if (recvbuf.bytes < std::max<uint64_t>(recv_total, 1) * rec * 8) recvbuf.alloc(std::max<uint64_t>(recv_total, 1) * rec * 8 * 5 / 4);
  EVQ_CUDA(cudaMemcpyAsync(counts.as<u64>() + 16, offs.data(), 16 * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (gnk >= 0) k_dpair_pack<<<grid, 256, 0, ctx->stream>>>(H, gnk, n, counts.as<u64>() + 16, counts.as<u64>() + 32, sendbuf.as<u64>());
  else k_merge_pack<<<grid, 256, 0, ctx->stream>>>(H, n, mo, counts.as<u64>() + 16, counts.as<u64>() + 32, sendbuf.as<u64>());
  EVQ_CUDA(cudaGetLastError());
  // 4. all-to-all over NVLink
  comm_all_to_all(ctx, sendbuf.p, send_off.data(), send_bytes.data(), recvbuf.p, recv_off.data(), recv_bytes.data());
  return recv_total;
}

static void merge_hash(evqgpu_query& q) {
  evqgpu_ctx* ctx = q.ctx;
  const int n = ctx->nranks;
  if (n > 16) fail(EVQGPU_ERR_UNSUPPORTED, "merge: at most 16 ranks");
  const MergeOps mo = merge_ops_of(q);
  if (q.pending) finish_query(q);   // the local table must be complete (and large enough) before it is shipped

  EvqHashTable H = q.emit.ht;
  DevBuf& recvbuf = q.merge_recv;
  const uint64_t recv_total = exchange_by_owner(q, H, mo, q.merge_send, recvbuf);
  // 5. owner-side merge into a fresh table.  The merged table has its own status word: a full MERGE table (a probe run
  // longer than the bound) is grown and refilled from the received records here - it must never be mistaken for a full
  // scan table, whose remedy (re-running the local scan) would leave the records un-merged.
  uint64_t cap = std::max(q.merge_cap, next_pow2_(std::max<uint64_t>(1024, recv_total * 2)));
  DevBuf& slots = q.merge_slots;
  if (q.merge_status.bytes < 16) q.merge_status.alloc(16);
  EvqHashTable M = H;
  // (the scan table of a sliced plan is compact; the merge table is probed with 16-byte vector accesses: whole sectors)
  M.stride = (H.stride + 3u) & ~3u;
  for (;;) {
    if (slots.bytes < cap * 8 * M.stride) slots.alloc(cap * 8 * M.stride);
    M.slots = slots.as<u64>();
    M.cap = cap;
    k_merge_init<<<(unsigned) ((cap + 255) / 256), 256, 0, ctx->stream>>>(M, mo);
    EVQ_CUDA(cudaGetLastError());
    EVQ_CUDA(cudaMemsetAsync(q.counters.as<u64>() + 1, 0, 8, ctx->stream));
    EVQ_CUDA(cudaMemsetAsync(q.merge_status.p, 0, 16, ctx->stream));
    ctx->kernel_launches += 1;
    q.stats.kernel_launches += 1;
    if (!recv_total) break;
    u64* cnt = q.counters.as<u64>();
    u32* st = q.merge_status.as<u32>();
    const u64* recs = recvbuf.as<u64>();
    switch (mo.nkeys) {
      case 0: launch_insert<0>(ctx, M, mo, recs, recv_total, cnt, st); break;
      case 1: launch_insert<1>(ctx, M, mo, recs, recv_total, cnt, st); break;
      case 2: launch_insert<2>(ctx, M, mo, recs, recv_total, cnt, st); break;
      case 3: launch_insert<3>(ctx, M, mo, recs, recv_total, cnt, st); break;
      case 4: launch_insert<4>(ctx, M, mo, recs, recv_total, cnt, st); break;
      case 5: launch_insert<5>(ctx, M, mo, recs, recv_total, cnt, st); break;
      case 6: launch_insert<6>(ctx, M, mo, recs, recv_total, cnt, st); break;
      case 7: launch_insert<7>(ctx, M, mo, recs, recv_total, cnt, st); break;
      default: launch_insert<8>(ctx, M, mo, recs, recv_total, cnt, st); break;
    }
    EVQ_CUDA(cudaGetLastError());
    ctx->kernel_launches += 1;
    q.stats.kernel_launches += 1;
    u32 mst = 0;
    EVQ_CUDA(cudaMemcpyAsync(&mst, q.merge_status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!(mst & EVQ_ERR_TABLE_FULL)) break;
    if (cap >= (1ull << 31)) fail(EVQGPU_ERR_NOMEM, "merged group table exceeds 2^31 slots");
    cap *= 4;
  }
  q.merge_cap = cap;
  ctx->kernel_launches += 2;
  q.stats.kernel_launches += 2;
  // count_distinct: the sets' members follow their groups (the groups' distinct words travelled as 0); every owner unites
  // what it received per merged group and counts (count_distinct_uint64_merge, aggregate.cc:102-108)
  for (size_t d = 0; d < q.distinct_args.size(); ++d) {
    EvqHashTable D;
    D.slots = q.dt_slots[d].as<u64>();
    D.cap = q.dt_cap;
    D.stride = 4;
    D.nkeys = 2;
    const uint64_t np = exchange_by_owner(q, D, mo, q.merge_send, q.merge_recv, mo.nkeys);
    if (!np) continue;
    EvqHashTable S = D;
    S.cap = next_pow2_(std::max<uint64_t>(1024, np * 2));
    DevBuf set;
    set.alloc(S.cap * 8 * S.stride);
    S.slots = set.as<u64>();
    EVQ_CUDA(cudaMemsetAsync(S.slots, 0, S.cap * 8 * S.stride, ctx->stream));
    EVQ_CUDA(cudaMemsetAsync(q.merge_status.p, 0, 16, ctx->stream));
    const unsigned g2 = (unsigned) std::min<uint64_t>((np + 255) / 256, (uint64_t) ctx->sm_count * 8);
    const u64* pr = q.merge_recv.as<u64>();
    u32* st = q.merge_status.as<u32>();
    const int word = q.distinct_word[d];
    switch (mo.nkeys) {
      case 0: k_dpair_insert<0><<<g2, 256, 0, ctx->stream>>>(M, S, pr, np, word, st); break;
      case 1: k_dpair_insert<1><<<g2, 256, 0, ctx->stream>>>(M, S, pr, np, word, st); break;
      case 2: k_dpair_insert<2><<<g2, 256, 0, ctx->stream>>>(M, S, pr, np, word, st); break;
      case 3: k_dpair_insert<3><<<g2, 256, 0, ctx->stream>>>(M, S, pr, np, word, st); break;
      case 4: k_dpair_insert<4><<<g2, 256, 0, ctx->stream>>>(M, S, pr, np, word, st); break;
      case 5: k_dpair_insert<5><<<g2, 256, 0, ctx->stream>>>(M, S, pr, np, word, st); break;
      case 6: k_dpair_insert<6><<<g2, 256, 0, ctx->stream>>>(M, S, pr, np, word, st); break;
      case 7: k_dpair_insert<7><<<g2, 256, 0, ctx->stream>>>(M, S, pr, np, word, st); break;
      default: k_dpair_insert<8><<<g2, 256, 0, ctx->stream>>>(M, S, pr, np, word, st); break;
    }
    EVQ_CUDA(cudaGetLastError());
    ctx->kernel_launches += 3;
    q.stats.kernel_launches += 3;
    u32 mst = 0;
    EVQ_CUDA(cudaMemcpyAsync(&mst, q.merge_status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (mst) fail(EVQGPU_ERR_RUNTIME, "count_distinct merge: the union set overflowed (status %u)", mst);
  }
  // results are emitted from the merged table; the local table stays allocated for the next execution
  q.emit.ht = M;
  q.emit.slots = cap;
  q.emit_total_rows = std::max<uint64_t>(recv_total, 1);
  emit_results(q);
  q.pending = true;
  finish_query(q);
}

// ---- the coordinator of a cluster GROUP BY on the device (GroupByMergeExpression, groupby.cc:528-637) ---------------------
// The shards' rows were parsed into merge records on the host (wire.cc coordinator_parse_rows: the states are varuints,
// which only a sequential walk can delimit); here they are merged - the same owner-side insert kernel as the cross-GPU hash
// merge, keyed by the three words of the 20-byte SHA-1 group key - and emitted through the plan's own emit kernel.
// count_distinct at the coordinator (count_distinct_uint64_merge, aggregate.cc:102-108: the union of the shards' sets): every
// (group key, value) pair received goes into one device set; the thread that inserts a NEW pair counts it into the group's
// state word of the merged table.
__global__ void k_coord_distinct(EvqHashTable M, EvqHashTable S, const u64* __restrict__ pairs, u64 npairs, int word, u32* __restrict__ status) {
  for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += (u64) gridDim.x * blockDim.x) {
    const u64* p = pairs + i * 4;
    u64 k3[3] = {p[0], p[1], p[2]};
    const u32 t3[3] = {0u, 0u, 0u};
    u64* sp = evq_ht_upsert<3>(M, k3, t3, (u64*) 0);   // (the group exists: its row was merged before)
    if (!sp) { atomicOr(status, EVQ_ERR_TABLE_FULL); continue; }
    u64 k4[4] = {p[0], p[1], p[2], p[3]};
    const u32 t4[4] = {0u, 0u, 0u, 0u};
    if (!evq_ht_upsert<4>(S, k4, t4, sp + 1 + 3 + word)) atomicOr(status, EVQ_ERR_TABLE_FULL);
  }
}

void coordinator_finish(evqgpu_query& q) {
  evqgpu_ctx* ctx = q.ctx;
  use_device(ctx);
  if (q.state_ops.empty()) {   // (no rows were ever fed: the layout is still to be made)
    KernelShape none;
    layout_states(q, none);
  }
  const MergeOps mo = merge_ops_of(q);
  const uint64_t rec = 4 + (uint64_t) mo.nstate;
  const uint64_t n = q.coord_nrecords;
  if (!q.module) {
    float ms = 0;
    q.kernel_source = generate_coordinator_source(q);
    q.module = jit_compile(ctx, q.kernel_source, {"evq_emit"}, &ms);
    q.jit_ms_total += ms;
  }
  if (!q.status.p) q.status.alloc(16);
  if (!q.counters.p) q.counters.alloc(32);
  if (!q.out_count.p) q.out_count.alloc(8);
  EVQ_CUDA(cudaMemsetAsync(q.status.p, 0, 16, ctx->stream));
  EVQ_CUDA(cudaMemsetAsync(q.counters.p, 0, 32, ctx->stream));
  DevBuf& recs = q.merge_recv;
  if (recs.bytes < std::max<uint64_t>(n, 1) * rec * 8) recs.alloc(std::max<uint64_t>(n, 1) * rec * 8);
  if (n) EVQ_CUDA(cudaMemcpyAsync(recs.p, q.coord_records.data(), n * rec * 8, cudaMemcpyHostToDevice, ctx->stream));
  EvqHashTable M;
  M.stride = (u32) round_up(1 + 3 + (uint64_t) mo.nstate, 4);
  M.nkeys = 3;
  uint64_t cap = next_pow2_(std::max<uint64_t>(1024, n * 2));
  if (q.merge_status.bytes < 16) q.merge_status.alloc(16);
  for (;;) {
    if (q.merge_slots.bytes < cap * 8 * M.stride) q.merge_slots.alloc(cap * 8 * M.stride);
    M.slots = q.merge_slots.as<u64>();
    M.cap = cap;
    k_merge_init<<<(unsigned) ((cap + 255) / 256), 256, 0, ctx->stream>>>(M, mo);
    EVQ_CUDA(cudaGetLastError());
    EVQ_CUDA(cudaMemsetAsync(q.merge_status.p, 0, 16, ctx->stream));
    ctx->kernel_launches++;
    if (!n) break;
    launch_insert<3>(ctx, M, mo, recs.as<u64>(), n, q.counters.as<u64>(), q.merge_status.as<u32>());
    EVQ_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
    u32 mst = 0;
    EVQ_CUDA(cudaMemcpyAsync(&mst, q.merge_status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!(mst & EVQ_ERR_TABLE_FULL)) break;
    if (cap >= (1ull << 31)) fail(EVQGPU_ERR_NOMEM, "merged group table exceeds 2^31 slots");
    cap *= 4;
  }
  for (size_t d = 0; d < q.coord_pairs.size() && d < q.distinct_args.size(); ++d) {
    const uint64_t np = q.coord_pairs[d].size() / 4;
    if (!np) continue;
    DevBuf pairs, set;
    pairs.alloc(np * 32);
    EVQ_CUDA(cudaMemcpyAsync(pairs.p, q.coord_pairs[d].data(), np * 32, cudaMemcpyHostToDevice, ctx->stream));
    EvqHashTable S;
    S.stride = 8;   // fingerprint, 3 key words, value: two sectors
    S.nkeys = 4;
    S.cap = next_pow2_(std::max<uint64_t>(1024, np * 2));
    set.alloc(S.cap * 8 * S.stride);
    S.slots = set.as<u64>();
    EVQ_CUDA(cudaMemsetAsync(S.slots, 0, S.cap * 8 * S.stride, ctx->stream));
    EVQ_CUDA(cudaMemsetAsync(q.merge_status.p, 0, 16, ctx->stream));
    const unsigned grid = (unsigned) std::min<uint64_t>((np + 255) / 256, (uint64_t) ctx->sm_count * 8);
    k_coord_distinct<<<grid, 256, 0, ctx->stream>>>(M, S, pairs.as<u64>(), np, q.distinct_word[d], q.merge_status.as<u32>());
    EVQ_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
    u32 mst = 0;
    EVQ_CUDA(cudaMemcpyAsync(&mst, q.merge_status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (mst) fail(EVQGPU_ERR_RUNTIME, "count_distinct merge: the union set overflowed (status %u)", mst);
  }
  q.coord_pairs.clear();
  q.shape = KernelShape();
  q.shape.tier = 2;
  q.use_tail = false;
  memset(&q.emit, 0, sizeof(q.emit));
  q.emit.ht = M;
  q.emit.slots = cap;
  q.emit_total_rows = std::max<uint64_t>(n, 1);
  q.stats = evqgpu_query_stats();
  q.stats.strategy = 2;
  q.stats.rows_scanned = n;
  q.reordered = false;
  q.tables.clear();
  emit_results(q);
  q.pending = true;
  finish_query(q);
  q.merged = true;
  q.coord_records.clear();
  q.coord_records.shrink_to_fit();
  q.coord_nrecords = 0;
}

void merge_query(evqgpu_query& q) {
  if (!(q.flags & EVQGPU_QUERY_GROUPBY)) fail(EVQGPU_ERR_ARG, "evqgpu_query_merge: not an aggregate plan");
  if (q.tables.empty()) fail(EVQGPU_ERR_ARG, "evqgpu_query_merge: the query has not been executed");
  if (q.merged) return;
  use_device(q.ctx);
  if (q.ctx->nranks <= 1 || !q.ctx->nccl_comm) {
    if (!q.emitted) {
      if (q.use_tail) launch_tail(q, false);
      else emit_results(q);
      q.pending = true;
    }
    q.merged = true;
    return;
  }
  if (!(q.flags & EVQGPU_QUERY_PARTIAL))
    fail(EVQGPU_ERR_ARG, "evqgpu_query_merge: the plan was not created with EVQGPU_QUERY_PARTIAL");
  // (the ranks agreed on the aggregation strategy, the slot assignment and the state layout in evqgpu_query_prepare,
  // one unconditional collective; nothing here decides per rank whether to enter a collective)
  if (q.shape.tier == 1) {
    merge_dense(q);
  } else if (q.shape.dense_global) {
    // direct-addressed group array (the ranks agreed on the key box in evqgpu_query_prepare; every word is a wrapping
    // 64-bit sum): one all-reduce over NVLink, in place; every rank ends with the full result
    comm_all_reduce_sum_u64(q.ctx, q.dense_base, q.emit.slots * q.state_ops.size());
    q.emit_total_rows = q.emit.slots;   // (the merged groups are bounded by the box, not by THIS rank's rows)
    emit_results(q);
    q.pending = true;
  } else {
    merge_hash(q);
  }
  q.merged = true;
}

}  // namespace evq
