// debug.cc - evqgpu_debug_generate: codegen + NVRTC without a device (the "does it build" check of the JIT text).
#include <string.h>
#include "query.h"

using namespace evq;

namespace evq {
void query_intake(evqgpu_query* q, const evqgpu_query_desc* desc);
}

extern "C" int evqgpu_debug_generate(const evqgpu_query_desc* desc, const evqgpu_debug_column* columns, uint32_t tier,
                                     uint32_t dense_slots, char* src_out, uint64_t src_cap, uint64_t* src_len_out, int compile,
                                     uint64_t* cubin_bytes_out) {
  return guarded([&] {
    if (!desc || !columns) fail(EVQGPU_ERR_ARG, "evqgpu_debug_generate: null argument");
    evqgpu_query q;
    query_intake(&q, desc);
    if (q.coordinator) {   // the emit kernel of a coordinator (GroupByMergeExpression) query
      KernelShape none;
      layout_states(q, none);
      const std::string src = generate_coordinator_source(q);
      uint64_t cubin_total = 0;
      if (compile) {
        std::string log;
        cubin_total = jit_compile_to_cubin(src, &log).size();
      }
      if (src_len_out) *src_len_out = src.size();
      if (src_out && src_cap > src.size()) memcpy(src_out, src.c_str(), src.size() + 1);
      if (cubin_bytes_out) *cubin_bytes_out = cubin_total;
      return;
    }
    KernelShape s;
    s.cols.resize(q.input_columns.size());
    for (size_t i = 0; i < q.input_columns.size(); ++i) {
      if (!q.col_used[i]) continue;
      ColSig& cs = s.cols[i];
      cs.used = true;
      cs.sql_type = columns[i].sql_type;
      switch (columns[i].encoding) {
        case EVQ_ENC_UINT64_PLAIN:
        case EVQ_ENC_FLOAT_IEEE754: cs.kind = EVQ_KIND_PLAIN64; break;
        case EVQ_ENC_UINT32_PLAIN: cs.kind = EVQ_KIND_PLAIN32; break;
        case EVQ_ENC_UINT64_LEB128: cs.kind = EVQ_KIND_LEB128; break;
        default: cs.kind = EVQ_KIND_BITPACK; break;
      }
      cs.nullable = columns[i].dlevel_max > 0;
      cs.dmax = columns[i].dlevel_max;
      cs.bits = columns[i].value_bits ? columns[i].value_bits : 64;
      cs.vmax = cs.bits >= 64 ? ~0ull : (1ull << cs.bits) - 1;
      if (columns[i].value_max) { cs.vmax = columns[i].value_max; cs.vmin = columns[i].value_min; cs.vmin_present = columns[i].value_min; }
      if (cs.nullable) cs.vmin = 0;   // NULL rows read as 0
      cs.leb_len = columns[i].leb_max_len ? columns[i].leb_max_len : 10;
      cs.data_stream = s.nstreams++;
      if (cs.nullable) { cs.level_stream = s.nstreams++; cs.null_slot = s.nnull++; }
      if (cs.kind == EVQ_KIND_LEB128) cs.leb_slot = s.nleb++;
    }
    s.fast = !getenv("EVQGPU_NO_FAST_NULL");
    for (const auto& c : s.cols)
      if (c.used && c.nullable && c.dmax != 1) s.fast = false;
    if (s.nnull > 0 && getenv("EVQGPU_NO_SUBIDX")) s.fast = false;
    s.use_subidx = s.fast && !getenv("EVQGPU_NO_SUBIDX");
    for (auto& c : s.cols) {
      c.gen_slot = -1;
      if (!(s.fast && c.used && c.kind == EVQ_KIND_LEB128 && c.leb_len >= 2)) continue;
      if (s.use_subidx) c.sub_stream = s.nstreams++;
      else c.gen_slot = s.ngen++;
    }
    for (auto& c : s.cols)   // as finish_shape (query.cu): optional variable-length columns of <= 4 bytes are decoded by value ordinal
      if (s.fast && c.used && c.nullable && c.kind == EVQ_KIND_LEB128 && c.leb_len >= 2 && c.leb_len <= 4 && c.sub_stream >= 0 &&
          !getenv("EVQGPU_NO_STAGED_NULLS"))
        c.nv_slot = s.nnv++;
    layout_states(q, s);
    const bool groupby = q.flags & EVQGPU_QUERY_GROUPBY;
    std::vector<int> tiers;
    if (!groupby) tiers = {0, 3};
    else tiers = {(int) tier};
    std::string all;
    uint64_t cubin_total = 0;
    for (int t : tiers) {
      s.tier = t;
      s.part_bits = 0;
      s.slice_slots = 0;
      s.rec_cols.clear();
      if (t == 4) {   // the hash tier as partitioned aggregation (64 record partitions)
        s.tier = 2;
        s.part_bits = 6;
        s.slice_slots = 1024;   // (+ the second partitioning level and the shared-memory table slices)
        std::vector<bool> used(q.input_columns.size(), false);
        for (const auto& g : q.group) collect_columns(g.get(), used);
        for (const auto& item : q.select)
          if (item.agg) for (const auto& a : item.agg->args) collect_columns(a.get(), used);
        for (size_t i = 0; i < used.size(); ++i)
          if (used[i]) s.rec_cols.push_back((int) i);
      }
      s.g1 = 1;
      if (t == 1 && !q.group.empty()) {
        s.g1 = 2;
        while ((uint32_t) s.g1 < dense_slots) s.g1 <<= 1;
      }
      if (s.tier == 1 && s.g1 > 1) {   // a plausible dense map: every key spans [0, 1] (+ NULL for a bare optional column), last key fastest
        uint64_t stride = 1;
        for (size_t k = q.group.size(); k-- > 0;) {
          const Expr* g = q.group[k].get();
          const bool may_null = g->op == EVQ_X_INPUT && g->col < s.cols.size() && s.cols[g->col].nullable;
          s.dense.key_min[k] = 0;
          s.dense.key_range[k] = may_null ? 3 : 2;
          s.dense.key_null_idx[k] = may_null ? 2 : ~0ull;
          s.dense.key_stride[k] = stride;
          stride *= s.dense.key_range[k];
        }
        s.dense.slots = stride;
        while ((uint64_t) s.g1 < s.dense.slots) s.g1 <<= 1;
      }
      layout_states(q, s);
      layout_narrow(q, s);
      s.ncons = s.fast ? 128 : 256;   // what fit_shape picks first
      s.nstages = s.fast ? 2 : 3;
      s.kt = s.fast ? 2 : 1;
      s.min_ctas = s.fast ? (q.nnarrow * std::max(1, q.plane_groups) > 60 ? 3 : 4) : 2;   // as fit_shape caps it
      std::string src = generate_source(q, s);
      if (compile) {
        std::string log;
        std::vector<char> cubin = jit_compile_to_cubin(src, &log);
        cubin_total += cubin.size();
      }
      all += src;
    }
    if (src_len_out) *src_len_out = all.size();
    if (src_out && src_cap > all.size()) memcpy(src_out, all.c_str(), all.size() + 1);
    if (cubin_bytes_out) *cubin_bytes_out = cubin_total;
  });
}
