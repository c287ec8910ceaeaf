// codegen.cc - spells the query-specific part of the scan kernel as CUDA C text.
//
// What the reference does per row with an interpreter (Compiler::compile -> vm::Program, sql/runtime/compiler.cc:50-104;
// VM::evaluate, sql/runtime/vm.cc:107-157), this file does once per query: the WHERE program, the GROUP BY
// expressions and every aggregate's accumulate program become straight-line C inside the hand-written kernel of
// kernels/evq_scan_kernel.cuh, and the `get` side of the select list becomes the emit kernel.
#include <sstream>
#include "query.h"

namespace evq {

int state_words_of(const FnInfo& fi) {
  switch (fi.fn) {
    case Fn::COUNT: return 0;   // shares the per-group row counter (state word 0)
    case Fn::SUM: return 1;
    case Fn::MIN:
    case Fn::MAX:
    case Fn::MEAN: return 2;
    default: return 0;
  }
}

static int minmax_op(Fn fn, int type) {
  const bool mx = fn == Fn::MAX;
  switch (type) {
    case EVQ_INT64: return mx ? OP_MAX_I64 : OP_MIN_I64;
    case EVQ_FLOAT64: return mx ? OP_MAX_F64 : OP_MIN_F64;
    default: return mx ? OP_MAX_U64 : OP_MIN_U64;
  }
}

static const char* col_ctype(uint32_t sql_type) {
  switch (sql_type) {
    case EVQ_FLOAT64: return "f64";
    case EVQ_BOOL: return "u32";
    case EVQ_INT64: return "i64";
    default: return "u64";
  }
}

// C type a column is carried in by the fast kernel
static const char* fast_ctype(const ColSig& c) {
  if (c.sql_type == EVQ_FLOAT64) return "f64";
  if (c.sql_type == EVQ_BOOL) return "u32";
  if (c.bits <= 32) return "u32";
  return c.sql_type == EVQ_INT64 ? "i64" : "u64";
}

static CodegenEnv row_env(const KernelShape& shape) {
  CodegenEnv env;
  env.col_value.resize(shape.cols.size());
  env.col_tag.resize(shape.cols.size());
  for (size_t i = 0; i < shape.cols.size(); ++i) {
    if (!shape.cols[i].used) continue;
    // columns carried as u32 (values known to fit) are widened at every use: the query's arithmetic stays 64-bit, the
    // compiler narrows what the known-zero upper halves allow
    const bool widen = shape.fast && fast_ctype(shape.cols[i]) == std::string("u32") && shape.cols[i].sql_type != EVQ_BOOL;
    env.col_value[i] = widen ? "((u64) row.c" + std::to_string(i) + ")" : "row.c" + std::to_string(i);
    env.col_tag[i] = shape.cols[i].nullable ? "row.t" + std::to_string(i) : std::string("0u");
  }
  return env;
}

// one aggregate update: "state[st] = combine<OP>(state[st], bits)" spelled through the UPD macro of the variant
static void gen_updates(std::ostringstream& os, const evqgpu_query& q, const KernelShape& shape) {
  CodegenEnv env = row_env(shape);
  os << "  EVQ_UPD(0, " << OP_ADD_U64 << ", 1ull);\n";   // rows per group (count(...) reads this word too)
  for (const auto& item : q.select) {
    if (!item.agg) continue;
    const FnInfo& fi = item.agg->info();
    const Expr* arg = item.agg->args.empty() ? nullptr : item.agg->args[0].get();
    if (fi.fn == Fn::COUNT) {
      // count(nil): the argument is evaluated for its side effects only (aggregate.cc:35-38, conversion.cc:29-90)
      if (arg && arg->op != EVQ_X_LITERAL && !(arg->op == EVQ_X_CALL && arg->args.size() == 1 && arg->args[0]->op == EVQ_X_LITERAL)) {
        Code c = gen_expr(arg, env);
        os << "  (void) (" << c.value << ");\n";
      }
      continue;
    }
    Code c = gen_expr(arg, env);
    const int ty = fi.args[0];
    os << "  {\n";
    if (fi.fn == Fn::SUM) {
      // sum_*: acc += v, NULL contributes its value bits 0 (aggregate.cc:184-219)
      os << "    const u64 v = " << as_bits(c, ty) << ";\n";
      os << "    EVQ_UPD(" << item.state0 << ", " << (ty == EVQ_FLOAT64 ? OP_ADD_F64 : OP_ADD_U64) << ", v);\n";
    } else if (fi.fn == Fn::MIN || fi.fn == Fn::MAX) {
      os << "    if (!(" << c.tag << ")) {\n";
      os << "      const u64 v = " << as_bits(c, ty) << ";\n";
      os << "      EVQ_UPD(" << item.state0 << ", " << minmax_op(fi.fn, ty) << ", v);\n";
      os << "      EVQ_UPD(" << item.state0 + 1 << ", " << OP_ADD_U64 << ", 1ull);\n";
      os << "    }\n";
    } else {   // MEAN
      os << "    if (!(" << c.tag << ")) {\n";
      os << "      const u64 v = evq_bits((f64) (" << c.value << "));\n";
      os << "      EVQ_UPD(" << item.state0 << ", " << OP_ADD_F64 << ", v);\n";
      os << "      EVQ_UPD(" << item.state0 + 1 << ", " << OP_ADD_U64 << ", 1ull);\n";
      os << "    }\n";
    }
    os << "  }\n";
  }
}

static void gen_general_layout(std::ostringstream& os, const KernelShape& shape) {
  const size_t ncols = shape.cols.size();
  // ---- row + prep structs
  os << "struct EvqRow {\n";
  for (size_t i = 0; i < ncols; ++i)
    if (shape.cols[i].used) os << "  " << col_ctype(shape.cols[i].sql_type) << " c" << i << "; u32 t" << i << ";\n";
  os << "  u32 _unused;\n};\n";
  os << "struct EvqPrep {\n  EvqLebState leb[" << std::max(1, shape.nleb) << "];\n  u32 lebcount[" << std::max(1, shape.nleb)
     << "];\n  u32 nullpfx[" << std::max(1, shape.nnull) << "];\n};\n";

  // ---- cooperative per-tile phases
  os << "__device__ __forceinline__ void evq_prep_a(const EvqTile& T, const EvqScanParams& P, EvqScratch* scr, u32* flagword, EvqPrep& prep) {\n";
  for (size_t i = 0; i < ncols; ++i) {
    const ColSig& c = shape.cols[i];
    if (!c.used) continue;
    if (c.leb_slot >= 0)
      os << "  evq_leb_phase_a<" << c.data_stream << ", " << c.leb_slot << ">(T, P, scr, flagword, prep.leb[" << c.leb_slot
         << "], prep.lebcount[" << c.leb_slot << "]);\n";
    if (c.null_slot >= 0)
      os << "  evq_null_phase_a<" << c.level_stream << ", " << c.null_slot << ">(T, P, scr, " << c.dmax << "u);\n";
  }
  os << "}\n";
  os << "__device__ __forceinline__ void evq_prep_b(const EvqTile& T, const EvqScanParams& P, EvqScratch* scr, EvqPrep& prep, u32 flags) {\n";
  for (size_t i = 0; i < ncols; ++i) {
    const ColSig& c = shape.cols[i];
    if (c.used && c.leb_slot >= 0)
      os << "  if (flags & " << (1u << c.leb_slot) << "u) evq_leb_phase_b<" << c.data_stream << ", " << c.leb_slot
         << ">(T, P, scr, prep.leb[" << c.leb_slot << "], prep.lebcount[" << c.leb_slot << "]);\n";
  }
  os << "}\n";
  os << "__device__ __forceinline__ void evq_prep_c(const EvqTile& T, const EvqScanParams& P, EvqScratch* scr, EvqPrep& prep) {\n";
  for (size_t i = 0; i < ncols; ++i) {
    const ColSig& c = shape.cols[i];
    if (c.used && c.null_slot >= 0) os << "  prep.nullpfx[" << c.null_slot << "] = evq_null_prefix<" << c.null_slot << ">(scr);\n";
  }
  os << "}\n";

  // ---- row decode: FastCSTableScan::fetchColumn* (sql/CSTableScan.cc:860-968) for one row
  os << "__device__ __forceinline__ void evq_load_row(const EvqTile& T, const EvqScanParams& P, const EvqScratch* scr, const EvqPrep& prep, u32 r, EvqRow& row) {\n";
  for (size_t i = 0; i < ncols; ++i) {
    const ColSig& c = shape.cols[i];
    if (!c.used) continue;
    const std::string S = std::to_string(c.data_stream);
    std::string idx = "r";
    os << "  {\n";
    if (c.nullable) {
      os << "    u32 rank;\n    const bool present = evq_null_rank<" << c.null_slot << ">(scr, prep.nullpfx[" << c.null_slot
         << "], r, rank);\n";
      idx = "rank";
    }
    std::string load;
    switch (c.kind) {
      case EVQ_KIND_PLAIN64: load = "evq_ld_plain64(T, " + S + ", P.streams[" + S + "].smem_off, " + idx + ")"; break;
      case EVQ_KIND_PLAIN32: load = "evq_ld_plain32(T, " + S + ", P.streams[" + S + "].smem_off, " + idx + ")"; break;
      case EVQ_KIND_BITPACK:
        load = "evq_ld_bitpack(T, " + S + ", P.streams[" + S + "].smem_off, " + idx + ", P.streams[" + S + "].bits)";
        break;
      default:
        load = "evq_ld_leb<" + std::to_string(c.leb_slot) + ">(T, scr, prep.leb[" + std::to_string(c.leb_slot) + "], " + idx + ")";
        break;
    }
    std::string conv;
    switch (c.sql_type) {
      case EVQ_FLOAT64: conv = "evq_f64(raw)"; break;
      case EVQ_BOOL: conv = "(u32) (raw > 0)"; break;   // column_reader_uint.cc:76-90: readBoolean = value > 0
      case EVQ_INT64: conv = "(i64) raw"; break;
      default: conv = "raw"; break;
    }
    if (c.nullable) {
      // NULL: value 0, tag STAG_NULL (CSTableScan.cc:877-890)
      os << "    u64 raw = 0;\n    if (present) raw = " << load << ";\n";
      os << "    row.c" << i << " = " << conv << ";\n    row.t" << i << " = present ? 0u : 1u;\n";
    } else {
      os << "    const u64 raw = " << load << ";\n";
      os << "    row.c" << i << " = " << conv << ";\n    row.t" << i << " = 0u;\n";
    }
    os << "  }\n";
  }
  os << "}\n";

}

// fast kernel (kernels/evq_scan_fast.cuh): structs + cooperative boundary search + per-thread decode of 4 consecutive rows
static void gen_fast_layout(std::ostringstream& os, const KernelShape& shape) {
  const size_t ncols = shape.cols.size();
  os << "struct EvqRow {\n";
  for (size_t i = 0; i < ncols; ++i)
    if (shape.cols[i].used) os << "  " << fast_ctype(shape.cols[i]) << " c" << i << ";\n";
  os << "  u32 _unused;\n};\n";
  os << "struct EvqCols {\n";
  for (size_t i = 0; i < ncols; ++i)
    if (shape.cols[i].used) os << "  " << fast_ctype(shape.cols[i]) << " c" << i << "[EVQ_RPT];\n";
  os << "  u32 _unused;\n};\n";
  const int ngen = std::max(1, shape.ngen);
  os << "struct EvqFastPrep {\n  bool general[" << ngen << "];\n  u32 count[" << ngen << "];\n  u32 incl[" << ngen << "];\n};\n";

  // ---- boundary search of the variable-length columns: 2 consumer barriers per tile, only when a tile needs them
  os << "__device__ __forceinline__ void evq_fast_prep(const EvqTile& T, const EvqScanParams& P, EvqFastScratch* scr, EvqFastPrep& prep) {\n";
  if (shape.ngen > 0) {
    os << "  bool any = false;\n";
    for (size_t i = 0; i < ncols; ++i) {
      const ColSig& c = shape.cols[i];
      if (!c.used || c.gen_slot < 0) continue;
      os << "  evq_fast_count<" << c.data_stream << ", " << c.leb_len << ">(T, P, prep.general[" << c.gen_slot << "], prep.count["
         << c.gen_slot << "]);\n  any = any || prep.general[" << c.gen_slot << "];\n";
    }
    os << "  if (any) {\n";
    for (size_t i = 0; i < ncols; ++i) {
      const ColSig& c = shape.cols[i];
      if (!c.used || c.gen_slot < 0) continue;
      os << "    if (prep.general[" << c.gen_slot << "]) prep.incl[" << c.gen_slot << "] = evq_fast_publish<" << c.gen_slot
         << ">(T, scr, prep.count[" << c.gen_slot << "]);\n";
    }
    os << "    evq_cons_sync();\n";
    for (size_t i = 0; i < ncols; ++i) {
      const ColSig& c = shape.cols[i];
      if (!c.used || c.gen_slot < 0) continue;
      os << "    if (prep.general[" << c.gen_slot << "]) evq_fast_write_starts<" << c.data_stream << ", " << c.gen_slot
         << ">(T, P, scr, prep.count[" << c.gen_slot << "], prep.incl[" << c.gen_slot << "]);\n";
    }
    os << "    evq_cons_sync();\n  }\n";
  }
  os << "}\n";

  // ---- FastCSTableScan::fetchColumn* (sql/CSTableScan.cc:860-968) for the thread's 4 rows
  os << "__device__ __forceinline__ void evq_fast_decode(const EvqTile& T, const EvqScanParams& P, const EvqFastScratch* scr, const EvqFastPrep& prep, EvqCols& cols) {\n";
  for (size_t i = 0; i < ncols; ++i) {
    const ColSig& c = shape.cols[i];
    if (!c.used) continue;
    const std::string S = std::to_string(c.data_stream);
    const std::string ct = fast_ctype(c);
    const bool narrow = ct == "u32";
    const std::string raw_t = (c.kind == EVQ_KIND_PLAIN64 && !(narrow && c.sql_type != EVQ_BOOL && c.sql_type != EVQ_FLOAT64)) || (c.kind == EVQ_KIND_LEB128 && c.leb_len > 4) ? "u64" : "u32";
    os << "  {\n    " << raw_t << " raw[EVQ_RPT];\n";
    switch (c.kind) {
      case EVQ_KIND_PLAIN64:
        if (raw_t == "u32") os << "    evq_fast_ld_plain64_lo<" << S << ">(T, P, raw);\n";
        else os << "    evq_fast_ld_plain64<" << S << ">(T, P, raw);\n";
        break;
      case EVQ_KIND_PLAIN32: os << "    evq_fast_ld_plain32<" << S << ">(T, P, raw);\n"; break;
      case EVQ_KIND_BITPACK: os << "    evq_fast_ld_bitpack<" << S << ">(T, P, raw);\n"; break;
      default:
        if (c.leb_len <= 1) os << "    evq_fast_ld_leb1<" << S << ">(T, P, raw);\n";
        else if (c.leb_len <= 4)
          os << "    evq_fast_ld_leb32<" << S << ", " << c.gen_slot << ", " << c.leb_len << ">(T, P, scr, prep.general[" << c.gen_slot << "], raw);\n";
        else
          os << "    evq_fast_ld_leb64<" << S << ", " << c.gen_slot << ", " << c.leb_len << ">(T, P, scr, prep.general[" << c.gen_slot << "], raw);\n";
        break;
    }
    std::string conv;
    switch (c.sql_type) {
      case EVQ_FLOAT64: conv = "evq_f64(raw[k])"; break;
      case EVQ_BOOL: conv = "(u32) (raw[k] > 0)"; break;   // column_reader_uint.cc:76-90: readBoolean = value > 0
      default: conv = std::string("(") + ct + ") raw[k]"; break;
    }
    os << "#pragma unroll\n    for (int k = 0; k < EVQ_RPT; ++k) cols.c" << i << "[k] = " << conv << ";\n  }\n";
  }
  os << "}\n";
  os << "__device__ __forceinline__ void evq_fast_row(const EvqCols& cols, int k, EvqRow& row) {\n";
  for (size_t i = 0; i < ncols; ++i)
    if (shape.cols[i].used) os << "  row.c" << i << " = cols.c" << i << "[k];\n";
  os << "}\n";
}

static std::string gen_row_functions(const evqgpu_query& q, const KernelShape& shape) {
  std::ostringstream os;
  if (shape.fast) gen_fast_layout(os, shape);
  else gen_general_layout(os, shape);

  // ---- WHERE
  CodegenEnv env = row_env(shape);
  os << "__device__ __forceinline__ bool evq_where(const EvqRow& row, u32& err) {\n";
  if (q.where) {
    Code c = gen_expr(q.where.get(), env);
    os << "  return (" << c.value << ") != 0;\n";   // popBool drops the tag (vm.cc:259-262)
  } else {
    os << "  return true;\n";
  }
  os << "}\n";

  // ---- GROUP BY key tuple: raw value bits + tag per expression (groupby.cc:112-135)
  os << "__device__ __forceinline__ void evq_keys(const EvqRow& row, u64* key, u32* ktag, u32& err) {\n";
  for (size_t i = 0; i < q.group.size(); ++i) {
    Code c = gen_expr(q.group[i].get(), env);
    os << "  key[" << i << "] = " << as_bits(c, q.group[i]->type) << ";\n  ktag[" << i << "] = " << c.tag << ";\n";
    // a NULL key keeps value bits 0 so that equal tuples have equal bytes
    if (c.tag != "0u") os << "  if (ktag[" << i << "]) key[" << i << "] = 0ull;\n";
  }
  os << "}\n";

  const int nstate = (int) q.state_ops.size();
  if (shape.tier == 1) {
    if (shape.g1 > 1) {
      os << "#define EVQ_SIDX(g, st) ((((g) * " << nstate << ") + (st)) * EVQ_NCONS + tid)\n";
      os << "__device__ __forceinline__ void evq_state_init_slot(u64* sacc, u32 g, u32 tid) {\n";
      for (int s = 0; s < nstate; ++s) os << "  sacc[EVQ_SIDX(g, " << s << ")] = evq_state_identity<" << q.state_ops[s] << ">();\n";
      os << "}\n";
      os << "#define EVQ_UPD(st, op, v) sacc[EVQ_SIDX(g, st)] = evq_state_combine<op>(sacc[EVQ_SIDX(g, st)], (v))\n";
      os << "__device__ __forceinline__ void evq_accumulate_smem(const EvqRow& row, u64* sacc, u32 g, u32 tid, u32& err) {\n";
      gen_updates(os, q, shape);
      os << "}\n#undef EVQ_UPD\n";
      os << "__device__ __forceinline__ void evq_state_flush_smem(u64* sacc, u32 g, u32 tid, u64* dense_state) {\n";
      for (int s = 0; s < nstate; ++s) {
        os << "  {\n    u64 v = sacc[EVQ_SIDX(g, " << s << ")];\n";
        os << "#pragma unroll\n    for (int o = 16; o > 0; o >>= 1) v = evq_state_combine<" << q.state_ops[s]
           << ">(v, __shfl_xor_sync(0xffffffffu, v, o));\n";
        os << "    if (evq_lane() == 0 && v != evq_state_identity<" << q.state_ops[s] << ">()) evq_state_atomic<" << q.state_ops[s]
           << ">(dense_state + (u64) g * " << nstate << " + " << s << ", v);\n  }\n";
      }
      os << "}\n";
    } else {
      os << "__device__ __forceinline__ void evq_state_init_regs(u64* acc) {\n";
      for (int s = 0; s < nstate; ++s) os << "  acc[" << s << "] = evq_state_identity<" << q.state_ops[s] << ">();\n";
      os << "}\n";
      os << "#define EVQ_UPD(st, op, v) acc[st] = evq_state_combine<op>(acc[st], (v))\n";
      os << "__device__ __forceinline__ void evq_accumulate_regs(const EvqRow& row, u64* acc, u32& err) {\n";
      gen_updates(os, q, shape);
      os << "}\n#undef EVQ_UPD\n";
      os << "__device__ __forceinline__ void evq_state_flush_regs(u64* acc, u64* dense_state) {\n";
      for (int s = 0; s < nstate; ++s) {
        os << "  {\n    u64 v = acc[" << s << "];\n";
        os << "#pragma unroll\n    for (int o = 16; o > 0; o >>= 1) v = evq_state_combine<" << q.state_ops[s]
           << ">(v, __shfl_xor_sync(0xffffffffu, v, o));\n";
        os << "    if (evq_lane() == 0 && v != evq_state_identity<" << q.state_ops[s] << ">()) evq_state_atomic<" << q.state_ops[s]
           << ">(dense_state + " << s << ", v);\n  }\n";
      }
      os << "}\n";
    }
  } else if (shape.tier == 2) {
    os << "#define EVQ_UPD(st, op, v) evq_state_atomic<op>(state + (u64) (st) * cap + slot, (v))\n";
    os << "__device__ __forceinline__ void evq_accumulate_global(const EvqRow& row, u64* state, u64 cap, u64 slot, u32& err) {\n";
    gen_updates(os, q, shape);
    os << "}\n#undef EVQ_UPD\n";
  } else if (shape.tier == 3) {
    // scan-only projection: select list evaluated on the rows that pass, packed SVector elements in table order
    os << "__device__ __forceinline__ void evq_project(const EvqRow& row, const EvqScanParams& P, u64 out_row, u32& err) {\n";
    for (size_t i = 0; i < q.select.size(); ++i) {
      Code c = gen_expr(q.select[i].expr.get(), env);
      const int ty = q.select[i].expr->type;
      if (ty == EVQ_BOOL)
        os << "  evq_store_packed2(P.out_cols[" << i << "] + out_row * 2, " << as_bits(c, ty) << ", " << c.tag << ");\n";
      else
        os << "  evq_store_packed9(P.out_cols[" << i << "] + out_row * 9, " << as_bits(c, ty) << ", " << c.tag << ");\n";
    }
    os << "}\n";
  }
  return os.str();
}

// typed C expression of a raw 64-bit pattern
static std::string from_bits(const std::string& bits, int type) {
  switch (type) {
    case EVQ_FLOAT64: return "evq_f64(" + bits + ")";
    case EVQ_INT64: return "((i64) (" + bits + "))";
    case EVQ_BOOL: return "((u32) ((" + bits + ") != 0))";
    default: return "(" + bits + ")";
  }
}

// init + emit kernels: GroupByExpression::nextBatch (groupby.cc:187-220): per group evaluate every select item's
// `get` program and append it to the output columns in the packed SVector encoding
static std::string gen_group_kernels(const evqgpu_query& q, const KernelShape& shape) {
  std::ostringstream os;
  const int nstate = (int) q.state_ops.size();
  const int nk = (int) q.group.size();
  os << "struct EvqInitParams { u64* dense_state; EvqHashTable ht; u64 slots; };\n";
  os << "extern \"C\" __global__ void evq_init(const __grid_constant__ EvqInitParams I) {\n";
  os << "  const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;\n  if (i >= I.slots) return;\n";
  if (shape.tier == 1) {
    for (int s = 0; s < nstate; ++s)
      os << "  I.dense_state[i * " << nstate << " + " << s << "] = evq_state_identity<" << q.state_ops[s] << ">();\n";
  } else {
    os << "  I.ht.fp[i] = 0ull;\n";
    for (int s = 0; s < nstate; ++s)
      os << "  I.ht.state[(u64) " << s << " * I.ht.cap + i] = evq_state_identity<" << q.state_ops[s] << ">();\n";
  }
  os << "}\n";

  os << "struct EvqEmitParams { const u64* dense_state; EvqHashTable ht; u64 key_min[EVQ_MAX_KEYS]; u64 key_stride[EVQ_MAX_KEYS]; "
        "u64 key_null_idx[EVQ_MAX_KEYS]; u64 key_range[EVQ_MAX_KEYS]; u64 slots; u64* out_count; u64 out_capacity; "
        "u8* out_cols[EVQ_MAX_STREAMS]; };\n";
  os << "extern \"C\" __global__ void evq_emit(const __grid_constant__ EvqEmitParams E) {\n";
  os << "  const u64 slot = (u64) blockIdx.x * blockDim.x + threadIdx.x;\n  if (slot >= E.slots) return;\n";
  os << "  u64 st[" << std::max(1, nstate) << "];\n  u64 key[" << std::max(1, nk) << "];\n  u32 ktag[" << std::max(1, nk) << "];\n";
  os << "  u32 err = 0;\n";
  if (shape.tier == 1) {
    for (int s = 0; s < nstate; ++s) os << "  st[" << s << "] = E.dense_state[slot * " << nstate << " + " << s << "];\n";
    os << "  if (st[0] == 0) return;\n";   // no row reached this group: it does not exist
    for (int i = 0; i < nk; ++i) {
      os << "  {\n    const u64 idx = (slot / E.key_stride[" << i << "]) % E.key_range[" << i << "];\n";
      os << "    ktag[" << i << "] = idx == E.key_null_idx[" << i << "] ? 1u : 0u;\n";
      os << "    key[" << i << "] = ktag[" << i << "] ? 0ull : E.key_min[" << i << "] + idx;\n  }\n";
    }
  } else {
    os << "  if (E.ht.fp[slot] == 0) return;\n";
    for (int s = 0; s < nstate; ++s) os << "  st[" << s << "] = E.ht.state[(u64) " << s << " * E.ht.cap + slot];\n";
    for (int i = 0; i < nk; ++i) {
      os << "  key[" << i << "] = E.ht.keys[(u64) " << i << " * E.ht.cap + slot];\n";
      os << "  ktag[" << i << "] = E.ht.ktags[(u64) " << i << " * E.ht.cap + slot];\n";
    }
  }
  os << "  const u64 out_row = atomicAdd(E.out_count, 1ull);\n  if (out_row >= E.out_capacity) return;\n";
  // substitutions: GROUP BY expression i -> stored key; the aggregate call -> its finished value
  CodegenEnv env;
  env.col_value.assign(q.input_columns.size(), "");
  env.col_tag.assign(q.input_columns.size(), "");
  for (int i = 0; i < nk; ++i) {
    const std::string k = "key[" + std::to_string(i) + "]";
    env.subst.push_back({q.group[i]->signature(), {from_bits(k, q.group[i]->type), "ktag[" + std::to_string(i) + "]"}});
  }
  for (size_t i = 0; i < q.select.size(); ++i) {
    const SelectItem& item = q.select[i];
    CodegenEnv e2 = env;
    if (item.agg) {
      const FnInfo& fi = item.agg->info();
      const std::string s0 = "st[" + std::to_string(std::max(0, item.state0)) + "]";
      const std::string s1 = "st[" + std::to_string(item.state0 + 1) + "]";
      std::string val;
      switch (fi.fn) {
        case Fn::COUNT: val = "st[0]"; break;                                       // count_get (aggregate.cc:40-42)
        case Fn::SUM: val = from_bits(s0, fi.ret); break;                           // sum_*_get
        case Fn::MIN:
        case Fn::MAX: val = "(" + s1 + " ? " + from_bits(s0, fi.ret) + " : " + from_bits("0ull", fi.ret) + ")"; break;
        case Fn::MEAN: val = "(evq_f64(" + s0 + ") / (f64) " + s1 + ")"; break;
        default: fail(EVQGPU_ERR_UNSUPPORTED, "aggregate %s", fi.symbol.c_str());
      }
      e2.subst.push_back({item.agg->signature(), {val, "0u"}});
    }
    Code c = gen_expr(item.expr.get(), e2);
    const int ty = item.expr->type;
    if (ty == EVQ_BOOL)
      os << "  evq_store_packed2(E.out_cols[" << i << "] + out_row * 2, " << as_bits(c, ty) << ", " << c.tag << ");\n";
    else
      os << "  evq_store_packed9(E.out_cols[" << i << "] + out_row * 9, " << as_bits(c, ty) << ", " << c.tag << ");\n";
  }
  os << "  (void) err;\n}\n";
  return os.str();
}

std::string generate_source(const evqgpu_query& q, const KernelShape& shape) {
  std::ostringstream os;
  os << "// generated by eventql_b200 csrc/codegen.cc - one fused scan kernel per (plan, column layout)\n";
  os << "#define EVQ_NCONS " << shape.ncons << "\n#define EVQ_NSTAGES " << shape.nstages << "\n#define EVQ_NSTREAMS "
     << shape.nstreams << "\n#define EVQ_TIER " << shape.tier << "\n#define EVQ_G1 " << shape.g1 << "\n#define EVQ_NSTATE "
     << std::max<size_t>(1, q.state_ops.size()) << "\n#define EVQ_NKEYS " << q.group.size() << "\n#define EVQ_NLEB "
     << shape.nleb << "\n#define EVQ_NNULL " << shape.nnull << "\n#define EVQ_HAS_PREP "
     << ((shape.nleb > 0 || shape.nnull > 0) ? 1 : 0) << "\n#define EVQ_MIN_CTAS " << shape.min_ctas << "\n#define EVQ_NGEN "
     << shape.ngen << "\n";
  os << kSrcAbi << "\n" << kSrcPrelude << "\n";
  const std::string kern = shape.fast ? kSrcScanFast : kSrcScanKernel;
  const std::string marker = "//@@EVQ_GENERATED@@";
  const size_t pos = kern.find(marker);
  if (pos == std::string::npos) fail(EVQGPU_ERR_RUNTIME, "kernel text lacks the generated-code marker");
  os << kern.substr(0, pos) << "\n" << gen_row_functions(q, shape) << "\n" << kern.substr(pos + marker.size()) << "\n";
  if (shape.tier == 1 || shape.tier == 2) os << gen_group_kernels(q, shape);
  return os.str();
}

}  // namespace evq
