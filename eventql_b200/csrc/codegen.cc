// codegen.cc - spells the query-specific part of the scan kernel as CUDA C text.
//
// What the reference does per row with an interpreter (Compiler::compile -> vm::Program, sql/runtime/compiler.cc:50-104;
// VM::evaluate, sql/runtime/vm.cc:107-157), this file does once per query: the WHERE program, the GROUP BY
// expressions and every aggregate's accumulate program become straight-line C inside the hand-written kernel of
// kernels/evq_scan_kernel.cuh, and the `get` side of the select list becomes the emit kernel.
#include <functional>
#include <sstream>
#include "query.h"

namespace evq {

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
  if (shape.fast) {   // column statistics -> narrowing hints for the expression code
    env.col_max.assign(shape.cols.size(), ~0ull);
    for (size_t i = 0; i < shape.cols.size(); ++i)
      if (shape.cols[i].used) env.col_max[i] = shape.cols[i].vmax;
    env.col_min.assign(shape.cols.size(), 0ull);
    for (size_t i = 0; i < shape.cols.size(); ++i)
      if (shape.cols[i].used && !shape.cols[i].nullable) env.col_min[i] = shape.cols[i].vmin;
  }
  return env;
}

// Can the value of `e` carry a NULL tag?  Only bare column references (possibly through `if`) do.
static bool tag_is_static_zero(const Expr* e, const KernelShape& shape) {
  switch (e->op) {
    case EVQ_X_INPUT: return !(e->col < shape.cols.size() && shape.cols[e->col].nullable);
    case EVQ_X_IF: return tag_is_static_zero(e->args[1].get(), shape) && tag_is_static_zero(e->args[2].get(), shape);
    case EVQ_X_LITERAL: return e->type != EVQ_NIL;
    default: return true;
  }
}

// Aggregate state layout of one execution.  Words are shared wherever two aggregates accumulate the same thing:
//   word 0            rows of the group: count(...) of any argument (aggregate.cc:35-38) and the "seen" counter of
//                     min / max / mean whenever the argument cannot be NULL
//   "sum:<expr>"      sum_uint64 / sum_int64 (wrapping, aggregate.cc:184-219) AND the low 64 bits of mean(<expr>) over a
//                     uint64 argument; mean adds "carry:<expr>", the number of wraps of that word, so its sum is the exact
//                     128-bit integer sum (converted to double once, in the emit kernel) instead of one I2F + DADD per row
//   "fsum:<expr>"     double sums (sum_float64, mean over int64 / float64 arguments)
//   "min:" / "max:"   extrema, "seen:<expr>" their non-NULL counters
// The layout depends on the plan and on which columns are optional; evqgpu_query_merge checks that all ranks agree.
void layout_states(evqgpu_query& q, const KernelShape& shape) {
  q.state_ops.clear();
  q.state_keys.clear();
  q.state_carry_of.clear();
  q.state_global.clear();
  q.distinct_args.clear();
  q.distinct_word.clear();
  auto word = [&](const std::string& key, int op, int carry_of = -1) -> int {
    for (size_t i = 0; i < q.state_keys.size(); ++i)
      if (q.state_keys[i] == key) return (int) i;
    q.state_keys.push_back(key);
    q.state_ops.push_back(op);
    q.state_carry_of.push_back(carry_of);
    q.state_global.push_back(carry_of >= 0);
    return (int) q.state_keys.size() - 1;
  };
  word("rows", OP_ADD_U64);
  if (q.coordinator) {
    // the coordinator of a cluster GROUP BY merges SAVED states (one per select item, in the reference's formats): every item
    // has words of its own - count / sum one word, min / max {value, seen}, mean {double sum, n} - and a non-aggregate item
    // a first-row pair (below)
    int i = 0;
    for (auto& item : q.select) {
      item.state0 = item.state_seen = item.state_carry = item.distinct = -1;
      const std::string id = std::to_string(i++);
      if (!item.agg) continue;
      const FnInfo& fi = item.agg->info();
      const int ty = fi.args.empty() ? EVQ_NIL : fi.args[0];
      switch (fi.fn) {
        case Fn::COUNT: item.state0 = word("c" + id, OP_ADD_U64); break;
        case Fn::COUNT_DISTINCT:   // the size of the union of the shards' sets, counted by merge.cu k_coord_distinct
          if (q.distinct_args.size() >= EVQ_MAX_DISTINCT)
            fail(EVQGPU_ERR_UNSUPPORTED, "more than %d count_distinct aggregates in one query", EVQ_MAX_DISTINCT);
          item.state0 = word("cd" + id, OP_ADD_U64);
          item.distinct = (int) q.distinct_args.size();
          q.distinct_args.push_back(item.agg->args.empty() ? nullptr : item.agg->args[0].get());
          q.distinct_word.push_back(item.state0);
          break;
        case Fn::SUM: item.state0 = word("s" + id, fi.ret == EVQ_FLOAT64 ? OP_ADD_F64 : OP_ADD_U64); break;
        case Fn::MIN:
        case Fn::MAX:
          item.state0 = word("m" + id, minmax_op(fi.fn, ty));
          item.state_seen = word("seen" + id, OP_ADD_U64);
          break;
        case Fn::MEAN:
          item.state0 = word("fsum" + id, OP_ADD_F64);
          item.state_seen = word("n" + id, OP_ADD_U64);
          break;
        default: fail(EVQGPU_ERR_UNSUPPORTED, "aggregate %s has no partial state format", fi.symbol.c_str());
      }
    }
  }
  for (auto& item : q.select) {
    if (q.coordinator) break;
    item.state0 = item.state_seen = item.state_carry = item.distinct = -1;
    if (!item.agg) continue;
    const FnInfo& fi = item.agg->info();
    const Expr* arg = item.agg->args.empty() ? nullptr : item.agg->args[0].get();
    const std::string sig = arg ? arg->signature() : std::string();
    const int ty = fi.args.empty() ? EVQ_NIL : fi.args[0];
    const bool never_null = arg && tag_is_static_zero(arg, shape);
    switch (fi.fn) {
      case Fn::COUNT: item.state0 = 0; break;
      case Fn::COUNT_DISTINCT: {
        // count_distinct_uint64 (aggregate.cc:80-137) keeps a std::set of the values per group.  Here one global hash set
        // of (group, value) pairs per distinct argument (EvqScanParams::dt); the thread that inserts a pair - and only that
        // one - adds 1 to the group's "cd:" word, which lives in global memory only.
        item.state0 = word("cd:" + sig, OP_ADD_U64);
        q.state_global[item.state0] = true;
        for (size_t d = 0; d < q.distinct_word.size(); ++d)
          if (q.distinct_word[d] == item.state0) item.distinct = (int) d;
        if (item.distinct < 0) {
          if (q.distinct_args.size() >= EVQ_MAX_DISTINCT)
            fail(EVQGPU_ERR_UNSUPPORTED, "more than %d count_distinct aggregates in one query", EVQ_MAX_DISTINCT);
          item.distinct = (int) q.distinct_args.size();
          q.distinct_args.push_back(arg);
          q.distinct_word.push_back(item.state0);
        }
        break;
      }
      case Fn::SUM:
        item.state0 = ty == EVQ_FLOAT64 ? word("fsum:" + sig, OP_ADD_F64) : word("sum:" + sig, OP_ADD_U64);
        break;
      case Fn::MIN:
      case Fn::MAX:
        item.state0 = word(std::string(fi.fn == Fn::MAX ? "max:" : "min:") + sig, minmax_op(fi.fn, ty));
        item.state_seen = never_null ? 0 : word("seen:" + sig, OP_ADD_U64);
        break;
      case Fn::MEAN:
        if (ty == EVQ_UINT64) {
          item.state0 = word("sum:" + sig, OP_ADD_U64);
          item.state_carry = word("carry:" + sig, OP_ADD_U64, item.state0);
        } else {
          item.state0 = word((ty == EVQ_FLOAT64 ? "fsum:" : "fsumi:") + sig, OP_ADD_F64);
        }
        item.state_seen = never_null ? 0 : word("seen:" + sig, OP_ADD_U64);
        break;
      default: fail(EVQGPU_ERR_UNSUPPORTED, "aggregate %s", fi.symbol.c_str());
    }
  }
  // first-row items: a pair of adjacent words [ordinal | tag][value], global memory only, 16-byte aligned in both state
  // layouts: word w of a hash slot sits at slot + 1 + nkeys + w (slots are 32-byte aligned), word w of a dense slot at
  // base + g * nstate + w with an even nstate and a base offset by (1 + nkeys) & 1 words (query.cu) - so (1 + nkeys + w) even
  if (q.has_first) {
    const size_t nk = q.group.size();
    int npad = 0;
    auto pad = [&]() {
      word("pad:" + std::to_string(npad++), OP_ADD_U64);
      q.state_global.back() = true;
    };
    for (auto& item : q.select) {
      item.state_first = -1;
      if (!item.first) continue;
      const std::string sig = item.expr->signature();
      bool found = false;
      for (size_t i = 0; i < q.state_keys.size(); ++i)
        if (q.state_keys[i] == "first_ord:" + sig) { item.state_first = (int) i; found = true; }
      if (found) continue;
      if ((1 + nk + q.state_keys.size()) & 1) pad();
      item.state_first = word("first_ord:" + sig, OP_FIRST_ORD);
      q.state_global.back() = true;
      word("first_val:" + sig, OP_FIRST_VAL);
      q.state_global.back() = true;
    }
    if (q.state_keys.size() & 1) pad();
  }
  // carry words are touched once in 2^64 / value rows: they live in the global state only, never in thread-private storage
  q.state_smem.assign(q.state_ops.size(), -1);
  q.nstate_smem = 0;
  for (size_t i = 0; i < q.state_ops.size(); ++i)
    if (!q.state_global[i]) q.state_smem[i] = q.nstate_smem++;
  q.state_narrow.assign(q.state_ops.size(), -1);
  q.narrow_col.clear();
  q.nnarrow = 0;
}

// Byte-plane sums of the fast dense kernel with a handful of groups (<= 4).  A sum(W * B) - W an expression below 2^32,
// B one below 256 or absent - is kept as u32 registers per (byte plane of W, group) and fed 4 rows at a time with dp4a:
// the 4 rows' bytes of plane p against the 4 rows' B bytes masked by "row belongs to group g".  That replaces a 64-bit
// shared-memory read-modify-write per row and word (the fast kernel's shared-memory bandwidth limit) by about
// planes / 4 dot products per row, and the value is  SUM_p 256^p * acc[p][g].  The rows counter is W = B = 1; sums of
// 1-byte LEB128 columns are one plane: the column's raw bytes.  Only the thread-private storage changes, not the global
// state layout.
static bool is_packed_byte_column(const Expr* e, const KernelShape& shape) {
  if (e->op != EVQ_X_INPUT || e->col >= shape.cols.size()) return false;
  const ColSig& c = shape.cols[e->col];
  // (optional columns too, in the fast layout: their packed bytes hold 0 where the row is NULL, which is what every
  // consumer of the value sees)
  return c.used && c.kind == EVQ_KIND_LEB128 && c.leb_len == 1 && (!c.nullable || c.dmax == 1) && c.sql_type != EVQ_BOOL &&
         c.sql_type != EVQ_FLOAT64 && c.sql_type != EVQ_INT64;
}

void layout_narrow(evqgpu_query& q, const KernelShape& shape) {
  q.state_narrow.assign(q.state_ops.size(), -1);
  q.narrow_col.clear();
  q.nnarrow = 0;
  q.plane_w.clear();
  q.plane_b.clear();
  q.plane_sums.clear();
  q.swar_slots = false;
  q.plane_sig.clear();
  q.plane_groups = 0;
  if (!shape.fast || shape.tier != 1 || shape.g1 < 2 || shape.dense.slots > 7 || getenv("EVQGPU_NO_NARROW")) return;
  if (q.has_first) return;   // first-row items are updated per row (evq_first_update): the per-row accumulate call stays
  // up to 4 slots: a power of two of groups (the 0xff << 8g masks of one register); 5..7 slots (a NULL-able key: flag x
  // status with a NULL flag = 6): exactly that many, the masks come from a register pair
  int ng = shape.g1;
  if (shape.g1 > 4) {
    if (shape.g1 != 8 || shape.dense.slots < 5 || getenv("EVQGPU_NO_NARROW8")) return;
    ng = (int) shape.dense.slots;
  }
  q.plane_groups = ng;
  const CodegenEnv env = row_env(shape);
  int budget = ng <= 4 ? 56 : 92;   // u32 accumulator registers per thread
  if (const char* e = getenv("EVQGPU_PLANE_BUDGET")) budget = atoi(e);
  auto need_packed = [&](int col) {
    for (int c : q.narrow_col)
      if (c == col) return;
    q.narrow_col.push_back(col);
  };
  auto operand = [&](std::vector<evqgpu_query::PlaneOperand>& list, const evqgpu_query::PlaneOperand& o) -> int {
    for (size_t i = 0; i < list.size(); ++i)
      if (list[i].expr && o.expr && list[i].expr->signature() == o.expr->signature()) return (int) i;
    list.push_back(o);
    return (int) list.size() - 1;
  };
  auto planes_of = [&](const Expr* e) -> int {
    if (is_packed_byte_column(e, shape)) return 1;
    const uint32_t b = expr_value_bits(e, env);
    return (int) ((b + 7) / 8);
  };
  {   // the rows counter
    evqgpu_query::PlaneSum ps;
    ps.word = 0;
    ps.plane0 = q.nnarrow;
    q.nnarrow += 1;
    q.state_narrow[0] = 0;
    q.plane_sums.push_back(ps);
  }
  for (const auto& item : q.select) {
    if (!item.agg || item.state0 <= 0 || q.state_narrow[item.state0] >= 0) continue;
    if (q.state_keys[item.state0].compare(0, 4, "sum:") != 0) continue;
    const FnInfo& fi = item.agg->info();
    if (fi.args.empty() || fi.args[0] != EVQ_UINT64) continue;
    const Expr* arg = item.agg->args[0].get();
    if (expr_may_raise(arg)) continue;
    // candidates (W, B): the whole argument, or the two ways to split a top-level product
    struct Cand { const Expr* w; const Expr* b; };
    std::vector<Cand> cands = {{arg, nullptr}};
    if (arg->op == EVQ_X_CALL && arg->info().fn == Fn::MUL && arg->args.size() == 2) {
      cands.push_back({arg->args[0].get(), arg->args[1].get()});
      cands.push_back({arg->args[1].get(), arg->args[0].get()});
    }
    const Cand* best = nullptr;
    int best_planes = 0;
    for (const auto& c : cands) {
      if (expr_value_max(c.w, env) > 0xffffffffull) continue;
      if (c.b && expr_value_max(c.b, env) > 255) continue;
      const int np = planes_of(c.w);
      if (!best || np < best_planes) { best = &c; best_planes = np; }
    }
    if (!best || (q.nnarrow + best_planes) * ng > budget) continue;
    evqgpu_query::PlaneOperand w;
    w.expr = best->w;
    w.nplanes = best_planes;
    if (is_packed_byte_column(best->w, shape)) { w.packed_col = (int) best->w->col; need_packed(w.packed_col); }
    evqgpu_query::PlaneSum ps;
    ps.word = item.state0;
    ps.w = operand(q.plane_w, w);
    ps.nplanes = best_planes;
    if (best->b) {
      evqgpu_query::PlaneOperand b;
      b.expr = best->b;
      if (is_packed_byte_column(best->b, shape)) {
        b.packed_col = (int) best->b->col;
      } else if (best->b->op == EVQ_X_CALL && best->b->args.size() == 2 &&
                 (best->b->info().fn == Fn::ADD || best->b->info().fn == Fn::SUB)) {
        // literal +- 1-byte column, decided per byte without carries or borrows (the value range says so): done on the
        // packed bytes of 4 rows at once
        const Expr* x = best->b->args[0].get();
        const Expr* y = best->b->args[1].get();
        const bool add = best->b->info().fn == Fn::ADD;
        if (x->op == EVQ_X_LITERAL && x->type == EVQ_UINT64 && x->imm <= 255 && is_packed_byte_column(y, shape) &&
            (add || expr_value_max(y, env) <= x->imm)) {
          b.packed_col = (int) y->col; b.swar = add ? 2 : 1; b.swar_lit = x->imm;
        } else if (add && y->op == EVQ_X_LITERAL && y->type == EVQ_UINT64 && y->imm <= 255 && is_packed_byte_column(x, shape)) {
          b.packed_col = (int) x->col; b.swar = 2; b.swar_lit = y->imm;
        }
      }
      if (b.packed_col >= 0) need_packed(b.packed_col);
      ps.b = operand(q.plane_b, b);
    }
    ps.plane0 = q.nnarrow;
    q.nnarrow += best_planes;
    q.state_narrow[item.state0] = (int) q.plane_sums.size();
    q.plane_sums.push_back(ps);
  }
  // "seen" counters (non-NULL arguments of min / max / mean) of a bare optional column: W = 1, B = the column's presence byte
  for (const auto& item : q.select) {
    if (!item.agg || item.state_seen <= 0 || q.state_narrow[item.state_seen] >= 0 || item.agg->args.empty()) continue;
    const Expr* arg = item.agg->args[0].get();
    if (arg->op != EVQ_X_INPUT || arg->col >= shape.cols.size()) continue;
    const ColSig& c = shape.cols[arg->col];
    if (!c.used || !c.nullable || c.dmax != 1) continue;
    if ((q.nnarrow + 1) * ng > budget) continue;
    evqgpu_query::PlaneOperand b;
    b.expr = arg;
    b.presence_col = (int) arg->col;
    evqgpu_query::PlaneSum ps;
    ps.word = item.state_seen;
    ps.w = -1;
    ps.b = -1;
    for (size_t i = 0; i < q.plane_b.size(); ++i)
      if (q.plane_b[i].presence_col == b.presence_col) ps.b = (int) i;
    if (ps.b < 0) { q.plane_b.push_back(b); ps.b = (int) q.plane_b.size() - 1; }
    ps.nplanes = 1;
    ps.plane0 = q.nnarrow;
    q.nnarrow += 1;
    q.state_narrow[item.state_seen] = (int) q.plane_sums.size();
    q.plane_sums.push_back(ps);
  }
  // dense slots on the packed key bytes: every key a bare 1-byte column whose range the statistics bound inside the slots
  {
    // (a NULL-able key: its packed byte is 0 where the row is NULL and the slot index of NULL is added from the presence
    // byte - index = value + null_idx * (1 - present), bytewise without carries)
    bool ok = !q.group.empty() && shape.dense.slots <= 7;
    for (size_t i = 0; ok && i < q.group.size(); ++i) {
      const DenseMap& dm = shape.dense;
      const bool may_null = dm.key_null_idx[i] != ~0ull;
      ok = is_packed_byte_column(q.group[i].get(), shape) && dm.key_min[i] == 0 &&
           (may_null || !shape.cols[q.group[i]->col].nullable) &&
           expr_value_max(q.group[i].get(), env) <= dm.key_range[i] - (may_null ? 2 : 1);
      if (ok && may_null) {   // the presence bytes of the key column are needed
        bool have = false;
        for (const auto& b : q.plane_b) have = have || b.presence_col == (int) q.group[i]->col;
        if (!have) {
          evqgpu_query::PlaneOperand b;
          b.expr = q.group[i].get();
          b.presence_col = (int) q.group[i]->col;
          q.plane_b.push_back(b);
        }
      }
    }
    if (ok && q.distinct_args.empty() && !getenv("EVQGPU_NO_SWAR_SLOTS")) {   // (count_distinct needs the slot per row)
      q.swar_slots = true;
      for (const auto& g : q.group) need_packed((int) g->col);
    }
  }
  // re-number the words that stay in shared memory
  q.nstate_smem = 0;
  for (size_t i = 0; i < q.state_ops.size(); ++i) {
    q.state_smem[i] = -1;
    if (!q.state_global[i] && q.state_narrow[i] < 0) q.state_smem[i] = q.nstate_smem++;
  }
  std::ostringstream sg;
  sg << "S" << (int) q.swar_slots << "G" << ng << ";";
  for (const auto& ps : q.plane_sums)
    sg << ps.word << ":" << (ps.w >= 0 ? q.plane_w[ps.w].expr->signature() : std::string("1")) << "*"
       << (ps.b >= 0 ? (q.plane_b[ps.b].presence_col >= 0 ? "present:" : "") + q.plane_b[ps.b].expr->signature() : std::string("1")) << "/" << ps.nplanes << ";";
  q.plane_sig = sg.str();
}

// Upper bound on the bit length of a uint64-valued expression, from the column statistics of the scanned tables
// (Column::value_bits).  64 = unknown / may wrap.
static uint32_t expr_bits(const Expr* e, const KernelShape& shape) {
  auto sat = [](uint32_t b) { return b > 64u ? 64u : b; };
  switch (e->op) {
    case EVQ_X_INPUT:
      if (e->col >= shape.cols.size()) return 64;
      if (shape.cols[e->col].sql_type == EVQ_BOOL) return 1;
      if (shape.cols[e->col].sql_type == EVQ_FLOAT64) return 64;
      return sat(shape.cols[e->col].bits);
    case EVQ_X_LITERAL: {
      if (e->type == EVQ_BOOL) return 1;
      if (e->type != EVQ_UINT64 && e->type != EVQ_TIMESTAMP64) return 64;
      uint32_t b = 0;
      for (uint64_t v = e->imm; v; v >>= 1) ++b;
      return b ? b : 1;
    }
    case EVQ_X_IF: return std::max(expr_bits(e->args[1].get(), shape), expr_bits(e->args[2].get(), shape));
    case EVQ_X_CALL: break;
    default: return 64;
  }
  const FnInfo& fi = e->info();
  if (fi.ret != EVQ_UINT64 && fi.ret != EVQ_TIMESTAMP64 && fi.ret != EVQ_BOOL) return 64;
  if (fi.ret == EVQ_BOOL) return 1;
  if (fi.args.empty() || (fi.args[0] != EVQ_UINT64 && fi.args[0] != EVQ_TIMESTAMP64)) return 64;
  switch (fi.fn) {
    case Fn::ADD: return sat(std::max(expr_bits(e->args[0].get(), shape), expr_bits(e->args[1].get(), shape)) + 1);
    case Fn::MUL: return sat(expr_bits(e->args[0].get(), shape) + expr_bits(e->args[1].get(), shape));
    case Fn::DIV: return expr_bits(e->args[0].get(), shape);
    case Fn::MOD: return std::min(expr_bits(e->args[0].get(), shape), expr_bits(e->args[1].get(), shape));
    case Fn::DATE_TRUNC: return expr_bits(e->args[1].get(), shape);
    default: return 64;   // sub may wrap, conversions may reinterpret
  }
}

static int carry_word_of(const evqgpu_query& q, int sum_word) {
  for (size_t i = 0; i < q.state_carry_of.size(); ++i)
    if (q.state_carry_of[i] == sum_word) return (int) i;
  return -1;
}

// The per-row aggregate updates, one per state WORD (not per select item): spelled through the macros of the variant
//   EVQ_UPD(word, op, v)            state[word] = combine<op>(state[word], v)
//   EVQ_UPD_C(word, carry, v)       state[word] += v; if that wrapped: state[carry] += 1
static void gen_updates(std::ostringstream& os, const evqgpu_query& q, const KernelShape& shape) {
  CodegenEnv env = row_env(shape);
  std::vector<bool> done(q.state_ops.size(), false);
  for (size_t w = 0; w < q.state_ops.size(); ++w)
    if (w < q.state_narrow.size() && q.state_narrow[w] >= 0) done[w] = true;   // fed by evq_accumulate_narrow
  if (!done[0]) os << "  EVQ_UPD(0, " << OP_ADD_U64 << ", 1ull);\n";   // rows per group
  done[0] = true;
  for (const auto& item : q.select) {
    if (item.first && item.state_first >= 0 && !done[item.state_first]) {
      // the value of the group's first row (groupby.cc:161-172): smallest row ordinal wins, one 128-bit CAS
      done[item.state_first] = true;
      Code c = gen_expr(item.expr.get(), env);
      os << "  {\n    const u64 t = (u64) ((" << c.tag << ") != 0u);\n    const u64 v = t ? 0ull : " << as_bits(c, item.expr->type) << ";\n";
      os << "    evq_first_update(EVQ_GPTR(" << item.state_first << "), (row.ord << 1) | t, v);\n  }\n";
    }
    if (!item.agg) continue;
    const FnInfo& fi = item.agg->info();
    const Expr* arg = item.agg->args.empty() ? nullptr : item.agg->args[0].get();
    if (fi.fn == Fn::COUNT_DISTINCT) continue;   // evq_accumulate_distinct
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
    const int w = item.state0;
    if (!done[w]) {
      done[w] = true;
      os << "  {\n";
      const int op = q.state_ops[w];
      if (q.state_keys[w].compare(0, 4, "sum:") == 0) {
        // sum_*: acc += v; a NULL contributes its value bits, which are 0 (aggregate.cc:184-219)
        int cw = carry_word_of(q, w);
        // the wrap check per row is dropped where the accumulator provably cannot wrap: a thread-private word needs more
        // than 2^32 rows of < 2^32 values, a global word more than 2^40 rows of < 2^24 values (flushes and merges still
        // propagate carries, the layout does not change)
        const uint32_t vb = ty == EVQ_UINT64 ? expr_bits(arg, shape) : 64;
        if (cw >= 0 && ((shape.tier == 1 && vb <= 32) || (shape.tier == 2 && vb <= 24))) cw = -1;
        os << "    const u64 v = " << as_bits(c, ty) << ";\n";
        if (cw >= 0) os << "    EVQ_UPD_C(" << w << ", " << cw << ", v);\n";
        else os << "    EVQ_UPD(" << w << ", " << OP_ADD_U64 << ", v);\n";
      } else if (op == OP_ADD_F64) {
        // double sums; NULL rows are skipped (their value bits are 0 anyway, but -0.0 + 0.0 and NaN payloads differ)
        const bool skip_null = c.tag != "0u" && fi.fn == Fn::MEAN;
        if (skip_null) os << "    if (!(" << c.tag << "))\n  ";
        os << "    EVQ_UPD(" << w << ", " << OP_ADD_F64 << ", evq_bits((f64) (" << c.value << ")));\n";
      } else {   // min / max
        if (c.tag != "0u") os << "    if (!(" << c.tag << "))\n  ";
        os << "    EVQ_UPD(" << w << ", " << op << ", " << as_bits(c, ty) << ");\n";
      }
      os << "  }\n";
    }
    const int sw = item.state_seen;
    if (sw > 0 && !done[sw]) {
      done[sw] = true;
      os << "  if (!(" << c.tag << ")) EVQ_UPD(" << sw << ", " << OP_ADD_U64 << ", 1ull);\n";
    }
  }
}

static void gen_general_layout(std::ostringstream& os, const KernelShape& shape) {
  const size_t ncols = shape.cols.size();
  // ---- row + prep structs
  os << "struct EvqRow {\n";
  for (size_t i = 0; i < ncols; ++i)
    if (shape.cols[i].used) os << "  " << col_ctype(shape.cols[i].sql_type) << " c" << i << "; u32 t" << i << ";\n";
  os << "  u64 ord;\n  u32 _unused;\n};\n";
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
  os << "  row.ord = P.ord_base + P.tile_row_base * EVQ_TILE_ROWS + T.row0 + r;\n";   // table order over the partitions of the query
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
  for (size_t i = 0; i < ncols; ++i) {
    if (!shape.cols[i].used) continue;
    os << "  " << fast_ctype(shape.cols[i]) << " c" << i << ";\n";
    if (shape.cols[i].nullable) os << "  u32 t" << i << ";\n";   // STag: 1 = NULL
  }
  os << "  u64 ord;\n  u32 _unused;\n};\n";
  os << "struct EvqCols {\n  u64 ord0;\n  u32 nbuf;\n";
  for (size_t i = 0; i < ncols; ++i)
    if (shape.cols[i].used && shape.cols[i].nullable) os << "  u32 n" << i << ", r" << i << ";\n";   // presence bits of the thread's rows, ordinal of its first value
  for (size_t i = 0; i < ncols; ++i)
    if (shape.cols[i].used) os << "  " << fast_ctype(shape.cols[i]) << " c" << i << "[EVQ_RPT];\n";
  for (size_t i = 0; i < ncols; ++i)
    if (shape.cols[i].used && shape.cols[i].packed) os << "  u32 p" << i << "[EVQ_RPT / 4];\n";   // the raw bytes, 4 rows per word
  for (size_t i = 0; i < ncols; ++i)
    if (shape.cols[i].used && shape.cols[i].presence) os << "  u32 q" << i << "[EVQ_RPT / 4];\n";  // presence bytes (1 = not NULL), 4 rows per word
  os << "  u32 _unused;\n};\n";
  const int ngen = std::max(1, shape.ngen);
  os << "struct EvqFastPrep {\n  bool general[" << ngen << "];\n  u32 start[" << ngen << "];\n};\n";

  // ---- boundary search of the variable-length columns: ONE consumer barrier per tile, only when a tile needs it
  os << "__device__ __forceinline__ void evq_fast_prep(const EvqTile& T, const EvqScanParams& P, EvqFastScratch* scr, EvqFastPrep& prep) {\n";
  if (shape.ngen > 0) {
    os << "  bool any = false;\n";
    for (size_t i = 0; i < ncols; ++i) {
      const ColSig& c = shape.cols[i];
      if (!c.used || c.gen_slot < 0) continue;
      os << "  evq_fast_prep_a<" << c.data_stream << ", " << c.gen_slot << ", " << c.leb_len << ">(T, P, scr, prep.general[" << c.gen_slot
         << "]);\n  any = any || prep.general[" << c.gen_slot << "];\n  prep.start[" << c.gen_slot << "] = 0u;\n";
    }
    os << "  if (any) {\n    evq_cons_sync();\n";
    for (size_t i = 0; i < ncols; ++i) {
      const ColSig& c = shape.cols[i];
      if (!c.used || c.gen_slot < 0) continue;
      os << "    if (prep.general[" << c.gen_slot << "]) prep.start[" << c.gen_slot << "] = evq_fast_prep_b<" << c.data_stream << ", "
         << c.gen_slot << ">(T, P, scr);\n";
    }
    os << "  }\n";
  }
  os << "}\n";

  // ---- optional columns: presence bits of the thread's rows + ordinal of its first value (one barrier per tile)
  os << "__device__ __forceinline__ void evq_fast_nulls(const EvqTile& T, const EvqScanParams& P, EvqFastScratch* scr, u32 buf, u32 nvalid, EvqCols& cols) {\n";
  if (shape.nnull > 0) {
    // presence bits of the thread's rows, then ONE warp scan per three optional columns: the per-thread counts (<= 8) travel
    // in 10-bit fields of one word (a thread's exclusive prefix is at most 8 * 127 = 1016)
    std::vector<size_t> ncol;
    for (size_t i = 0; i < ncols; ++i)
      if (shape.cols[i].used && shape.cols[i].nullable) ncol.push_back(i);
    for (size_t i : ncol)
      os << "  cols.n" << i << " = evq_fast_presence<" << shape.cols[i].level_stream << ">(T, P) & ((1u << nvalid) - 1u);\n";
    for (size_t p0 = 0; p0 < ncol.size(); p0 += 3) {
      os << "  u32 pk" << p0 / 3 << " = evq_fast_null_scan(";
      for (size_t j = p0; j < std::min(p0 + 3, ncol.size()); ++j) os << (j > p0 ? " | (" : "") << "__popc(cols.n" << ncol[j] << ")" << (j > p0 ? " << " + std::to_string(10 * (j - p0)) + ")" : "");
      os << ", scr, buf, " << p0 / 3 << ", T.ctid);\n";
    }
    os << "  cols.nbuf = buf;\n";
    for (size_t i = 0; i < ncols; ++i) {   // values of the staged columns, by value ordinal (needs no ranks)
      const ColSig& c = shape.cols[i];
      if (!c.used || c.nv_slot < 0) continue;
      uint32_t lmin = 1;   // (of the values in the stream: NULL rows have none)
      for (uint64_t m = c.vmin_present >> 7; m; m >>= 7) ++lmin;
      lmin = std::min(lmin, c.leb_len);
      os << "  evq_fast_stage_vals<" << c.data_stream << ", " << c.sub_stream << ", " << c.leb_len << ", " << lmin << ", " << c.nv_slot
         << ">(T, P, scr, buf);\n";
    }
    os << "  evq_cons_sync();\n";
    for (size_t p0 = 0; p0 < ncol.size(); p0 += 3) {
      os << "  pk" << p0 / 3 << " = evq_fast_null_rank(pk" << p0 / 3 << ", scr, buf, " << p0 / 3 << ", T.ctid);\n";
      for (size_t j = p0; j < std::min(p0 + 3, ncol.size()); ++j)
        os << "  cols.r" << ncol[j] << " = (pk" << p0 / 3 << " >> " << 10 * (j - p0) << ") & 1023u;\n";
    }
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
    if (c.nullable) {
      // NULL: value 0, tag STAG_NULL (CSTableScan.cc:877-890); only the present values are in the data stream
      const std::string pbr = "cols.n" + std::to_string(i) + ", cols.r" + std::to_string(i);
      std::string raw_tn = "u32";
      os << "  {\n";
      switch (c.kind) {
        case EVQ_KIND_PLAIN64: raw_tn = "u64"; os << "    u64 raw[EVQ_RPT];\n    evq_fast_ldn_plain64<" << S << ">(T, P, " << pbr << ", raw);\n"; break;
        case EVQ_KIND_PLAIN32: os << "    u32 raw[EVQ_RPT];\n    evq_fast_ldn_plain32<" << S << ">(T, P, " << pbr << ", raw);\n"; break;
        case EVQ_KIND_BITPACK: os << "    u32 raw[EVQ_RPT];\n    evq_fast_ldn_bitpack<" << S << ">(T, P, " << pbr << ", raw);\n"; break;
        default:
          if (c.nv_slot >= 0) {
            os << "    u32 raw[EVQ_RPT];\n    evq_fast_gather_vals<" << c.nv_slot << ">(scr, cols.nbuf, " << pbr << ", raw);\n";
          } else if (c.leb_len <= 1 && c.packed) {
            os << "    u32 raw[EVQ_RPT];\n    evq_fast_ldn_leb1p<" << S << ">(T, P, " << pbr << ", raw, cols.p" << i << ");\n";
          } else if (c.leb_len <= 1) {
            os << "    u32 raw[EVQ_RPT];\n    evq_fast_ldn_leb1<" << S << ">(T, P, " << pbr << ", raw);\n";
          } else {
            raw_tn = "u64";
            os << "    u64 raw[EVQ_RPT];\n    evq_fast_ldn_leb<" << S << ", " << (c.leb_uniform ? -1 : c.sub_stream) << ", " << c.leb_len << ">(T, P, " << pbr
               << ", raw);\n";
          }
          break;
      }
      std::string conv;
      switch (c.sql_type) {
        case EVQ_FLOAT64: conv = "evq_f64(raw[k])"; break;
        case EVQ_BOOL: conv = "(u32) (raw[k] > 0)"; break;
        default: conv = std::string("(") + ct + ") raw[k]"; break;
      }
      os << "#pragma unroll\n    for (int k = 0; k < EVQ_RPT; ++k) cols.c" << i << "[k] = " << conv << ";\n  }\n";
      if (c.presence) os << "  evq_presence_bytes(cols.n" << i << ", cols.q" << i << ");\n";
      (void) raw_tn;
      continue;
    }
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
        if (c.leb_len <= 1 && c.packed) os << "    evq_fast_ld_leb1p<" << S << ">(T, P, raw, cols.p" << i << ");\n";
        else if (c.leb_len <= 1) os << "    evq_fast_ld_leb1<" << S << ">(T, P, raw);\n";
        else
        {
          // where this thread's 8 values start: from the column's sub-index (staged with the tile), or searched by evq_fast_prep
          std::string general, start;
          if (c.leb_uniform) {   // every value has leb_len bytes: the fixed-stride paths only
            general = "false";
            start = "0u";
          } else if (c.sub_stream >= 0) {
            general = "evq_fast_general<" + S + ", " + std::to_string(c.leb_len) + ">(T)";
            start = "evq_fast_substart<" + std::to_string(c.sub_stream) + ">(T, P)";
          } else {
            general = "prep.general[" + std::to_string(c.gen_slot) + "]";
            start = "prep.start[" + std::to_string(c.gen_slot) + "]";
          }
          // second template argument: 2 = two sub-index entry points packed in `start`, 1 = one searched entry point
          // (leb32 only) fourth: the fewest bytes a value of the column can have, from its minimum
          uint32_t lmin = 1;
          for (uint64_t m = c.vmin >> 7; m; m >>= 7) ++lmin;
          lmin = std::min(lmin, c.leb_len);
          os << "    evq_fast_ld_leb" << (c.leb_len <= 4 ? "32" : "64") << "<" << S << ", " << (c.sub_stream >= 0 ? 2 : 1) << ", " << c.leb_len;
          if (c.leb_len <= 4) os << ", " << lmin;
          os << ">(T, P, " << general << ", " << start << ", raw);\n";
        }
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
  os << "__device__ __forceinline__ void evq_fast_row(const EvqCols& cols, int k, EvqRow& row) {\n  row.ord = cols.ord0 + k;\n";
  for (size_t i = 0; i < ncols; ++i) {
    if (!shape.cols[i].used) continue;
    os << "  row.c" << i << " = cols.c" << i << "[k];\n";
    if (shape.cols[i].nullable) os << "  row.t" << i << " = ((cols.n" << i << " >> k) & 1u) ^ 1u;\n";
  }
  os << "}\n";
}

static std::string gen_row_functions(const evqgpu_query& q, const KernelShape& shape) {
  std::ostringstream os;
  if (shape.fast) gen_fast_layout(os, shape);
  else gen_general_layout(os, shape);

  // ---- WHERE
  CodegenEnv env = row_env(shape);
  {
    // a WHERE program without integer division cannot raise: rows past the end of a short tile may then run it too,
    // which lets the kernel fold the validity test into the predicate instead of branching around it
    std::function<bool(const Expr*)> may_raise = [&](const Expr* e) -> bool {
      if (!e) return false;
      if (e->op == EVQ_X_CALL && (e->info().fn == Fn::DIV || e->info().fn == Fn::MOD) && e->info().args[0] != EVQ_FLOAT64) return true;
      for (const auto& a : e->args)
        if (may_raise(a.get())) return true;
      return false;
    };
    if (!may_raise(q.where.get())) os << "#define EVQ_WHERE_PURE 1\n";
  }
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

  // ---- partitioned aggregation: a row that passed WHERE as a record (the columns the keys and aggregate arguments read)
  if (shape.part_bits > 0) {
    // the record is PACKED: every column takes the bits its largest value needs (the column statistics are part of the
    // kernel's signature), a NULL-able column one more for its tag - C4's (24-bit key, 20-bit value) rows are ONE word
    // instead of two, which halves the traffic of all three passes
    const RecordLayout L = record_layout(shape);
    std::vector<std::string> words(L.nwords, "0ull");
    for (const auto& f : L.fields) {
      const std::string ct = fast_ctype(shape.cols[f.col]);
      std::string v = f.is_tag ? "(u64) row.t" + std::to_string(f.col)
                               : (ct == "f64" ? "evq_bits(row.c" + std::to_string(f.col) + ")" : "(u64) row.c" + std::to_string(f.col));
      if (f.is_tag) v = "(" + v + " & 1ull)";
      words[f.word] += " | (" + v + " << " + std::to_string(f.shift) + ")";
    }
    os << "__device__ __forceinline__ void evq_row_store(const EvqRow& row, u64* rec) {\n";
    if (words.size() == 2) {   // one 16-byte store (into the tile's bin in shared memory)
      os << "  *(ulonglong2*) rec = make_ulonglong2(" << words[0] << ", " << words[1] << ");\n";
    } else {
      for (size_t j = 0; j < words.size(); ++j) os << "  rec[" << j << "] = " << words[j] << ";\n";
    }
    os << "}\n";
    os << "__device__ __forceinline__ void evq_row_load(const u64* rec, EvqRow& row) {\n";
    if (words.size() == 2) {
      os << "  u64 w0, w1;\n  asm volatile(\"ld.global.cs.v2.u64 {%0, %1}, [%2];\" : \"=l\"(w0), \"=l\"(w1) : \"l\"(rec));\n  const u64 w[2] = {w0, w1};\n";
    } else {
      os << "  u64 w[" << words.size() << "];\n";
      for (size_t j = 0; j < words.size(); ++j) os << "  w[" << j << "] = __ldcs(rec + " << j << ");\n";
    }
    for (const auto& f : L.fields) {
      const std::string ct = fast_ctype(shape.cols[f.col]);
      std::string x = "(w[" + std::to_string(f.word) + "] >> " + std::to_string(f.shift) + ")";
      if (f.bits < 64) x = "(" + x + " & " + std::to_string((1ull << f.bits) - 1) + "ull)";
      if (f.is_tag) os << "  row.t" << f.col << " = (u32) " << x << ";\n";
      else os << "  row.c" << f.col << " = " << (ct == "f64" ? "evq_f64(" + x + ")" : "(" + ct + ") " + x) << ";\n";
    }
    os << "  row.ord = 0;\n}\n";
  }

  // ---- count_distinct: insert (group, value) into the set of the argument; the inserting thread counts it
  if (!q.distinct_args.empty() && (q.flags & EVQGPU_QUERY_GROUPBY)) {
    os << "#define EVQ_NDISTINCT " << q.distinct_args.size() << "\n";
    os << "__device__ __forceinline__ void evq_accumulate_distinct(const EvqRow& row, u64 gid, u64* state, const EvqScanParams& P, u32& err) {\n";
    for (size_t d = 0; d < q.distinct_args.size(); ++d) {
      Code c = gen_expr(q.distinct_args[d], env);
      os << "  {\n    u64 k2[2] = {gid, " << as_bits(c, EVQ_UINT64) << "};\n    const u32 t2[2] = {0u, 0u};\n";
      os << "    if (!evq_ht_upsert<2>(P.dt[" << d << "], k2, t2, state + " << q.distinct_word[d] << ")) err |= EVQ_ERR_TABLE_FULL;\n  }\n";
    }
    os << "}\n";
  }

  // ---- direct-addressed group array (tier 2 without a hash table): key tuple -> slot from the key bounds, which are
  // kernel PARAMETERS here (one kernel for every table set); a key outside the bounds is an internal error
  if (shape.tier == 2 && shape.dense_global) {
    os << "#define EVQ_DENSE_GLOBAL 1\n";
    os << "__device__ __forceinline__ u64 evq_dense_slot_rt(const u64* key, const u32* ktag, const EvqScanParams& P, u32& err) {\n"
          "  u64 slot = 0;\n  bool ok = true;\n";
    for (size_t i = 0; i < q.group.size(); ++i) {
      os << "  {\n    const u64 d = key[" << i << "] - P.key_min[" << i << "];\n";
      os << "    const bool isnull = ktag[" << i << "] != 0u;\n";
      os << "    ok = ok && (isnull ? P.key_null_idx[" << i << "] != ~0ull : d <= P.key_span[" << i << "]);\n";
      os << "    slot += (isnull ? P.key_null_idx[" << i << "] : d) * P.key_stride[" << i << "];\n  }\n";
    }
    os << "  if (!ok) {\n    err |= EVQ_ERR_SLOT_RANGE;\n    return ~0ull;\n  }\n  return slot;\n}\n";
  }

  // ---- dense tier: group key tuple -> accumulator slot, with the key bounds of this execution as constants
  if (shape.tier == 1 && shape.g1 > 1) {
    bool all_proven = true;
    for (size_t i = 0; i < q.group.size(); ++i) {
      const DenseMap& dm = shape.dense;
      const bool may_null = dm.key_null_idx[i] != ~0ull;
      const uint64_t span = dm.key_range[i] - (may_null ? 2 : 1);
      if (may_null || dm.key_min[i] != 0 || expr_value_max(q.group[i].get(), env) > span) all_proven = false;
    }
    if (all_proven) os << "#define EVQ_SLOT_ALWAYS_VALID 1\n";
    os << "__device__ __forceinline__ u32 evq_dense_slot(const u64* key, const u32* ktag, u32& err) {\n  u32 slot = 0;\n  bool ok = true;\n";
    for (size_t i = 0; i < q.group.size(); ++i) {
      const DenseMap& dm = shape.dense;
      const bool may_null = dm.key_null_idx[i] != ~0ull;
      const uint64_t span = dm.key_range[i] - (may_null ? 2 : 1);   // largest non-NULL index
      os << "  {\n    const u64 d = key[" << i << "] - " << dm.key_min[i] << "ull;\n";
      // the range check is dropped when the column statistics already bound the key inside the slot range
      const bool proven = !may_null && dm.key_min[i] == 0 && expr_value_max(q.group[i].get(), env) <= span;
      if (proven) {
        os << "    const u32 idx = (u32) d;\n";
      } else if (may_null) {
        os << "    const u32 idx = ktag[" << i << "] ? " << dm.key_null_idx[i] << "u : (u32) d;\n";
        os << "    ok = ok && (ktag[" << i << "] || d <= " << span << "ull);\n";
      } else {
        os << "    const u32 idx = (u32) d;\n    ok = ok && d <= " << span << "ull;\n";
      }
      os << "    slot += idx * " << dm.key_stride[i] << "u;\n  }\n";
    }
    os << "  if (!ok) {\n    err |= EVQ_ERR_SLOT_RANGE;\n    return ~0u;\n  }\n  return slot;\n}\n";
  }

  const int nstate = (int) q.state_ops.size();
  // thread-private -> warp -> global merge of one accumulator word; sum words with a carry partner count the wraps of
  // every addition on the way (the butterfly leaves the same totals in all lanes)
  auto gen_flush_word = [&](int w, const std::string& load, const std::string& global_base) {
    const int op = q.state_ops[w];
    const int cw = carry_word_of(q, w);
    os << "  {\n    u64 v = " << load << ";\n";
    if (cw >= 0) {
      os << "    u64 c = 0;\n#pragma unroll\n    for (int o = 16; o > 0; o >>= 1) {\n      const u64 n = __shfl_xor_sync(0xffffffffu, v, o);\n"
            "      c += __shfl_xor_sync(0xffffffffu, c, o);\n      v += n;\n      c += v < n ? 1ull : 0ull;\n    }\n";
      os << "    if (evq_lane() == 0 && (v | c)) {\n      const u64 old = atomicAdd(" << global_base << " + " << w << ", v);\n"
            "      if (old + v < v) ++c;\n      if (c) atomicAdd(" << global_base << " + " << cw << ", c);\n    }\n  }\n";
    } else {
      os << "#pragma unroll\n    for (int o = 16; o > 0; o >>= 1) v = evq_state_combine<" << op << ">(v, __shfl_xor_sync(0xffffffffu, v, o));\n";
      os << "    if (evq_lane() == 0 && v != evq_state_identity<" << op << ">()) evq_state_atomic<" << op << ">(" << global_base
         << " + " << w << ", v);\n  }\n";
    }
  };
  if (shape.tier == 1) {
    const int nsm = std::max(1, q.nstate_smem);
    if (shape.g1 > 1) {
      os << "#define EVQ_SIDX(g, st) ((((g) * " << nsm << ") + (st)) * EVQ_NCONS + tid)\n";
      os << "__device__ __forceinline__ void evq_state_init_slot(u64* sacc, u32 g, u32 tid) {\n";
      for (int w = 0; w < nstate; ++w)
        if (q.state_smem[w] >= 0) os << "  sacc[EVQ_SIDX(g, " << q.state_smem[w] << ")] = evq_state_identity<" << q.state_ops[w] << ">();\n";
      os << "}\n";
      for (int w = 0; w < nstate; ++w) os << "#define EVQ_SM_" << w << " " << q.state_smem[w] << "\n";
      // one base address per row (group g, this thread); the words of the group sit at constant offsets from it
      os << "#define EVQ_UPD(st, op, v) _acc[EVQ_SM_##st * EVQ_NCONS] = evq_state_combine<op>(_acc[EVQ_SM_##st * EVQ_NCONS], (v))\n";
      os << "#define EVQ_UPD_C(st, cw, v) { const u64 _o = _acc[EVQ_SM_##st * EVQ_NCONS]; const u64 _n = _o + (v); "
            "_acc[EVQ_SM_##st * EVQ_NCONS] = _n; if (_n < _o) atomicAdd(dense_state + (u64) g * " << nstate << " + (cw), 1ull); }\n";
      os << "#define EVQ_GPTR(st) (dense_state + (u64) g * " << nstate << " + (st))\n";
      os << "__device__ __forceinline__ void evq_accumulate_smem(const EvqRow& row, u64* sacc, u32 g, u32 tid, u64* dense_state, u32& err) {\n";
      os << "  u64* _acc = sacc + g * " << nsm << "u * EVQ_NCONS + tid;\n";
      gen_updates(os, q, shape);
      os << "}\n#undef EVQ_UPD\n#undef EVQ_UPD_C\n#undef EVQ_GPTR\n";
      os << "__device__ __forceinline__ void evq_state_flush_smem(u64* sacc, u32 g, u32 tid, u64* dense_state) {\n";
      for (int w = 0; w < nstate; ++w)
        if (q.state_smem[w] >= 0)
          gen_flush_word(w, "sacc[EVQ_SIDX(g, " + std::to_string(q.state_smem[w]) + ")]", "dense_state + (u64) g * " + std::to_string(nstate));
      os << "}\n";
      if (q.nnarrow > 0) {
        // byte-plane sums (layout_narrow): u32 registers per (plane, group), fed 4 rows at a time.  `selector` holds one
        // nibble per row of the quad: its dense slot, or >= 4 when the row did not pass WHERE; PRMT turns it into the byte
        // mask of group g (0xff where the row belongs to it).  Sums without a byte operand use the mask itself as the
        // signed operand -1, i.e. they accumulate the NEGATED sum (mod 2^32), which evq_narrow_flush undoes.
        if (q.swar_slots) {
          os << "#define EVQ_SWAR_SLOTS 1\n";
          os << "__device__ __forceinline__ u32 evq_quad_slots(const EvqCols& cols, int j) {\n  const u32 s = 0u";
          for (size_t i = 0; i < q.group.size(); ++i) {
            os << " + cols.p" << q.group[i]->col << "[j] * " << shape.dense.key_stride[i] << "u";
            if (shape.dense.key_null_idx[i] != ~0ull) {   // NULL rows (packed byte 0): + null_idx, bytewise
              char lit[32];
              snprintf(lit, sizeof(lit), "0x%08xu", (unsigned) (shape.dense.key_null_idx[i] * 0x01010101u));
              os << " + (" << lit << " - cols.q" << q.group[i]->col << "[j] * " << shape.dense.key_null_idx[i] << "u) * " << shape.dense.key_stride[i] << "u";
            }
          }
          os << ";   // the slot of each row in its byte (< 8: no carries)\n";
          os << "  return __byte_perm(s | (s >> 4), 0u, 0x4420u);   // one nibble per row\n}\n";
        }
        // a W / B expression for row kk of quad j
        auto quad_env = [&](int kk) {
          CodegenEnv e = row_env(shape);
          for (size_t i = 0; i < shape.cols.size(); ++i) {
            if (!shape.cols[i].used) continue;
            const std::string ref = "cols.c" + std::to_string(i) + "[4 * j + " + std::to_string(kk) + "]";
            const bool widen = fast_ctype(shape.cols[i]) == std::string("u32") && shape.cols[i].sql_type != EVQ_BOOL;
            e.col_value[i] = widen ? "((u64) " + ref + ")" : ref;
          }
          return e;
        };
        os << "__device__ __forceinline__ void evq_accumulate_narrow(const EvqCols& cols, int j, u32 selector, u32* nacc) {\n";
        os << "  u32 m[EVQ_NG];\n#pragma unroll\n  for (int g = 0; g < EVQ_NG; ++g)\n"
              "    m[g] = __byte_perm(g < 4 ? 0xffu << (8 * (g & 3)) : 0u, g < 4 ? 0u : 0xffu << (8 * (g & 3)), selector);\n";
        // byte planes of every distinct W
        for (size_t wi = 0; wi < q.plane_w.size(); ++wi) {
          const auto& W = q.plane_w[wi];
          if (W.packed_col >= 0) {
            os << "  const u32 w" << wi << "p0 = cols.p" << W.packed_col << "[j];\n";
            continue;
          }
          for (int kk = 0; kk < 4; ++kk) {
            Code c = gen_expr(W.expr, quad_env(kk));
            os << "  const u32 w" << wi << "r" << kk << " = (u32) (" << c.value << ");\n";
          }
          const std::string w = "w" + std::to_string(wi);
          if (W.nplanes == 1) {
            os << "  const u32 " << w << "p0 = __byte_perm(__byte_perm(" << w << "r0, " << w << "r1, 0x0040u), __byte_perm(" << w << "r2, "
               << w << "r3, 0x0040u), 0x5410u);\n";
          } else {
            os << "  const u32 " << w << "a = __byte_perm(" << w << "r0, " << w << "r1, 0x5140u), " << w << "b = __byte_perm(" << w << "r2, "
               << w << "r3, 0x5140u);\n";
            os << "  const u32 " << w << "p0 = __byte_perm(" << w << "a, " << w << "b, 0x5410u), " << w << "p1 = __byte_perm(" << w << "a, "
               << w << "b, 0x7632u);\n";
            if (W.nplanes > 2) {
              os << "  const u32 " << w << "c = __byte_perm(" << w << "r0, " << w << "r1, 0x7362u), " << w << "d = __byte_perm(" << w
                 << "r2, " << w << "r3, 0x7362u);\n";
              os << "  const u32 " << w << "p2 = __byte_perm(" << w << "c, " << w << "d, 0x5410u);\n";
              if (W.nplanes > 3) os << "  const u32 " << w << "p3 = __byte_perm(" << w << "c, " << w << "d, 0x7632u);\n";
            }
          }
        }
        // the 4 rows' bytes of every distinct B
        for (size_t bi = 0; bi < q.plane_b.size(); ++bi) {
          const auto& B = q.plane_b[bi];
          os << "  const u32 b" << bi << " = ";
          if (B.presence_col >= 0) {
            os << "cols.q" << B.presence_col << "[j];\n";
          } else if (B.packed_col >= 0 && B.swar == 0) {
            os << "cols.p" << B.packed_col << "[j];\n";
          } else if (B.packed_col >= 0) {
            char lit[32];
            snprintf(lit, sizeof(lit), "0x%08xu", (unsigned) (B.swar_lit * 0x01010101u));
            os << lit << (B.swar == 1 ? " - " : " + ") << "cols.p" << B.packed_col << "[j];\n";
          } else {
            std::string r[4];
            for (int kk = 0; kk < 4; ++kk) r[kk] = "(u32) (" + gen_expr(B.expr, quad_env(kk)).value + ")";
            os << "__byte_perm(__byte_perm(" << r[0] << ", " << r[1] << ", 0x0040u), __byte_perm(" << r[2] << ", " << r[3]
               << ", 0x0040u), 0x5410u);\n";
          }
        }
        os << "#pragma unroll\n  for (int g = 0; g < EVQ_NG; ++g) {\n";
        for (size_t bi = 0; bi < q.plane_b.size(); ++bi) os << "    const u32 b" << bi << "m = b" << bi << " & m[g];\n";
        for (const auto& ps : q.plane_sums) {
          for (int pl = 0; pl < ps.nplanes; ++pl) {
            const std::string acc = "nacc[" + std::to_string(ps.plane0 + pl) + " * EVQ_NG + g]";
            const std::string plane = ps.w < 0 ? std::string("0x01010101u") : "w" + std::to_string(ps.w) + "p" + std::to_string(pl);
            if (ps.b >= 0 && ps.w < 0) os << "    " << acc << " = __dp4a(0x01010101u, b" << ps.b << "m, " << acc << ");\n";   // a count of B's bytes
            else if (ps.b < 0) os << "    " << acc << " = evq_dp4a_us(" << plane << ", m[g], " << acc << ");\n";
            else os << "    " << acc << " = __dp4a(" << plane << ", b" << ps.b << "m, " << acc << ");\n";
          }
        }
        os << "  }\n}\n";
        // rows that passed WHERE, from the (negated) rows counters
        os << "__device__ __forceinline__ u64 evq_narrow_rows(const u32* nacc) {\n  u64 n = 0;\n#pragma unroll\n"
              "  for (int g = 0; g < EVQ_NG; ++g) n += (u64) (0u - nacc[g]);\n  return n;\n}\n";
        os << "__device__ __forceinline__ void evq_narrow_flush(const u32* nacc, u64* dense_state) {\n";
        for (int g = 0; g < q.plane_groups; ++g)
          for (const auto& ps : q.plane_sums) {
            std::string v;
            for (int pl = 0; pl < ps.nplanes; ++pl) {
              std::string a = "nacc[" + std::to_string((ps.plane0 + pl) * q.plane_groups + g) + "]";
              if (ps.b < 0) a = "(0u - " + a + ")";
              a = "((u64) " + a + " << " + std::to_string(8 * pl) + ")";
              v += (pl ? " + " : "") + a;
            }
            gen_flush_word(ps.word, v, "dense_state + " + std::to_string((uint64_t) g * nstate));
          }
        os << "}\n";
      }
    } else {
      os << "__device__ __forceinline__ void evq_state_init_regs(u64* acc) {\n";
      for (int w = 0; w < nstate; ++w)
        if (q.state_smem[w] >= 0) os << "  acc[" << q.state_smem[w] << "] = evq_state_identity<" << q.state_ops[w] << ">();\n";
      os << "}\n";
      for (int w = 0; w < nstate; ++w) os << "#define EVQ_SM_" << w << " " << q.state_smem[w] << "\n";
      os << "#define EVQ_UPD(st, op, v) acc[EVQ_SM_##st] = evq_state_combine<op>(acc[EVQ_SM_##st], (v))\n";
      os << "#define EVQ_UPD_C(st, cw, v) { const u64 _o = acc[EVQ_SM_##st]; const u64 _n = _o + (v); acc[EVQ_SM_##st] = _n; "
            "if (_n < _o) atomicAdd(dense_state + (cw), 1ull); }\n";
      os << "#define EVQ_GPTR(st) (dense_state + (st))\n";
      os << "__device__ __forceinline__ void evq_accumulate_regs(const EvqRow& row, u64* acc, u64* dense_state, u32& err) {\n";
      gen_updates(os, q, shape);
      os << "}\n#undef EVQ_UPD\n#undef EVQ_UPD_C\n#undef EVQ_GPTR\n";
      os << "__device__ __forceinline__ void evq_state_flush_regs(u64* acc, u64* dense_state) {\n";
      for (int w = 0; w < nstate; ++w)
        if (q.state_smem[w] >= 0) gen_flush_word(w, "acc[" + std::to_string(q.state_smem[w]) + "]", "dense_state");
      os << "}\n";
    }
  } else if (shape.tier == 2) {
    // `state` = the group's state words inside its slot (kernels/evq_abi.h EvqHashTable)
    os << "#define EVQ_UPD(st, op, v) evq_state_atomic<op>(state + (st), (v))\n";
    os << "#define EVQ_UPD_C(st, cw, v) { const u64 _v = (v); const u64 _o = atomicAdd(state + (st), _v); "
          "if (_o + _v < _v) atomicAdd(state + (cw), 1ull); }\n";
    os << "#define EVQ_GPTR(st) (state + (st))\n";
    os << "__device__ __forceinline__ void evq_accumulate_global(const EvqRow& row, u64* state, u32& err) {\n";
    gen_updates(os, q, shape);
    os << "}\n#undef EVQ_UPD\n#undef EVQ_UPD_C\n#undef EVQ_GPTR\n";
    if (shape.slice_slots > 0) {   // the same on a slot of a table slice in shared memory (`state` = its shared-window address)
      os << "#define EVQ_UPD(st, op, v) evq_state_atomic_smem<op>(state + 8u * (st), (v))\n";
      os << "#define EVQ_UPD_C(st, cw, v) { const u64 _v = (v); const u64 _o = evq_add_ret_smem(state + 8u * (st), _v); "
            "if (_o + _v < _v) evq_state_atomic_smem<EVQ_OP_ADD_U64>(state + 8u * (cw), 1ull); }\n";
      os << "#define EVQ_GPTR(st) ((u64*) 0)\n";
      os << "__device__ __forceinline__ void evq_accumulate_smem(const EvqRow& row, u32 state, u32& err) {\n";
      gen_updates(os, q, shape);
      os << "}\n#undef EVQ_UPD\n#undef EVQ_UPD_C\n#undef EVQ_GPTR\n";
    }
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
  if (shape.tier == 1 || shape.dense_global) {
    for (int s = 0; s < nstate; ++s)
      os << "  I.dense_state[i * " << nstate << " + " << s << "] = evq_state_identity<" << q.state_ops[s] << ">();\n";
  } else {
    os << "  u64* sp = I.ht.slots + i * I.ht.stride;\n  sp[0] = 0ull;\n";
    for (int s = 0; s < nstate; ++s)
      os << "  sp[" << 1 + nk + s << "] = evq_state_identity<" << q.state_ops[s] << ">();\n";
  }
  os << "}\n";

  os << "struct EvqEmitParams { const u64* dense_state; EvqHashTable ht; u64 key_min[EVQ_MAX_KEYS]; u64 key_stride[EVQ_MAX_KEYS]; "
        "u64 key_null_idx[EVQ_MAX_KEYS]; u64 key_range[EVQ_MAX_KEYS]; u64 slots; u64* out_count; u64 out_capacity; "
        "u8* out_cols[EVQ_MAX_STREAMS]; u8* out_sha; u64* out_state; };\n";
  // one group -> one result row.  `dstate`: the dense state array the group's words are read from (the tail kernel passes
  // the merged copy in shared memory); unused by the hash tier, which reads its slot
  os << "__device__ __forceinline__ void evq_emit_group(const EvqEmitParams& E, const u64 slot, const u64* dstate) {\n";
  os << "  u64 st[" << std::max(1, nstate) << "];\n  u64 key[" << std::max(1, nk) << "];\n  u32 ktag[" << std::max(1, nk) << "];\n";
  os << "  u32 err = 0;\n";
  if (shape.tier == 1 || shape.dense_global) {
    for (int s = 0; s < nstate; ++s) os << "  st[" << s << "] = dstate[slot * " << nstate << " + " << s << "];\n";
    os << "  if (st[0] == 0) return;\n";   // no row reached this group: it does not exist
    for (int i = 0; i < nk; ++i) {
      os << "  {\n    const u64 idx = (slot / E.key_stride[" << i << "]) % E.key_range[" << i << "];\n";
      os << "    ktag[" << i << "] = idx == E.key_null_idx[" << i << "] ? 1u : 0u;\n";
      os << "    key[" << i << "] = ktag[" << i << "] ? 0ull : E.key_min[" << i << "] + idx;\n  }\n";
    }
  } else {
    os << "  const u64* sp = E.ht.slots + slot * E.ht.stride;\n  const u64 fp = sp[0];\n  if (fp == 0) return;\n";
    for (int s = 0; s < nstate; ++s) os << "  st[" << s << "] = sp[" << 1 + nk + s << "];\n";
    for (int i = 0; i < nk; ++i) {
      os << "  key[" << i << "] = sp[" << 1 + i << "];\n";
      os << "  ktag[" << i << "] = (u32) (fp >> " << 2 + i << ") & 1u;\n";
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
      const std::string s1 = "st[" + std::to_string(std::max(0, item.state_seen)) + "]";
      std::string val;
      switch (fi.fn) {
        case Fn::COUNT: val = s0; break;                                            // count_get (aggregate.cc:40-42); word 0 = rows
        case Fn::COUNT_DISTINCT: val = s0; break;                                   // count_distinct_uint64_get
        case Fn::SUM: val = from_bits(s0, fi.ret); break;                           // sum_*_get
        case Fn::MIN:
        case Fn::MAX: val = "(" + s1 + " ? " + from_bits(s0, fi.ret) + " : " + from_bits("0ull", fi.ret) + ")"; break;
        case Fn::MEAN:
          // uint64 arguments: the exact 128-bit integer sum (carry word : sum word), rounded to double once
          if (item.state_carry >= 0)
            val = "((((f64) st[" + std::to_string(item.state_carry) + "]) * 18446744073709551616.0 + (f64) " + s0 + ") / (f64) " + s1 + ")";
          else
            val = "(evq_f64(" + s0 + ") / (f64) " + s1 + ")";
          break;
        default: fail(EVQGPU_ERR_UNSUPPORTED, "aggregate %s", fi.symbol.c_str());
      }
      e2.subst.push_back({item.agg->signature(), {val, "0u"}});
    }
    const int ty = item.expr->type;
    if (item.first) {   // the boxed value of the group's first row (groupby.cc:161-172, :213-215)
      const std::string v = "st[" + std::to_string(item.state_first + 1) + "]", t = "(u32) (st[" + std::to_string(item.state_first) + "] & 1ull)";
      if (ty == EVQ_BOOL) os << "  evq_store_packed2(E.out_cols[" << i << "] + out_row * 2, " << v << ", " << t << ");\n";
      else os << "  evq_store_packed9(E.out_cols[" << i << "] + out_row * 9, " << v << ", " << t << ");\n";
      continue;
    }
    Code c = gen_expr(item.expr.get(), e2);
    if (ty == EVQ_BOOL)
      os << "  evq_store_packed2(E.out_cols[" << i << "] + out_row * 2, " << as_bits(c, ty) << ", " << c.tag << ");\n";
    else
      os << "  evq_store_packed9(E.out_cols[" << i << "] + out_row * 9, " << as_bits(c, ty) << ", " << c.tag << ");\n";
  }
  if (q.flags & EVQGPU_QUERY_WIRE) {
    // PartialGroupByExpression rows: the group key is the SHA-1 of the group expressions' packed stack bytes, last
    // expression first (groupby.cc:112-135: [8 B value][1 B tag], BOOL [1 B][1 B]); the states travel raw
    os << "  {\n    u8 msg[" << std::max(1, 9 * nk) << "];\n    u32 len = 0;\n";
    for (int i = nk; i-- > 0;) {
      if (q.group[i]->type == EVQ_BOOL)
        os << "    msg[len++] = (u8) (key[" << i << "] != 0ull);\n    msg[len++] = (u8) ktag[" << i << "];\n";
      else
        os << "    for (int b = 0; b < 8; ++b) msg[len++] = (u8) (key[" << i << "] >> (8 * b));\n    msg[len++] = (u8) ktag[" << i << "];\n";
    }
    if (q.string_keys)   // (a string key's bytes are in the host's dictionary: the tuple travels, wire.cc hashes it)
      os << "    for (u32 b = 0; b < len; ++b) E.out_sha[out_row * " << wire_key_stride(q) << " + b] = msg[b];\n  }\n";
    else
      os << "    evq_sha1(msg, len, E.out_sha + out_row * 20);\n  }\n";
    for (int s = 0; s < nstate; ++s) {
      // count_distinct: the saved state is the SET (aggregate.cc:110-116), which the host reads from the (group id, value)
      // table - the word carries the group's id there (its count is the size of the set)
      const bool is_distinct = std::find(q.distinct_word.begin(), q.distinct_word.end(), s) != q.distinct_word.end();
      const std::string v = !is_distinct ? "st[" + std::to_string(s) + "]" : (shape.tier == 1 || shape.dense_global) ? "slot" : "(u64) sp";
      os << "  E.out_state[out_row * " << nstate << " + " << s << "] = " << v << ";\n";
    }
  }
  os << "  (void) err;\n}\n";
  os << "extern \"C\" __global__ void evq_emit(const __grid_constant__ EvqEmitParams E) {\n";
  os << "  const u64 slot = (u64) blockIdx.x * blockDim.x + threadIdx.x;\n  if (slot >= E.slots) return;\n";
  os << "  evq_emit_group(E, slot, E.dense_state);\n}\n";

  // ---- partitioned aggregation, pass 2: the records of ONE partition into the group table.  All their home slots lie in
  // one slice of the table (the partition is the top bits of the slot index), which stays L2-resident while the partition
  // is processed: probes hit L2 and the aggregate updates are L2 atomics - the only HBM traffic is the sequential read of
  // the records.  Two records per thread in flight (both first probes are issued before either is resolved).
  if (shape.tier == 2 && shape.part_bits > 0) {
    os << "struct EvqAggParams { EvqHashTable ht; const u64* part_buf; const u32* part_cursor; u64 part_cap; u32 nparts; u32 nseg; u32* status; u64* counters; u32* bar; u32 window; u32 pad; };\n";
    os << "extern \"C\" __global__ void __launch_bounds__(256) evq_agg_part(const __grid_constant__ EvqAggParams A) {\n";
    os << "  u32 err = 0;\n";
    // ONE (cooperative: all CTAs resident) launch walks the partitions in order, all CTAs striding over a partition's
    // records together.  The CTAs stay within `window` partitions of each other - before partition p
    // a CTA waits until every CTA is done with partition p - window (a counter every CTA bumps once per partition) - so
    // that at most window + 1 table slices are being worked on: they stay L2-resident, and nobody idles at a barrier.
    os << "  for (u32 part = 0; part < A.nparts; ++part) {\n";
    os << "    if (part >= A.window) {\n      if (threadIdx.x == 0) {\n        const u32 need = (part - A.window + 1u) * gridDim.x;\n        u32 seen;\n"
          "        do { asm volatile(\"ld.acquire.gpu.global.u32 %0, [%1];\" : \"=r\"(seen) : \"l\"(A.bar) : \"memory\"); if (seen < need) __nanosleep(200); } while (seen < need);\n"
          "      }\n      __syncthreads();\n    }\n";
    os << "    {\n";
    os << "      const u32 cur = A.part_cursor[part];\n      const u32 n = cur < A.part_cap ? cur : (u32) A.part_cap;\n";
    os << "      const u64* base = A.part_buf + (u64) part * A.part_cap * EVQ_NREC;\n";
    os << "      const u32 stride = gridDim.x * 256u;\n";
    os << "      for (u32 i = blockIdx.x * 256u + threadIdx.x; i < n; i += 4u * stride) {\n";
    os << "        EvqRow row[4];\n        u64 key[4][EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];\n        u64 fpv[4], slot[4], w0[4], w1[4];\n        bool have[4];\n";
    os << "#pragma unroll\n        for (int j = 0; j < 4; ++j) {\n          have[j] = i + j * stride < n;\n          fpv[j] = slot[j] = w0[j] = w1[j] = 0;\n"
          "          if (have[j]) evq_row_load(base + (u64) (i + j * stride) * EVQ_NREC, row[j]);\n        }\n";
    os << "#pragma unroll\n        for (int j = 0; j < 4; ++j) {\n          if (have[j]) {\n            u32 ktag[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];\n"
          "            evq_keys(row[j], key[j], ktag, err);\n            evq_ht_hash<EVQ_NKEYS>(A.ht, key[j], ktag, fpv[j], slot[j]);\n"
          "            evq_ht_prefetch(A.ht, slot[j], w0[j], w1[j]);\n          }\n        }\n";
    os << "#pragma unroll\n        for (int j = 0; j < 4; ++j) {\n          if (have[j]) {\n"
          "            u64* sp = evq_ht_upsert_from<EVQ_NKEYS>(A.ht, key[j], fpv[j], slot[j], w0[j], w1[j], (u64*) 0);\n"
          "            if (!sp) err |= EVQ_ERR_TABLE_FULL;\n            else evq_accumulate_global(row[j], sp + 1 + EVQ_NKEYS, err);\n          }\n        }\n      }\n    }\n"
          "    __syncthreads();\n    if (threadIdx.x == 0) { __threadfence(); atomicAdd(A.bar, 1u); }\n  }\n";
    os << "  if (err) atomicOr(A.status, err);\n}\n";
  }

  // ---- partitioned aggregation with shared-memory table slices (the default form of the hash tier beyond L2).
  // The L2-resident form above is bound by the rate of scattered L2 requests (a probe + one atomic per state word and row,
  // profiles/README.md round 2).  Here the records are partitioned once more (evq_repart: every first-level partition into
  // 2^sub_bits sub-partitions by the next bits of the home slot), until the table slice of one sub-partition
  // (slice_slots slots) fits shared memory; evq_agg_smem then takes one sub-partition at a time: slice in (coalesced),
  // records aggregated with shared-memory atomics (probing wraps inside the slice), slice out.  Per row no request
  // leaves the SM except the sequential read of its record.
  if (shape.tier == 2 && shape.part_bits > 0 && shape.slice_slots > 0) {
    os << "#define EVQ_SLOT_WORDS " << 1 + nk + nstate << "\n#define EVQ_RP_TILE 2048\n";
    os << "struct EvqRepartParams { EvqHashTable ht; const u64* in; const u32* in_cursor; u64 in_cap; u64* out; u32* out_cursor; u64 out_cap; "
          "u32* status; u32 nparts; u32 sub_bits; u32 sub_shift; u32 pad; };\n";
    os << R"EVQ(extern "C" __global__ void __launch_bounds__(256) evq_repart(const __grid_constant__ EvqRepartParams R) {
  extern __shared__ __align__(128) u8 evq_smem[];
  u64* prec = (u64*) evq_smem;                          // the tile's records ordered by sub-partition
  u32* hist = (u32*) (prec + EVQ_RP_TILE * EVQ_NREC);   // [256] records of the tile per sub-partition
  u32* scan = hist + 256;                               // [256] ... before the sub-partition
  u32* base = scan + 256;                               // [256] where the run goes (claimed from the global cursors)
  u32* wsum = base + 256;                               // [8] warp totals of the scan
  u32* tpre = wsum + 8;                                 // [257] tiles before every first-level partition
  u8* ppart = (u8*) (tpre + 260);                       // [EVQ_RP_TILE] the sub-partition of every staged record
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const u32 hist_sa = evq_smem_u32(hist);
  const u32 nsub = 1u << R.sub_bits;
  u32 err = 0;
  hist[tid] = 0u;
  if (tid == 0) {
    u32 t = 0;
    for (u32 p = 0; p < R.nparts; ++p) {
      tpre[p] = t;
      const u32 cur = R.in_cursor[p];
      const u32 n = cur < R.in_cap ? cur : (u32) R.in_cap;
      t += (n + EVQ_RP_TILE - 1u) / EVQ_RP_TILE;
    }
    tpre[R.nparts] = t;
  }
  __syncthreads();
  const u32 ntiles = tpre[R.nparts];
  for (u32 g = blockIdx.x; g < ntiles; g += gridDim.x) {
    u32 lo = 0, hi = R.nparts;   // the partition of tile g: last p with tpre[p] <= g
    while (hi - lo > 1u) { const u32 mid = (lo + hi) >> 1; if (tpre[mid] <= g) lo = mid; else hi = mid; }
    const u32 part = lo;
    const u32 cur = R.in_cursor[part];
    const u32 n = cur < R.in_cap ? cur : (u32) R.in_cap;
    const u64* src = R.in + (u64) part * R.in_cap * EVQ_NREC;
    const u32 t0 = (g - tpre[part]) * EVQ_RP_TILE;
    EvqRow row[8];
    u32 pr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const u32 i = t0 + (u32) j * 256u + tid;
      pr[j] = ~0u;
      if (i < n) evq_row_load(src + (u64) i * EVQ_NREC, row[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const u32 i = t0 + (u32) j * 256u + tid;
      if (i < n) {
        u64 key[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
        u32 ktag[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
        u64 fpv, slot;
        evq_keys(row[j], key, ktag, err);
        evq_ht_hash<EVQ_NKEYS>(R.ht, key, ktag, fpv, slot);
        const u32 sub = (u32) (slot >> R.sub_shift) & (nsub - 1u);
        u32 rank;
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(rank) : "r"(hist_sa + 4u * sub));
        pr[j] = sub | (rank << 8);
      }
    }
    __syncthreads();
    const u32 c = hist[tid];
    u32 incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (u32) o) incl += v;
    }
    if (lane == 31u) wsum[warp] = incl;
    __syncthreads();
    u32 before = incl - c, total = 0;
#pragma unroll
    for (u32 w = 0; w < 8; ++w) {
      const u32 x = wsum[w];
      if (w < warp) before += x;
      total += x;
    }
    hist[tid] = 0u;
    scan[tid] = before;
    base[tid] = c ? atomicAdd(R.out_cursor + (((u64) part << R.sub_bits) | tid), c) : 0u;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (pr[j] != ~0u) {
        const u32 sub = pr[j] & 255u;
        const u32 at = scan[sub] + (pr[j] >> 8);
        evq_row_store(row[j], prec + (size_t) at * EVQ_NREC);
        ppart[at] = (u8) sub;
      }
    }
    __syncthreads();
#if EVQ_NREC % 2 == 0
    {
      constexpr u32 U = EVQ_NREC / 2;
      const ulonglong2* st = (const ulonglong2*) prec;
      for (u32 j = tid; j < total * U; j += 256u) {
        const u32 i = j / U, w = j % U;
        const u32 sub = ppart[i];
        const u64 pos = (u64) base[sub] + (i - scan[sub]);
        if (pos < R.out_cap) ((ulonglong2*) (R.out + ((((u64) part << R.sub_bits) | sub) * R.out_cap + pos) * EVQ_NREC))[w] = st[j];
        else err |= EVQ_ERR_PART_FULL;
      }
    }
#else
    for (u32 j = tid; j < total * EVQ_NREC; j += 256u) {
      const u32 i = j / EVQ_NREC, w = j % EVQ_NREC;
      const u32 sub = ppart[i];
      const u64 pos = (u64) base[sub] + (i - scan[sub]);
      if (pos < R.out_cap) R.out[((((u64) part << R.sub_bits) | sub) * R.out_cap + pos) * EVQ_NREC + w] = prec[j];
      else err |= EVQ_ERR_PART_FULL;
    }
#endif
    __syncthreads();
  }
  if (err) atomicOr(R.status, err);
}
struct EvqAggSmemParams { EvqHashTable ht; const u64* buf; const u32* cursor; u64 cap; u32* status; u32 nsub_total; u32 slice_slots; };
#ifndef EVQ_AG_THREADS
#define EVQ_AG_THREADS 1024
#endif
// The table of a sliced plan is COMPACT (stride == EVQ_SLOT_WORDS): a slice is one contiguous block, moved by the TMA -
// one bulk copy in (completion on an mbarrier), one bulk copy out (bulk async-group) per sub-partition, issued by thread 0.
#define EVQ_BULK_CHUNK 32768u
__device__ __forceinline__ void evq_slice_load(const EvqAggSmemParams& A, u32 sp, u8* dst, u64* bar, u32 bytes) {
  // (the bulk store that last read this buffer must be done reading it)
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  evq_mbar_arrive_expect_tx(bar, bytes);
  const u8* g = (const u8*) (A.ht.slots + (u64) sp * A.slice_slots * EVQ_SLOT_WORDS);
  for (u32 off = 0; off < bytes; off += EVQ_BULK_CHUNK)
    evq_bulk_g2s(dst + off, g + off, bytes - off < EVQ_BULK_CHUNK ? bytes - off : EVQ_BULK_CHUNK, bar);
}
__device__ __forceinline__ void evq_slice_store(const EvqAggSmemParams& A, u32 sp, const u8* src, u32 bytes) {
  u8* g = (u8*) (A.ht.slots + (u64) sp * A.slice_slots * EVQ_SLOT_WORDS);
  for (u32 off = 0; off < bytes; off += EVQ_BULK_CHUNK) {
    const u32 nb = bytes - off < EVQ_BULK_CHUNK ? bytes - off : EVQ_BULK_CHUNK;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(g + off), "r"(evq_smem_u32(src + off)), "r"(nb) : "memory");
  }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ u32 evq_next_sub(const EvqAggSmemParams& A, u32 sp) {   // the next sub-partition of this CTA that has records
  while (sp < A.nsub_total && A.cursor[sp] == 0u) sp += gridDim.x;
  return sp;
}
extern "C" __global__ void __launch_bounds__(EVQ_AG_THREADS) evq_agg_smem(const __grid_constant__ EvqAggSmemParams A) {
  extern __shared__ __align__(128) u8 evq_smem[];
  // two slice buffers [slice_slots][EVQ_SLOT_WORDS] (fingerprint, keys, state words): the next sub-partition's slice is
  // copied in while the current one is aggregated; behind them the two mbarriers of the copies
  const u32 buf_bytes = A.slice_slots * (EVQ_SLOT_WORDS * 8u);
  u64* bars = (u64*) (evq_smem + 2u * buf_bytes);
  const u32 tid = threadIdx.x;
  const u32 S = A.slice_slots, smask = A.slice_slots - 1u;
  u32 err = 0;
  if (tid == 0) {
    evq_mbar_init(&bars[0], 1);
    evq_mbar_init(&bars[1], 1);
    evq_mbar_fence_init();
  }
  __syncthreads();
  u32 sp = evq_next_sub(A, blockIdx.x);
  u32 cur_buf = 0, parity = 0;   // bit b of parity: the phase buffer b's next copy completes
  if (tid == 0 && sp < A.nsub_total) evq_slice_load(A, sp, evq_smem, &bars[0], buf_bytes);
  while (sp < A.nsub_total) {
    const u32 nsp = evq_next_sub(A, sp + gridDim.x);
    if (tid == 0 && nsp < A.nsub_total) evq_slice_load(A, nsp, evq_smem + (cur_buf ^ 1u) * buf_bytes, &bars[cur_buf ^ 1u], buf_bytes);
    evq_mbar_wait(&bars[cur_buf], (parity >> cur_buf) & 1u);
    parity ^= 1u << cur_buf;
    const u32 tab_sa = evq_smem_u32(evq_smem) + cur_buf * buf_bytes;
    const u32 cur = A.cursor[sp];
    const u32 n = cur < A.cap ? cur : (u32) A.cap;
    const u64* src = A.buf + (u64) sp * A.cap * EVQ_NREC;
    // (the thread's next three records are on their way while the current one is aggregated)
    EvqRow q0, q1, q2;
    if (tid < n) evq_row_load(src + (u64) tid * EVQ_NREC, q0);
    if (tid + EVQ_AG_THREADS < n) evq_row_load(src + (u64) (tid + EVQ_AG_THREADS) * EVQ_NREC, q1);
    if (tid + 2u * EVQ_AG_THREADS < n) evq_row_load(src + (u64) (tid + 2u * EVQ_AG_THREADS) * EVQ_NREC, q2);
    for (u32 i = tid; i < n; i += EVQ_AG_THREADS) {
      const EvqRow rowj = q0;
      q0 = q1;
      q1 = q2;
      if (i + 3u * EVQ_AG_THREADS < n) evq_row_load(src + (u64) (i + 3u * EVQ_AG_THREADS) * EVQ_NREC, q2);
      u64 key[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
      u32 ktag[EVQ_NKEYS > 0 ? EVQ_NKEYS : 1];
      u64 fpv, slot;
      evq_keys(rowj, key, ktag, err);
      evq_ht_hash<EVQ_NKEYS>(A.ht, key, ktag, fpv, slot);
      u32 l = (u32) slot & smask;
      u32 found = 0;
      bool have = false;
      for (u32 probes = 0; probes < S; ++probes) {
        const u32 sa = tab_sa + l * (EVQ_SLOT_WORDS * 8u);
        u64 c = evq_lds64(sa);
        if (c == 0ull) {
          c = evq_cas_smem(sa, 0ull, fpv | 2ull);
          if (c == 0ull) {   // claimed: keys, then the final fingerprint
#pragma unroll
            for (int k = 0; k < EVQ_NKEYS; ++k) evq_sts64(sa + 8u * (1u + k), key[k]);
            __threadfence_block();
            evq_sts64(sa, fpv);
            found = sa;
            have = true;
            break;
          }
        }
        if ((c | 2ull) == (fpv | 2ull)) {
          while (c & 2ull) c = evq_lds64(sa);   // the claimant is still writing the keys
          bool same = true;
#pragma unroll
          for (int k = 0; k < EVQ_NKEYS; ++k) same = same && evq_lds64(sa + 8u * (1u + k)) == key[k];
          if (same) { found = sa; have = true; break; }
        }
        l = (l + 1u) & smask;
      }
      if (!have) err |= EVQ_ERR_TABLE_FULL;
      else evq_accumulate_smem(rowj, found + 8u * (1u + EVQ_NKEYS), err);
    }
    // the slice goes back: the threads' shared-memory writes become visible to the async proxy, then ONE bulk store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) evq_slice_store(A, sp, evq_smem + cur_buf * buf_bytes, buf_bytes);
    sp = nsp;
    cur_buf ^= 1u;
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (err) atomicOr(A.status, err);
}
)EVQ";
  }

  // ---- the tail of a dense-tier execution: ONE kernel (one CTA) behind the scan launches that
  //   (1) multi-rank: pushes this rank's state words into every peer's exchange buffer over NVLink (P2P stores), raises its
  //       flag there, waits for the peers' flags and combines all ranks' words in rank order (GroupByMergeExpression,
  //       groupby.cc:553-615; deterministic also for double sums) - the all-gather, the merge kernel and the emit kernel in one
  //   (2) emits the result rows (GroupByExpression::nextBatch)
  //   (3) publishes status / counters / row count in the query's result block and re-arms the state (identities, zeroed
  //       counters) for the next execution, which then needs no init kernel and no memsets
  if (shape.tier == 1 && !shape.dense_global) {
    const int slots = shape.g1 > 1 ? shape.g1 : 1;
    const int nwords = slots * nstate;
    os << "struct EvqTailParams { EvqEmitParams E; u64* dense_state; u64* ctl; u32 nranks; u32 rank; u64 epoch; u64* xbuf_local; "
          "u64* flags_local; u64* xbuf_peer[16]; u64* flags_peer[16]; u64* merged_out; };\n";
    os << "#define EVQ_TAIL_WORDS " << nwords << "\n#define EVQ_XSLOT_WORDS 8192\n";
    os << "extern \"C\" __global__ void __launch_bounds__(256, 1) evq_tail(const __grid_constant__ EvqTailParams T) {\n";
    os << "  __shared__ u64 m[EVQ_TAIL_WORDS];\n  __shared__ u64 nrows;\n  __shared__ u32 timed_out;\n";
    os << "  const u32 tid = threadIdx.x;\n  if (tid == 0) { nrows = 0; timed_out = 0; }\n";
    os << "  const u64 par = T.epoch & 1ull;\n";
    os << "  if (T.nranks > 1) {\n";
    os << "    for (u32 i = tid; i < EVQ_TAIL_WORDS; i += 256) {\n      const u64 v = T.dense_state[i];\n"
          "      for (u32 r = 0; r < T.nranks; ++r)\n        if (r != T.rank) T.xbuf_peer[r][(par * 16 + T.rank) * EVQ_XSLOT_WORDS + i] = v;\n    }\n";
    os << "    __threadfence_system();\n    __syncthreads();\n";
    os << "    if (tid < T.nranks && tid != T.rank) {\n"
          "      asm volatile(\"st.release.sys.global.u64 [%0], %1;\" :: \"l\"(T.flags_peer[tid] + T.rank), \"l\"(T.epoch) : \"memory\");\n"
          "      const long long t0 = clock64();\n      u64 seen = 0;\n"
          "      for (;;) {\n        asm volatile(\"ld.acquire.sys.global.u64 %0, [%1];\" : \"=l\"(seen) : \"l\"(T.flags_local + tid) : \"memory\");\n"
          "        if (seen >= T.epoch) break;\n"
          "        if (clock64() - t0 > 40000000000ll) { timed_out = 1; break; }\n        __nanosleep(64);\n      }\n    }\n";
    os << "    __syncthreads();\n  }\n";
    // combine in rank order
    os << "  for (u32 i = tid; i < EVQ_TAIL_WORDS; i += 256) {\n    const u32 w = i % " << nstate << "u;\n    u64 acc = 0;\n    switch (w) {\n";
    for (int w = 0; w < nstate; ++w) {
      const int op = q.state_ops[w];
      os << "      case " << w << ": {\n";
      if (op == OP_FIRST_ORD || op == OP_FIRST_VAL) {
        const int d = op == OP_FIRST_ORD ? 0 : 1;
        os << "        const u32 oi = i - " << d << "u;\n        u64 best = ~0ull, val = 0;\n"
              "        for (u32 r = 0; r < T.nranks; ++r) {\n"
              "          const u64* src = r == T.rank ? T.dense_state : (const u64*) (T.xbuf_local + (par * 16 + r) * EVQ_XSLOT_WORDS);\n"
              "          const u64 o = __ldcv(src + oi);\n          if (o < best) { best = o; val = __ldcv(src + oi + 1); }\n        }\n"
              "        acc = " << (d ? "val" : "best") << ";\n";
      } else {
        const int carry_of = q.state_carry_of[w];
        os << "        acc = evq_state_identity<" << op << ">();\n";
        if (carry_of >= 0) os << "        u64 lo = 0;\n";
        os << "        for (u32 r = 0; r < T.nranks; ++r) {\n"
              "          const u64* src = r == T.rank ? T.dense_state : (const u64*) (T.xbuf_local + (par * 16 + r) * EVQ_XSLOT_WORDS);\n"
              "          acc = evq_state_combine<" << op << ">(acc, __ldcv(src + i));\n";
        if (carry_of >= 0)   // add the wraps of re-summing the partner word in the same rank order
          os << "          { const u64 n = __ldcv(src + i - " << w - carry_of << "u); lo += n; acc += lo < n ? 1ull : 0ull; }\n";
        os << "        }\n";
      }
      os << "        break;\n      }\n";
    }
    os << "    }\n    m[i] = acc;\n  }\n  __syncthreads();\n";
    os << "  {\n    EvqEmitParams E = T.E;\n    E.out_count = &nrows;\n    for (u64 slot = tid; slot < E.slots; slot += 256) evq_emit_group(E, slot, m);\n  }\n";
    os << "  __syncthreads();\n";
    // publish + re-arm
    os << "  for (u32 i = tid; i < EVQ_TAIL_WORDS; i += 256) {\n    if (T.merged_out) T.merged_out[i] = m[i];\n    switch (i % " << nstate << "u) {\n";
    for (int w = 0; w < nstate; ++w) os << "      case " << w << ": T.dense_state[i] = evq_state_identity<" << q.state_ops[w] << ">(); break;\n";
    os << "    }\n  }\n";
    os << "  if (tid == 0) {\n    u64* c = T.ctl;\n    if (timed_out) ((u32*) c)[0] |= EVQ_ERR_PEER_TIMEOUT;\n"
          "    for (int j = 0; j < 6; ++j) { c[8 + j] = c[j]; c[j] = 0ull; }\n    c[8 + 6] = nrows;\n    c[8 + 7] += 1ull;\n  }\n";
    os << "}\n";
  }
  return os.str();
}

std::string generate_coordinator_source(const evqgpu_query& q) {
  KernelShape shape;
  shape.tier = 2;
  std::ostringstream os;
  os << "// generated by eventql_b200 csrc/codegen.cc - emit kernel of a coordinator (GroupByMergeExpression) query\n";
  os << "#define EVQ_NCONS 256\n#define EVQ_NKEYS " << q.group.size() << "\n";
  os << kSrcAbi << "\n" << kSrcPrelude << "\n" << gen_group_kernels(q, shape);
  return os.str();
}

// 16-byte chunks one tile of the longest variable-length column can span (+2: unaligned start, rounding)
int gen_chunks(const KernelShape& shape) {
  uint32_t L = 1;
  for (const auto& c : shape.cols)
    if (c.used && c.gen_slot >= 0) L = std::max(L, c.leb_len);
  return (int) ((L * EVQ_TILE_ROWS + 15) / 16 + 2);
}

// the packed record of the partitioned hash tier: the columns the GROUP BY expressions and the aggregate arguments read, each
// in the bits its largest value needs (+ a tag bit for a NULL-able column), fields never straddling a word
RecordLayout record_layout(const KernelShape& shape) {
  RecordLayout L;
  int word = 0, used = 0;
  auto place = [&](int col, int bits, bool is_tag) {
    if (used + bits > 64) { ++word; used = 0; }
    L.fields.push_back({col, word, used, bits, is_tag});
    used += bits;
  };
  for (int c : shape.rec_cols) {
    const ColSig& cs = shape.cols[c];
    int bits = 64;
    const bool wide = cs.sql_type == EVQ_FLOAT64 || cs.sql_type == EVQ_INT64 || getenv("EVQGPU_NO_PACKED_RECORDS");
    if (!wide) {
      bits = 1;
      while (bits < 64 && (cs.vmax >> bits) != 0) ++bits;
    }
    place(c, bits, false);
    if (cs.nullable) place(c, 1, true);
  }
  L.nwords = (size_t) word + 1;
  return L;
}

// pass 1 of the partitioned hash tier gathers a tile's records in one shared-memory bin per partition: twice the expected
// records of a 1024-row tile, within 8 .. 64, halved while the bins exceed 48 KB (wide records)
int part_bin_records(int part_bits, size_t nrec) {
  int bin = std::max(8, std::min(64, 2048 >> part_bits));
  while (((size_t) bin << part_bits) * nrec * 8 > 48 * 1024 && bin > 8) bin /= 2;
  return bin;
}

std::string generate_source(const evqgpu_query& q, const KernelShape& shape_in) {
  KernelShape shape = shape_in;
  for (int col : q.narrow_col)
    if (col >= 0) shape.cols[col].packed = true;
  for (const auto& b : q.plane_b)
    if (b.presence_col >= 0) shape.cols[b.presence_col].presence = true;
  std::ostringstream os;
  os << "// generated by eventql_b200 csrc/codegen.cc - one fused scan kernel per (plan, column layout)\n";
  os << "#define EVQ_NCONS " << shape.ncons << "\n#define EVQ_NSTAGES " << shape.nstages << "\n#define EVQ_NSTREAMS "
     << shape.nstreams << "\n#define EVQ_TIER " << shape.tier << "\n#define EVQ_G1 " << shape.g1 << "\n#define EVQ_NSTATE "
     << std::max(1, q.nstate_smem) << "\n#define EVQ_NKEYS " << q.group.size() << "\n#define EVQ_NLEB "
     << shape.nleb << "\n#define EVQ_NNULL " << shape.nnull << "\n#define EVQ_HAS_PREP "
     << ((shape.nleb > 0 || shape.nnull > 0) ? 1 : 0) << "\n#define EVQ_MIN_CTAS " << shape.min_ctas << "\n#define EVQ_NGEN "
     << shape.ngen << "\n#define EVQ_GEN_CHUNKS " << gen_chunks(shape) << "\n#define EVQ_NNV " << shape.nnv << "\n#define EVQ_NNARROW " << q.nnarrow << "\n#define EVQ_NG " << std::max(1, q.plane_groups) << "\n#define EVQ_NSTATE_SMEM " << q.nstate_smem << "\n#define EVQ_NSTATE_ALL "
     << std::max<size_t>(1, q.state_ops.size()) << "\n";
  if (shape.part_bits > 0) {
    const size_t nrec = record_layout(shape).nwords;
    os << "#define EVQ_PARTITION 1\n#define EVQ_MAX_PARTS " << (1 << shape.part_bits) << "\n#define EVQ_NREC " << nrec
       << "\n#define EVQ_PART_BIN " << part_bin_records(shape.part_bits, nrec) << "\n";
    if (shape.slice_slots > 0) {
      os << "#define EVQ_SMEM_SLICES 1\n";
      if (const char* e = getenv("EVQGPU_AGG_THREADS")) os << "#define EVQ_AG_THREADS " << atoi(e) << "\n";   // (sweep aids, scripts/c4_step.sh)
    }
  }
  if (getenv("EVQGPU_DRYRUN")) os << "#define EVQ_DRYRUN 1\n";
  if (shape.fast) os << "#define EVQ_KT " << shape.kt << "\n";
  if (shape.filter_stream >= 0) os << "#define EVQ_FILTER_STREAM " << shape.filter_stream << "\n";
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
