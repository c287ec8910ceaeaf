// query.cc - evqgpu_query: plan intake, kernel specialisation, launches, result fetch.
//
// Operator mapping (reference, src/eventql/sql/):
//   evqgpu_query_create   FastCSTableScan::execute (CSTableScan.cc:726-755) + Compiler::compile's split of every
//                         select item into an accumulate program and a get program (runtime/compiler.cc:50-104)
//   evqgpu_query_execute  the nextBatch loop of FastCSTableScan (CSTableScan.cc:757-858) drained by
//                         GroupByExpression::execute (statements/select/groupby.cc:69-185)
//   evqgpu_query_fetch    GroupByExpression::nextBatch (groupby.cc:187-220) / the scan's own output batches
#include "query.h"
#include <cub/device/device_scan.cuh>
#include <string.h>
#include <algorithm>
#include <cmath>
#include <functional>

using namespace evq;

namespace evq {

static uint64_t next_pow2(uint64_t v) {
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

static void launch(evqgpu_ctx* ctx, cudaKernel_t k, dim3 grid, dim3 block, size_t smem, void** args) {
  EVQ_CUDA(cudaLaunchKernel((const void*) k, grid, block, args, smem, ctx->stream));
  ctx->kernel_launches++;
}

static void ensure(DevBuf& b, uint64_t bytes) {
  if (b.bytes < bytes) b.alloc(bytes);
}

// ---- binding of the plan's input columns to one table -------------------------------------------------------------
struct Binding {
  std::vector<int> col_index;          // plan input column -> table column (or -1 when unused)
  std::vector<const Column*> cols;     // ... and the column the kernels read: the table's own, or the code column of a string column
};

static Binding bind_table(const evqgpu_query& q, evqgpu_table* t) {
  Binding b;
  b.col_index.assign(q.input_columns.size(), -1);
  b.cols.assign(q.input_columns.size(), nullptr);
  for (size_t i = 0; i < q.input_columns.size(); ++i) {
    if (!q.col_used[i]) continue;
    const int ci = t->find(q.input_columns[i].c_str());
    if (ci < 0) fail(EVQGPU_ERR_ARG, "column not found: %s", q.input_columns[i].c_str());
    b.col_index[i] = ci;
    if (q.col_is_string[i]) {
      if (!t->cols[ci].is_string)
        fail(EVQGPU_ERR_ARG, "column '%s' is used as a string but is not a flat string column", q.input_columns[i].c_str());
      b.cols[i] = ensure_code_column(t, t->cols[ci]);
      continue;
    }
    if (i < q.col_pred.size() && q.col_pred[i] >= 0) {   // the verdict column of a string predicate over this string column
      if (!t->cols[ci].is_string)
        fail(EVQGPU_ERR_ARG, "column '%s' is used as a string but is not a flat string column", q.input_columns[i].c_str());
      b.cols[i] = ensure_pred_column(t, t->cols[ci], q.string_preds[q.col_pred[i]]);
      continue;
    }
    if (!t->cols[ci].scannable)
      fail(EVQGPU_ERR_UNSUPPORTED, "column '%s' (logical type %u, encoding %u, rlevel_max %u) is outside the flat numeric scan path",
           q.input_columns[i].c_str(), t->cols[ci].meta.logical_type, t->cols[ci].meta.encoding, t->cols[ci].meta.rlevel_max);
    if (!t->cols[ci].loaded) table_load_column(t, t->cols[ci]);
    b.cols[i] = &t->cols[ci];
  }
  return b;
}

// The kernels are specialised on the value range of the columns (narrow arithmetic, comparisons that the range decides,
// range checks that cannot fail).  Small bounds are kept exact; large ones are coarsened to powers of two so that tables
// whose extremes differ a little share one kernel.
static uint64_t stat_ceil(uint64_t vmax) {
  if (vmax < 256) return vmax;
  const int b = 64 - __builtin_clzll(vmax);
  return b >= 64 ? ~0ull : (1ull << b) - 1;
}
static uint64_t stat_floor(uint64_t vmin) {
  if (vmin < 256) return vmin;
  return 1ull << (63 - __builtin_clzll(vmin));
}

static KernelShape shape_for(const evqgpu_query& q, evqgpu_table* t, const Binding& b) {
  KernelShape s;
  s.cols.resize(q.input_columns.size());
  s.fast = true;
  for (size_t i = 0; i < q.input_columns.size(); ++i) {
    if (b.col_index[i] < 0) continue;
    const Column& c = *b.cols[i];
    ColSig& cs = s.cols[i];
    cs.used = true;
    cs.sql_type = c.sql_type;
    cs.kind = c.data_kind;
    cs.nullable = c.meta.dlevel_max > 0;
    cs.dmax = c.meta.dlevel_max;
    cs.bits = c.value_bits;
    cs.vmax = stat_ceil(c.value_max);
    cs.vmin = stat_floor(c.value_min);
    cs.vmin_present = stat_floor(c.value_min_present);
    cs.leb_len = c.leb_max_len;
    cs.leb_uniform = c.data_kind == EVQ_KIND_LEB128 && c.leb_uniform && !getenv("EVQGPU_NO_UNIFORM");
    cs.data_stream = s.nstreams++;
    if (cs.nullable) {
      cs.level_stream = s.nstreams++;
      cs.null_slot = s.nnull++;
      // flat optional columns (one definition level bit per row) also run the fast layout
      if (c.meta.dlevel_max != 1 || c.level_bits != 1 || getenv("EVQGPU_NO_FAST_NULL")) s.fast = false;
    }
    if (cs.kind == EVQ_KIND_LEB128) cs.leb_slot = s.nleb++;
  }
  if (s.nstreams > EVQ_MAX_STREAMS)
    fail(EVQGPU_ERR_UNSUPPORTED, "query reads %d column streams; the scan kernel stages at most %d", s.nstreams, EVQ_MAX_STREAMS);
  return s;
}

// the value-range statistics only widen a kernel, they never make it wrong: take the widest over the partitions
static void widen_shape(KernelShape& s, const KernelShape& o) {
  for (size_t i = 0; i < s.cols.size(); ++i) {
    s.cols[i].bits = std::max(s.cols[i].bits, o.cols[i].bits);
    s.cols[i].vmax = std::max(s.cols[i].vmax, o.cols[i].vmax);
    s.cols[i].vmin = std::min(s.cols[i].vmin, o.cols[i].vmin);
    s.cols[i].vmin_present = std::min(s.cols[i].vmin_present, o.cols[i].vmin_present);
    s.cols[i].leb_uniform = s.cols[i].leb_uniform && o.cols[i].leb_uniform && s.cols[i].leb_len == o.cols[i].leb_len;
    s.cols[i].leb_len = std::max(s.cols[i].leb_len, o.cols[i].leb_len);
  }
}

static void finish_shape(KernelShape& s, bool have_subidx) {
  if (getenv("EVQGPU_NO_FAST")) s.fast = false;
  s.use_subidx = s.fast && have_subidx && !getenv("EVQGPU_NO_SUBIDX");
  if (s.nnull > 0 && !s.use_subidx) s.fast = false;   // the in-kernel boundary search covers required columns only
  s.ngen = 0;
  for (auto& c : s.cols) {
    c.gen_slot = -1;
    c.sub_stream = -1;
    if (!(s.fast && c.used && c.kind == EVQ_KIND_LEB128 && c.leb_len >= 2) || c.leb_uniform) continue;
    if (s.use_subidx) c.sub_stream = s.nstreams++;
    else c.gen_slot = s.ngen++;
  }
  s.nnv = 0;
  for (auto& c : s.cols) {
    c.nv_slot = -1;
    if (s.fast && c.used && c.nullable && c.kind == EVQ_KIND_LEB128 && c.leb_len >= 2 && c.leb_len <= 4 && c.sub_stream >= 0 &&
        !getenv("EVQGPU_NO_STAGED_NULLS"))
      c.nv_slot = s.nnv++;
  }
  if (s.nstreams > EVQ_MAX_STREAMS)
    fail(EVQGPU_ERR_UNSUPPORTED, "query reads %d column streams; the scan kernel stages at most %d", s.nstreams, EVQ_MAX_STREAMS);
}

// layout (not the value ranges) must agree between the partitions of one query
static std::string shape_key(const KernelShape& s) {
  std::string k;
  for (const auto& c : s.cols) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%d:%u:%u:%d:%u|", (int) c.used, c.sql_type, c.kind, (int) c.nullable, c.dmax);
    k += buf;
  }
  return k;
}

// per-table stage layout: where every stream lands inside one pipeline stage
struct StageLayout {
  uint32_t smem_off[EVQ_MAX_STREAMS] = {0};
  uint32_t smem_cap[EVQ_MAX_STREAMS] = {0};
  uint32_t stage_bytes = 0;
};

static StageLayout stage_layout(const evqgpu_query& q, evqgpu_table* t, const Binding& b, const KernelShape& s) {
  StageLayout L;
  L = StageLayout();
  uint32_t off = 0;
  for (size_t i = 0; i < s.cols.size(); ++i) {
    const ColSig& cs = s.cols[i];
    if (!cs.used) continue;
    const Column& c = *b.cols[i];
    auto place = [&](int stream, uint32_t cap) {
      // regions are packed at TMA granularity (16 bytes); decoders may read a few values past a payload (short last
      // tile), which lands in the next region or in the tail pad of the stage
      cap = (uint32_t) round_up(cap, 16);
      L.smem_off[stream] = off;
      L.smem_cap[stream] = cap;
      off += cap;
    };
    // (kt consecutive tiles are contiguous in a required column's stream: at most kt times the largest tile)
    place(cs.data_stream, c.data_tile_cap * (uint32_t) s.kt);
    if (cs.nullable) place(cs.level_stream, c.level_tile_cap * (uint32_t) s.kt);
    if (cs.sub_stream >= 0) place(cs.sub_stream, EVQ_SUB_ENTRIES * 2 * (uint32_t) s.kt);
  }
  if (s.filter_stream >= 0) {
    L.smem_off[s.filter_stream] = off;
    L.smem_cap[s.filter_stream] = (EVQ_TILE_ROWS / 8) * (uint32_t) s.kt;
    off += L.smem_cap[s.filter_stream];
  }
  L.stage_bytes = (uint32_t) round_up(off + 128, 128);
  (void) q;
  return L;
}

static size_t scratch_bytes(const KernelShape& s) {
  const size_t nwarps = s.ncons / 32;
  if (s.fast) {   // EvqFastScratch
    const size_t ngen = std::max(1, s.ngen);
    const size_t parts = s.part_bits > 0 ? ((size_t) 1 << s.part_bits) : 0;
    size_t staging = 0;   // partitioned aggregation: the tile's records ordered by partition + their partition bytes
    if (s.part_bits > 0) {
      const size_t nrec = record_layout(s).nwords;
      staging = ((size_t) part_bin_records(s.part_bits, nrec) << s.part_bits) * nrec * 8 + 32;
    }
    return round_up(4 * ngen * nwarps + 4 * ngen * (size_t) gen_chunks(s) + 4 * nwarps + 8 * std::max(1, s.nnull) * nwarps +
                    (size_t) s.nnv * 2 * (EVQ_TILE_ROWS + 8) * 4 + 16 + parts * (2 * 4 + 2 * 8) + 16 + staging, 128) + 128;
  }
  const size_t one = 4 * std::max(1, s.nleb) * nwarps + 4 * std::max(1, s.nnull) * (EVQ_TILE_ROWS / 32) +
                     2 * std::max(1, s.nleb) * EVQ_TILE_ROWS + 4 * nwarps;
  return round_up(2 * round_up(one, 8), 128) + 128;
}

static size_t header_bytes(const KernelShape& s) {
  const size_t raw = 8 * 4 + 8 * 4 + (s.fast ? 0 : 4 * 4) + 16 * 4 * (s.fast ? s.kt : 1) * std::max(1, s.nstreams);
  return round_up(raw, 128);
}

static size_t acc_bytes(const evqgpu_query& q, const KernelShape& s) {
  if (s.tier != 1 || s.g1 <= 1) return 0;
  return (size_t) s.g1 * q.nstate_smem * s.ncons * 8;
}

}  // namespace evq

// ---- plan intake -----------------------------------------------------------------------------------------------------

static bool is_function_of_keys(const evqgpu_query& q, const Expr* e) {
  // a non-aggregate select item that is a function of the GROUP BY key is evaluated from the stored key at emit time;
  // anything else takes the value of the group's first row (groupby.cc:161-172).  Decided by trying to generate the emit code.
  CodegenEnv env;
  env.col_value.assign(q.input_columns.size(), "");
  env.col_tag.assign(q.input_columns.size(), "");
  for (size_t i = 0; i < q.group.size(); ++i) env.subst.push_back({q.group[i]->signature(), {"k", "t"}});
  try {
    (void) gen_expr(e, env);
  } catch (const Error& err) {
    if (err.status == EVQGPU_ERR_UNSUPPORTED) return false;
    throw;
  }
  return true;
}

namespace evq {

// String support (strings.cu): eq / neq between string columns and string literals, and bare string columns as GROUP BY
// expressions / select items, are rewritten to the columns' dictionary codes (evqgpu_ctx::string_codes; NULL reads as code
// 0 = "", which is what eq_string sees for a NULL, boolean.cc:235-257).  Everything else typed STRING is refused.
static void lower_string_leaf(evqgpu_query* q, Expr* a) {
  if (a->op == EVQ_X_INPUT) {
    if (a->col >= q->col_is_string.size()) fail(EVQGPU_ERR_ARG, "expression references input column %u of %zu", a->col, q->col_is_string.size());
    q->col_is_string[a->col] = true;
  } else if (a->op == EVQ_X_LITERAL) {
    if (!q->ctx) fail(EVQGPU_ERR_UNSUPPORTED, "string literals need a context (dictionary codes)");
    a->imm = string_code(q->ctx, a->str);
    q->string_literals.push_back({a, a->str});   // (a multi-rank dictionary synchronisation may renumber the code)
    a->str.clear();
  } else {
    fail(EVQGPU_ERR_UNSUPPORTED, "string expressions other than columns and literals are outside the device path");
  }
  a->type = EVQ_UINT64;
}

// fn(string column, string literal) in either order, fn an ordering comparison or startswith / endswith: the node
// becomes a BOOL input column - the predicate's verdict column, bound per table by ensure_pred_column (strings.cu) - negated
// when the predicate holds for "": a NULL row reads 0 and must get what the reference computes for it, predicate("")
// (the tag is dropped, boolean.cc:441-452)
static bool lower_string_predicate(evqgpu_query* q, Expr* e) {
  if (e->op != EVQ_X_CALL || e->args.size() != 2) return false;
  const Fn fn = e->info().fn;
  if (!(fn == Fn::LT || fn == Fn::LTE || fn == Fn::GT || fn == Fn::GTE || fn == Fn::STARTSWITH || fn == Fn::ENDSWITH)) return false;
  if (e->info().args[0] != EVQ_STRING) return false;
  Expr* a = e->args[0].get();
  Expr* b = e->args[1].get();
  const bool column_first = a->op == EVQ_X_INPUT && b->op == EVQ_X_LITERAL;
  if (!column_first && !(a->op == EVQ_X_LITERAL && b->op == EVQ_X_INPUT))
    fail(EVQGPU_ERR_UNSUPPORTED, "string comparisons run on the device between a string column and a string literal");
  const Expr* col = column_first ? a : b;
  const Expr* lit = column_first ? b : a;
  if (col->col >= q->input_columns.size()) fail(EVQGPU_ERR_ARG, "expression references input column %u of %zu", col->col, q->input_columns.size());
  StringPredicate p;
  p.fn = (int) fn;
  p.column_first = column_first;
  p.literal = lit->str;
  p.invert = string_predicate_eval(p, std::string());
  // a pseudo input column (same table column name, so the binding finds the string column)
  q->string_preds.push_back(p);
  q->input_columns.push_back(q->input_columns[col->col]);
  q->col_is_string.push_back(false);
  q->col_pred.resize(q->input_columns.size(), -1);
  q->col_pred.back() = (int) q->string_preds.size() - 1;
  const uint32_t idx = (uint32_t) q->input_columns.size() - 1;
  // e := [neg] input(idx) : BOOL
  std::unique_ptr<Expr> in(new Expr());
  in->op = EVQ_X_INPUT;
  in->type = EVQ_BOOL;
  in->col = idx;
  e->args.clear();
  if (p.invert) {
    e->fn = function_lookup("neg#bool/bool;");
    e->args.push_back(std::move(in));
  } else {
    e->op = EVQ_X_INPUT;
    e->col = idx;
    e->fn = 0;
  }
  e->type = EVQ_BOOL;
  return true;
}

static void lower_strings(evqgpu_query* q, Expr* e) {
  if (!e) return;
  if (lower_string_predicate(q, e)) return;
  if (e->op == EVQ_X_CALL && (e->info().fn == Fn::EQ || e->info().fn == Fn::NEQ) && e->info().args[0] == EVQ_STRING) {
    lower_string_leaf(q, e->args[0].get());
    lower_string_leaf(q, e->args[1].get());
    e->fn = function_lookup(e->info().fn == Fn::EQ ? "eq#bool/uint64;uint64;" : "neq#bool/uint64;uint64;");
    return;
  }
  if (e->op == EVQ_X_CALL && e->info().fn == Fn::DATE_TRUNC) {   // its window literal stays a string
    lower_strings(q, e->args[1].get());
    return;
  }
  if (e->type == EVQ_STRING)
    fail(EVQGPU_ERR_UNSUPPORTED, "string values are outside the device path except in eq / neq with columns and literals and as bare GROUP BY / select columns");
  for (auto& a : e->args) lower_strings(q, a.get());
}

// a bare string column as a GROUP BY expression or select item
static bool lower_bare_string_column(evqgpu_query* q, Expr* e) {
  if (!e || e->op != EVQ_X_INPUT || e->type != EVQ_STRING) return false;
  lower_string_leaf(q, e);
  return true;
}

void query_intake(evqgpu_query* q, const evqgpu_query_desc* desc) {
  if (desc->struct_size != sizeof(evqgpu_query_desc)) fail(EVQGPU_ERR_ARG, "evqgpu_query_desc: struct_size mismatch");
  q->flags = desc->flags;
  q->expected_groups = desc->expected_groups;
  for (uint32_t i = 0; i < desc->num_input_columns; ++i) q->input_columns.push_back(desc->input_columns[i]);
  q->col_is_string.assign(q->input_columns.size(), false);
  q->where = parse_program(desc->where);
  lower_strings(q, q->where.get());
  if (q->where && q->where->type != EVQ_BOOL) fail(EVQGPU_ERR_ARG, "WHERE expression must be of type bool");
  if (q->where && find_aggregate(q->where.get())) fail(EVQGPU_ERR_ARG, "aggregate call in WHERE");
  if (desc->num_group > EVQ_MAX_KEYS) fail(EVQGPU_ERR_UNSUPPORTED, "at most %d GROUP BY expressions", EVQ_MAX_KEYS);
  if (desc->num_select == 0 || desc->num_select > EVQ_MAX_STREAMS)
    fail(EVQGPU_ERR_UNSUPPORTED, "select list must have 1..%d items", EVQ_MAX_STREAMS);
  for (uint32_t i = 0; i < desc->num_group; ++i) {
    ExprPtr g = parse_program(desc->group[i]);
    if (!g) fail(EVQGPU_ERR_ARG, "empty GROUP BY expression");
    if (find_aggregate(g.get())) fail(EVQGPU_ERR_ARG, "aggregate call in GROUP BY");
    const bool string_key = lower_bare_string_column(q, g.get());
    if (string_key) q->string_keys = true;   // (wire rows: the key hash covers the string bytes - made on the host, wire.cc)
    else lower_strings(q, g.get());
    q->group_is_string.push_back(string_key);
    if (g->type == EVQ_STRING || g->type == EVQ_NIL) fail(EVQGPU_ERR_UNSUPPORTED, "GROUP BY key type is outside the numeric device path");
    q->group.push_back(std::move(g));
  }
  const bool groupby = q->flags & EVQGPU_QUERY_GROUPBY;
  for (uint32_t i = 0; i < desc->num_select; ++i) {
    SelectItem item;
    item.expr = parse_program(desc->select[i]);
    if (!item.expr) fail(EVQGPU_ERR_ARG, "empty select expression");
    item.is_string = lower_bare_string_column(q, item.expr.get());
    if (!item.is_string) lower_strings(q, item.expr.get());
    if (item.expr->type == EVQ_STRING || item.expr->type == EVQ_NIL)
      fail(EVQGPU_ERR_UNSUPPORTED, "select item %u: result type is outside the numeric device path", i);
    item.agg = find_aggregate(item.expr.get());
    if (item.agg && !groupby) fail(EVQGPU_ERR_ARG, "aggregate call in a scan-only plan");
    if (item.agg)
      for (const auto& a : item.agg->args)
        if (find_aggregate(a.get())) fail(EVQGPU_ERR_ARG, "nested aggregate call");
    q->select.push_back(std::move(item));
  }
  if (desc->flags & EVQGPU_QUERY_COORDINATOR) {
    // the coordinator of a cluster GROUP BY (GroupByMergeExpression): nothing is scanned.  Groups are identified by the 20-byte
    // SHA-1 keys the shards send - three pseudo key words - and every non-aggregate item arrives as a value with each row
    if (!groupby) fail(EVQGPU_ERR_ARG, "EVQGPU_QUERY_COORDINATOR needs an aggregate plan");
    q->coordinator = true;
    q->group.clear();
    for (uint32_t i = 0; i < 3; ++i) {
      evqgpu_insn in;
      memset(&in, 0, sizeof(in));
      in.op = EVQ_X_LITERAL;
      in.type = EVQ_UINT64;
      in.imm = 0x5348413100000000ull + i;   // distinct signatures that no select item can spell
      evqgpu_expr e = {&in, 1, nullptr, 0};
      q->group.push_back(parse_program(e));
    }
    for (auto& item : q->select)
      if (!item.agg) { item.first = true; q->has_first = true; }
  } else if (groupby)
    for (auto& item : q->select)
      if (!item.agg && !is_function_of_keys(*q, item.expr.get())) {
        if (item.is_string) fail(EVQGPU_ERR_UNSUPPORTED, "a string select item of an aggregate plan must be a GROUP BY key");
        item.first = true;
        q->has_first = true;
      }
  // which input columns does the device actually have to read
  q->col_used.assign(q->input_columns.size(), false);
  collect_columns(q->where.get(), q->col_used);
  for (const auto& g : q->group) collect_columns(g.get(), q->col_used);
  for (const auto& item : q->select) {
    if (!groupby) collect_columns(item.expr.get(), q->col_used);
    else if (item.agg) for (const auto& a : item.agg->args) collect_columns(a.get(), q->col_used);
    else if (item.first) collect_columns(item.expr.get(), q->col_used);
  }
}

}  // namespace evq

extern "C" {

int evqgpu_query_create(evqgpu_ctx* ctx, const evqgpu_query_desc* desc, evqgpu_query** out) {
  return guarded([&] {
    if (!ctx || !desc || !out) fail(EVQGPU_ERR_ARG, "evqgpu_query_create: null argument");
    std::unique_ptr<evqgpu_query> q(new evqgpu_query());
    q->ctx = ctx;
    query_intake(q.get(), desc);
    use_device(ctx);
    q->status.alloc(16);
    q->counters.alloc(32);
    q->out_count.alloc(8);
    *out = q.release();
  });
}

void evqgpu_query_destroy(evqgpu_query* q) {
  if (!q) return;
  cudaSetDevice(q->ctx->device);
  delete q;
}

uint32_t evqgpu_query_num_columns(const evqgpu_query* q) { return q ? (uint32_t) q->select.size() : 0; }
uint32_t evqgpu_query_column_type(const evqgpu_query* q, uint32_t idx) {
  if (!q || idx >= q->select.size()) return 0;
  return q->select[idx].is_string ? (uint32_t) EVQ_STRING : (uint32_t) q->select[idx].expr->type;
}

}  // extern "C"

// ---- execution -------------------------------------------------------------------------------------------------------
namespace evq {

struct TablePlan {
  evqgpu_table* table;
  Binding binding;
  StageLayout layout;
  size_t smem = 0;
};

static KernelShape shape_of_plans(const evqgpu_query& q, const std::vector<TablePlan>& plans) {
  KernelShape s = shape_for(q, plans[0].table, plans[0].binding);
  for (size_t i = 1; i < plans.size(); ++i) {
    const KernelShape o = shape_for(q, plans[i].table, plans[i].binding);
    if (shape_key(o) != shape_key(s))
      fail(EVQGPU_ERR_UNSUPPORTED, "partitions of one query must share column encodings and nullability");
    widen_shape(s, o);
  }
  // every variable-length LEB128 column of every partition must carry a sub-index for the kernel to rely on it
  bool have_subidx = true;
  for (const auto& p : plans)
    for (size_t i = 0; i < s.cols.size(); ++i) {
      if (p.binding.col_index[i] < 0) continue;
      const Column& c = *p.binding.cols[i];
      if (c.data_kind == EVQ_KIND_LEB128 && s.cols[i].leb_len >= 2 && !s.cols[i].leb_uniform && !c.sub_index.p) have_subidx = false;
    }
  finish_shape(s, have_subidx);
  // external row filters (FastCSTableScan::setFilter): one more stream; a partition without one keeps all its rows
  bool any_filter = false;
  for (const auto& p : plans) any_filter = any_filter || p.table->has_filter;
  if (any_filter) {
    if (s.nstreams >= EVQ_MAX_STREAMS) fail(EVQGPU_ERR_UNSUPPORTED, "no stream left for the row filter");
    s.filter_stream = s.nstreams++;
    for (const auto& p : plans) {
      evqgpu_table* t = p.table;
      if (t->has_filter || t->filter.p) continue;
      use_device(t->ctx);
      t->filter.alloc(round_up((uint64_t) t->num_tiles * (EVQ_TILE_ROWS / 8), 256) + 256);
      EVQ_CUDA(cudaMemsetAsync(t->filter.p, 0xff, t->filter.bytes, t->ctx->stream));
    }
  }
  return s;
}

static void fill_streams(EvqScanParams& P, evqgpu_table* t, const Binding& b, const KernelShape& s, const StageLayout& L) {
  P.num_rows = t->num_rows;
  P.num_tiles = t->num_tiles;
  P.num_streams = (u32) s.nstreams;
  for (size_t i = 0; i < s.cols.size(); ++i) {
    const ColSig& cs = s.cols[i];
    if (!cs.used) continue;
    const Column& c = *b.cols[i];
    EvqStream& d = P.streams[cs.data_stream];
    d.base = c.data.buf.as<u8>();
    d.off_index = c.data_kind == EVQ_KIND_LEB128 ? c.off_index.as<u64>() : nullptr;
    d.val_index = cs.nullable ? c.val_index.as<u64>() : nullptr;
    d.nbytes = c.data.nbytes;
    d.kind = c.data_kind;
    d.bits = c.data_bits;
    d.smem_off = L.smem_off[cs.data_stream];
    d.smem_cap = L.smem_cap[cs.data_stream];
    if (cs.sub_stream >= 0) {
      EvqStream& x = P.streams[cs.sub_stream];
      x.base = c.sub_index.as<u8>();
      x.off_index = nullptr;
      x.val_index = nullptr;
      x.nbytes = c.sub_index.bytes;
      x.kind = EVQ_KIND_SUBIDX;
      x.bits = 0;
      x.smem_off = L.smem_off[cs.sub_stream];
      x.smem_cap = L.smem_cap[cs.sub_stream];
    }
    if (cs.nullable) {
      EvqStream& l = P.streams[cs.level_stream];
      l.base = c.dlevel.buf.as<u8>();
      l.off_index = nullptr;
      l.val_index = nullptr;
      l.nbytes = c.dlevel.nbytes;
      l.kind = EVQ_KIND_LEVEL;
      l.bits = c.level_bits;
      l.smem_off = L.smem_off[cs.level_stream];
      l.smem_cap = L.smem_cap[cs.level_stream];
    }
  }
  if (s.filter_stream >= 0) {
    EvqStream& f = P.streams[s.filter_stream];
    f.base = t->filter.as<u8>();
    f.off_index = nullptr;
    f.val_index = nullptr;
    f.nbytes = t->filter.bytes;
    f.kind = EVQ_KIND_FILTER;
    f.bits = 1;
    f.smem_off = L.smem_off[s.filter_stream];
    f.smem_cap = L.smem_cap[s.filter_stream];
  }
}

static uint64_t algorithmic_bytes(const evqgpu_query& q, const std::vector<TablePlan>& plans) {
  uint64_t n = 0;
  for (const auto& p : plans)
    for (size_t i = 0; i < q.input_columns.size(); ++i)
      if (p.binding.col_index[i] >= 0) {
        const Column& c = *p.binding.cols[i];
        n += c.data_payload_bytes + c.level_payload_bytes;
      }
  return n;
}

// choose thread count / stages so the CTA fits; returns dynamic smem bytes of the largest table
static void fit_shape(evqgpu_query& q, KernelShape& s, std::vector<TablePlan>& plans) {
  const size_t limit = (size_t) q.ctx->smem_optin;
  // fast kernel: 2 row tiles per stage first (half as many, twice as large bulk copies), if that costs no resident CTA
  int kt_first = s.fast ? 2 : 1;
  if (const char* e = getenv("EVQGPU_KT")) kt_first = s.fast ? std::max(1, std::min(4, atoi(e))) : 1;
  const int nattempts = kt_first > 1 ? 8 : 4;
  for (int attempt = 0; attempt < nattempts; ++attempt) {
    s.kt = (kt_first > 1 && attempt < 4) ? kt_first : 1;
    s.ncons = (attempt & 1) ? 128 : 256;
    s.nstages = (attempt & 2) ? 2 : 3;
    if (s.fast) {
      // 8 consecutive rows per thread first (halves the per-thread fixed work of every tile).  The kernel is bound by
      // dependent shared-memory / ALU chains, not by the depth of the copy pipeline (profiles/r01_sweep_q1.txt: 2, 3 and 4
      // stages time the same, every extra resident CTA adds throughput): 2 stages, as many CTAs per SM as fit
      s.ncons = (attempt & 1) ? 256 : 128;
      s.nstages = 2;
      if (const char* e = getenv("EVQGPU_FAST_NCONS")) s.ncons = atoi(e) == 256 ? 256 : 128;
      if (s.use_subidx) s.ncons = 128;   // the sub-index holds the entry point of every 8th value = one per thread
    }
    if (const char* e = getenv("EVQGPU_NSTAGES")) s.nstages = std::max(2, std::min(4, atoi(e)));
    size_t worst = 0;
    for (auto& p : plans) {
      p.layout = stage_layout(q, p.table, p.binding, s);
      p.smem = header_bytes(s) + (size_t) s.nstages * p.layout.stage_bytes + scratch_bytes(s) + acc_bytes(q, s);
      worst = std::max(worst, p.smem);
    }
    if (worst <= limit) {
      // as many CTAs per SM as shared memory allows (latency hiding for the decode phases), within the register budget
      // the launch bounds leave per thread: 2 x 288 threads or 4 x 160 threads
      const int by_smem = (int) ((227 * 1024) / (worst + 1024));
      // (the byte-plane accumulators of the dense tier are registers: fewer CTAs, more registers per thread)
      const int nacc = q.nnarrow * std::max(1, q.plane_groups);
      const int cap = s.ncons <= 128 ? (nacc > 60 ? 3 : nacc > 36 ? 4 : nacc > 16 ? 5 : 6) : 2;
      if (s.kt > 1 && by_smem < cap && !getenv("EVQGPU_KT")) { attempt |= 3; continue; }   // larger stages would cost residency: one tile per stage
      s.min_ctas = std::max(1, std::min(by_smem, cap));
      if (const char* e = getenv("EVQGPU_MAX_CTAS")) s.min_ctas = std::max(1, std::min(s.min_ctas, atoi(e)));
      return;
    }
  }
  fail(EVQGPU_ERR_UNSUPPORTED, "row tile of this query does not fit shared memory (%zu columns)", q.input_columns.size());
}

static void run_scan(evqgpu_query& q, const KernelShape& s, std::vector<TablePlan>& plans, EvqScanParams base,
                     const std::function<void(unsigned, uint64_t, EvqScanParams&)>& before_table = nullptr,
                     const std::function<void(unsigned)>& after_table = nullptr) {
  evqgpu_ctx* ctx = q.ctx;
  cudaKernel_t kern = q.module->kernels.at("evq_scan");
  size_t max_smem = 0;
  for (const auto& p : plans) max_smem = std::max(max_smem, p.smem);
  EVQ_CUDA(cudaFuncSetAttribute((const void*) kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) max_smem));
  uint64_t tile_row_base = 0;
  for (auto& p : plans) {
    if (p.table->num_tiles == 0) continue;
    EvqScanParams P = base;
    fill_streams(P, p.table, p.binding, s, p.layout);
    P.tile_row_base = tile_row_base;
    tile_row_base += p.table->num_tiles;
    u32 stage_bytes = p.layout.stage_bytes;
    const int ctas_per_sm = std::max<int>(1, std::min<size_t>(s.min_ctas, (227 * 1024) / (p.smem + 1024)));
    const uint64_t groups = (p.table->num_tiles + s.kt - 1) / s.kt;
    uint64_t grid64 = std::min<uint64_t>(groups, (uint64_t) ctx->sm_count * ctas_per_sm);
    // the u32 byte-plane accumulators of a thread take at most 8 * 255 * 255 per row tile: bound the tiles per CTA
    if (q.nnarrow > 0) grid64 = std::max<uint64_t>(grid64, (p.table->num_tiles + 7999) / 8000);
    const unsigned grid = (unsigned) grid64;
    void* args[] = {&P, &stage_bytes};
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->profiling) {
      EVQ_CUDA(cudaEventCreate(&e0));
      EVQ_CUDA(cudaEventCreate(&e1));
      EVQ_CUDA(cudaEventRecord(e0, ctx->stream));
    }
    if (before_table) before_table(grid, p.table->num_rows, P);
    launch(ctx, kern, dim3(grid), dim3(s.ncons + 32), p.smem, args);
    q.stats.kernel_launches++;
    if (after_table) after_table(grid);   // (partitioned aggregation: pass 2 belongs to the table's scan - and to its timing)
    if (e0) {
      EVQ_CUDA(cudaEventRecord(e1, ctx->stream));
      q.prof_events.push_back({e0, e1});
    }
  }
}

// smallest dense slot capacity the kernel is specialised for
static int g1_for(uint64_t slots) {
  int g = 2;
  while ((uint64_t) g < slots) g <<= 1;
  return g;
}

// only a bare column reference (possibly through `if`) keeps its NULL tag; every function call drops it
static bool key_may_be_null(const Expr* e, const std::vector<bool>& nullable_cols) {
  if (e->op == EVQ_X_INPUT) return e->col < nullable_cols.size() && nullable_cols[e->col];
  if (e->op == EVQ_X_IF) return key_may_be_null(e->args[1].get(), nullable_cols) || key_may_be_null(e->args[2].get(), nullable_cols);
  return false;
}

static bool key_dense_capable(const Expr* g) {
  return g->type == EVQ_UINT64 || g->type == EVQ_TIMESTAMP64 || g->type == EVQ_BOOL || g->type == EVQ_INT64;
}

// Bounds pre-pass: min/max of every GROUP BY expression over the tables, computed with the scan kernel itself
// (a single-group aggregate query), so that a handful of groups can be mapped to dense accumulator slots.
// local_key_bounds fills, per key, {valid, min, max, may be NULL}; false = the keys are not dense-capable at all.
static const size_t kBW = 4;

// -> 1 bounds known, 0 the key types rule the dense tiers out (a property of the plan), -1 the pre-pass failed HERE
// (e.g. a key expression divides by zero on this rank's rows)
static int local_key_bounds(evqgpu_query& q, std::vector<evqgpu_table*>& tables, std::vector<uint64_t>& bounds) {
  const size_t nk = q.group.size();
  bounds.assign(kBW * nk, 0);
  for (const auto& g : q.group)
    if (!key_dense_capable(g.get())) return 0;
  // plan: select min(k0), max(k0), ... ("has NULL" follows from the nullability of the key's columns)
  std::vector<std::vector<evqgpu_insn>> codes;
  std::vector<evqgpu_expr> sel;
  std::vector<const char*> names;
  for (const auto& n : q.input_columns) names.push_back(n.c_str());
  // serialise group expression i back to postfix and wrap it in min()/max()
  struct Ser {
    static void emit(const Expr* e, std::vector<evqgpu_insn>& out) {
      for (const auto& a : e->args) emit(a.get(), out);
      evqgpu_insn in;
      memset(&in, 0, sizeof(in));
      in.op = (uint8_t) e->op;
      in.type = (uint8_t) e->type;
      in.nargs = (uint16_t) e->args.size();
      in.arg = e->op == EVQ_X_CALL ? (uint32_t) e->fn : e->col;
      in.imm = e->imm;
      out.push_back(in);
    }
  };
  auto type_name = [](int t) { return t == EVQ_INT64 ? "int64" : "uint64"; };
  for (size_t i = 0; i < nk; ++i) {
    const Expr* g = q.group[i].get();
    // bool / timestamp keys are compared through their raw bits as uint64: wrap in a cast-free min/max on the bits
    for (int mm = 0; mm < 2; ++mm) {
      std::vector<evqgpu_insn> code;
      Ser::emit(g, code);
      if (code.back().type == EVQ_BOOL || code.back().type == EVQ_TIMESTAMP64) {
        // to_int64(bool|timestamp64) keeps the order of these non-negative values
        std::string conv = std::string("to_int64#int64/") + (g->type == EVQ_BOOL ? "bool;" : "timestamp64;");
        evqgpu_insn c;
        memset(&c, 0, sizeof(c));
        c.op = EVQ_X_CALL; c.type = EVQ_INT64; c.nargs = 1; c.arg = (uint32_t) function_lookup(conv);
        code.push_back(c);
      }
      const int ty = code.back().type;
      std::string sym = std::string(mm ? "max#" : "min#") + type_name(ty) + "/" + type_name(ty) + ";";
      evqgpu_insn c;
      memset(&c, 0, sizeof(c));
      c.op = EVQ_X_CALL; c.type = (uint8_t) ty; c.nargs = 1; c.arg = (uint32_t) function_lookup(sym);
      code.push_back(c);
      codes.push_back(std::move(code));
    }
  }
  for (auto& c : codes) sel.push_back({c.data(), (uint32_t) c.size(), nullptr, 0});
  evqgpu_query_desc d;
  memset(&d, 0, sizeof(d));
  d.struct_size = sizeof(d);
  d.flags = EVQGPU_QUERY_GROUPBY;
  d.num_input_columns = (uint32_t) names.size();
  d.input_columns = names.data();
  d.num_select = (uint32_t) sel.size();
  d.select = sel.data();
  evqgpu_query* bq = nullptr;
  if (evqgpu_query_create(q.ctx, &d, &bq) != EVQGPU_OK) return -1;
  std::unique_ptr<evqgpu_query, void (*)(evqgpu_query*)> guard(bq, evqgpu_query_destroy);
  bq->col_is_string = q.col_is_string;   // the re-serialised key expressions already read string columns as dictionary codes
  if (evqgpu_query_execute(bq, tables.data(), (uint32_t) tables.size()) != EVQGPU_OK) return -1;
  q.stats.kernel_launches += bq->stats.kernel_launches;
  q.jit_ms_total += bq->jit_ms_total;
  uint64_t nrows = 0;
  evqgpu_query_num_rows(bq, &nrows);
  std::vector<bool> nullable_cols(q.input_columns.size(), false);
  for (auto* t : tables)
    for (size_t i = 0; i < q.input_columns.size(); ++i) {
      const int ci = t->find(q.input_columns[i].c_str());
      if (ci >= 0 && t->cols[ci].meta.dlevel_max > 0) nullable_cols[i] = true;
    }
  // only a bare column reference keeps its NULL tag as a key: one extra slot index then
  for (size_t i = 0; i < nk; ++i) bounds[kBW * i + 3] = key_may_be_null(q.group[i].get(), nullable_cols) ? 1 : 0;
  if (nrows != 0) {
    std::vector<std::vector<uint8_t>> bufs(sel.size(), std::vector<uint8_t>(9));
    std::vector<void*> ptrs;
    for (auto& b : bufs) ptrs.push_back(b.data());
    uint64_t got = 0;
    if (evqgpu_query_fetch(bq, 0, 1, ptrs.data(), &got) != EVQGPU_OK || got != 1) return -1;
    for (size_t i = 0; i < nk; ++i) {
      bounds[kBW * i] = 1;
      memcpy(&bounds[kBW * i + 1], bufs[2 * i].data(), 8);
      memcpy(&bounds[kBW * i + 2], bufs[2 * i + 1].data(), 8);
    }
  }
  return 1;
}

// the bounds of several ranks -> the bounds of the whole job (SURVEY 8e "canonical slot assignment")
static void combine_key_bounds(const evqgpu_query& q, const std::vector<std::vector<uint64_t>>& per_rank, std::vector<uint64_t>& bounds) {
  const size_t nk = q.group.size();
  bounds.assign(kBW * nk, 0);
  for (size_t i = 0; i < nk; ++i) {
    const bool is_signed = q.group[i]->type != EVQ_UINT64;   // int64 / bool / timestamp64 went through int64
    uint64_t valid = 0, mn = 0, mx = 0, may_null = 0;
    for (const auto& rb : per_rank) {
      const uint64_t* b = &rb[kBW * i];
      may_null |= b[3];
      if (!b[0]) continue;
      if (!valid) { valid = 1; mn = b[1]; mx = b[2]; continue; }
      if (is_signed) {
        if ((int64_t) b[1] < (int64_t) mn) mn = b[1];
        if ((int64_t) b[2] > (int64_t) mx) mx = b[2];
      } else {
        if (b[1] < mn) mn = b[1];
        if (b[2] > mx) mx = b[2];
      }
    }
    bounds[kBW * i] = valid; bounds[kBW * i + 1] = mn; bounds[kBW * i + 2] = mx; bounds[kBW * i + 3] = may_null;
  }
}

static bool dense_map_from_bounds(const evqgpu_query& q, const std::vector<uint64_t>& bounds, DenseMap& dm) {
  const size_t nk = q.group.size();
  for (size_t i = 0; i < nk; ++i) {
    if (!bounds[kBW * i]) {   // no row anywhere: any mapping works
      dm.key_min[i] = 0; dm.key_range[i] = 2; dm.key_null_idx[i] = 1;
      continue;
    }
    const uint64_t mn = bounds[kBW * i + 1], mx = bounds[kBW * i + 2];
    // signed and unsigned keys alike: (key - min) as an unsigned difference
    const uint64_t span = mx - mn;
    if (span > (1ull << 24) - 2) return false;
    dm.key_min[i] = mn;
    // one extra index for NULL keys, only when the expression can carry a NULL tag at all
    const bool may_null = bounds[kBW * i + 3] != 0;
    dm.key_range[i] = span + (may_null ? 2 : 1);
    dm.key_null_idx[i] = may_null ? span + 1 : ~0ull;
  }
  uint64_t slots = 1;
  for (size_t i = nk; i-- > 0;) {
    dm.key_stride[i] = slots;
    if (dm.key_range[i] > (1ull << 24) || slots * dm.key_range[i] > (1ull << 24)) return false;
    slots *= dm.key_range[i];
  }
  dm.slots = slots;
  return true;
}

static bool is_multi_rank_partial(const evqgpu_query& q) {
  return (q.flags & EVQGPU_QUERY_PARTIAL) && (q.flags & EVQGPU_QUERY_GROUPBY) && q.ctx->nccl_comm && q.ctx->nranks > 1;
}

static uint64_t fnv1a(const std::string& s) {
  uint64_t h = 1469598103934665603ull;
  for (unsigned char ch : s) h = (h ^ ch) * 1099511628211ull;
  return h;
}

// GroupByExpression::nextBatch (groupby.cc:187-220) for all groups at once: the emit kernel evaluates every select
// item's `get` side per group and writes packed SVector columns
void emit_results(evqgpu_query& q) {
  evqgpu_ctx* ctx = q.ctx;
  const uint64_t emit_slots = q.emit.slots;
  const uint64_t out_cap = (q.shape.tier == 1 || q.shape.dense_global) ? std::min<uint64_t>(emit_slots, std::max<uint64_t>(q.emit_total_rows, 64)) : std::min<uint64_t>(emit_slots, std::max<uint64_t>(q.emit_total_rows, 1));
  q.out_cols.resize(q.select.size());
  for (size_t i = 0; i < q.select.size(); ++i)
    ensure(q.out_cols[i], out_cap * (q.select[i].expr->type == EVQ_BOOL ? 2 : 9) + 16);
  q.out_capacity = out_cap;
  EVQ_CUDA(cudaMemsetAsync(q.out_count.p, 0, 8, ctx->stream));
  EmitParams ep = q.emit;
  ep.out_count = q.out_count.as<u64>();
  ep.out_capacity = out_cap;
  for (size_t i = 0; i < q.select.size(); ++i) ep.out_cols[i] = q.out_cols[i].as<u8>();
  if (q.flags & EVQGPU_QUERY_WIRE) {
    ensure(q.out_sha, out_cap * wire_key_stride(q) + 16);
    ensure(q.out_state, out_cap * std::max<size_t>(1, q.state_ops.size()) * 8 + 16);
    ep.out_sha = q.out_sha.as<u8>();
    ep.out_state = q.out_state.as<u64>();
  }
  void* args[] = {&ep};
  launch(ctx, q.module->kernels.at("evq_emit"), dim3((unsigned) ((emit_slots + 255) / 256)), dim3(256), 0, args);
  q.stats.kernel_launches++;
  q.emitted = true;
}

// the tail of a dense-tier execution (codegen.cc evq_tail): with `merge` the ranks' state words are exchanged through the
// peer-mapped buffers of comm.cc and combined in rank order; then emit, publish, re-arm
void launch_tail(evqgpu_query& q, bool merge) {
  evqgpu_ctx* ctx = q.ctx;
  const uint64_t emit_slots = q.emit.slots;
  const uint64_t out_cap = std::min<uint64_t>(emit_slots, std::max<uint64_t>(q.emit_total_rows, 64));
  q.out_cols.resize(q.select.size());
  for (size_t i = 0; i < q.select.size(); ++i)
    ensure(q.out_cols[i], out_cap * (q.select[i].expr->type == EVQ_BOOL ? 2 : 9) + 16);
  q.out_capacity = out_cap;
  TailParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.E = q.emit;
  tp.E.out_capacity = out_cap;
  for (size_t i = 0; i < q.select.size(); ++i) tp.E.out_cols[i] = q.out_cols[i].as<u8>();
  if (q.flags & EVQGPU_QUERY_WIRE) {
    ensure(q.out_sha, out_cap * wire_key_stride(q) + 16);
    ensure(q.out_state, out_cap * std::max<size_t>(1, q.state_ops.size()) * 8 + 16);
    tp.E.out_sha = q.out_sha.as<u8>();
    tp.E.out_state = q.out_state.as<u64>();
  }
  tp.dense_state = q.dense_base;
  tp.ctl = q.ctl.as<u64>();
  tp.nranks = 1;
  tp.rank = 0;
  if (merge) {
    if (!ctx->p2p_ok) fail(EVQGPU_ERR_RUNTIME, "internal error: fused merge without peer-mapped buffers");
    tp.nranks = (u32) ctx->nranks;
    tp.rank = (u32) ctx->rank;
    tp.epoch = ++ctx->p2p_epoch;
    tp.xbuf_local = (u64*) ctx->p2p_local;
    tp.flags_local = (u64*) ((uint8_t*) ctx->p2p_local + EVQ_P2P_FLAGS_OFFSET);
    for (int r = 0; r < ctx->nranks; ++r) {
      tp.xbuf_peer[r] = (u64*) ctx->p2p_peer[r];
      tp.flags_peer[r] = (u64*) ((uint8_t*) ctx->p2p_peer[r] + EVQ_P2P_FLAGS_OFFSET);
    }
  }
  void* args[] = {&tp};
  launch(ctx, q.module->kernels.at("evq_tail"), dim3(1), dim3(256), 0, args);
  q.stats.kernel_launches++;
  q.emitted = true;
  q.tail_done = true;
  q.armed = true;
  q.armed_sig = q.module_sig;
}

static std::string layout_signature(const evqgpu_query& q) {
  std::string sig;
  for (size_t i = 0; i < q.state_keys.size(); ++i) sig += q.state_keys[i] + "#" + std::to_string(q.state_ops[i]) + ";";
  return sig;
}

// evqgpu_query_prepare: everything the ranks of a multi-rank job must agree on BEFORE the scan, in ONE unconditional
// collective (an all-gather every rank enters no matter what happened locally):
//   * did anything fail locally (binding the tables, the key-bounds pre-pass)?  -> the call fails on every rank
//   * the aggregate state layout (depends on which columns are optional in the rank's partitions)
//   * the key bounds -> the canonical key -> slot assignment of the dense tiers
// evqgpu_query_enqueue / _merge then run without any data-dependent decision about entering a collective.
void prepare_query(evqgpu_query& q, std::vector<evqgpu_table*>& tv) {
  if (!is_multi_rank_partial(q)) return;   // nothing to agree on
  evqgpu_ctx* ctx = q.ctx;
  const size_t nk = q.group.size();
  std::vector<uint64_t> uids;
  for (auto* t : tv) uids.push_back(t ? t->uid : 0);
  uint64_t failed = 0, capable = 0, layout = 0;
  std::vector<uint64_t> bounds(kBW * nk, 0);
  std::string msg;
  std::vector<TablePlan> plans;
  try {
    for (auto* t : tv) {
      if (!t) fail(EVQGPU_ERR_ARG, "null table");
      if (t->ctx != q.ctx) fail(EVQGPU_ERR_ARG, "table belongs to another context");
      TablePlan p;
      p.table = t;
      p.binding = bind_table(q, t);
      plans.push_back(std::move(p));
    }
    if (plans.empty()) fail(EVQGPU_ERR_ARG, "evqgpu_query_prepare: no tables");
  } catch (const Error& e) {
    failed = 1;
    msg = e.msg;
  }
  // a plan that reads string columns as dictionary codes: the ranks' dictionaries are made identical first (collective,
  // entered by every rank of such a plan whatever happened above), so that equal strings have equal codes everywhere -
  // string GROUP BY keys and string predicates then merge like integers
  bool uses_strings = !q.string_literals.empty();
  for (bool b : q.col_is_string) uses_strings = uses_strings || b;
  if (uses_strings) {
    sync_dictionary(ctx);
    for (auto& lit : q.string_literals) {
      const uint64_t code = string_code(ctx, lit.second);
      if (lit.first->imm != code) { lit.first->imm = code; q.module_sig.clear(); }
    }
  }
  uids.clear();   // (the synchronisation gives renumbered tables a new identity)
  for (auto* t : tv) uids.push_back(t ? t->uid : 0);
  try {
    if (failed) throw Error{EVQGPU_ERR_RUNTIME, msg};
    KernelShape s = shape_of_plans(q, plans);
    layout_states(q, s);
    layout = fnv1a(layout_signature(q));
    if (nk > 0 && q.expected_groups <= (1ull << 24)) {
      const int rc = local_key_bounds(q, tv, bounds);
      if (rc < 0) fail(EVQGPU_ERR_RUNTIME, "key bounds pre-pass failed: %s", last_error());
      capable = (uint64_t) rc;
    }
  } catch (const Error& e) {
    failed = 1;
    msg = e.msg;
  }
  std::vector<uint64_t> mine = {failed, capable, layout};
  mine.insert(mine.end(), bounds.begin(), bounds.end());
  const std::vector<uint64_t> all = comm_all_gather_host(ctx, mine);   // every rank, every time
  const size_t W = mine.size();
  for (int r = 0; r < ctx->nranks; ++r)
    if (all[(size_t) r * W])
      fail(EVQGPU_ERR_RUNTIME, "evqgpu_query_prepare: rank %d failed%s%s", r, r == ctx->rank ? ": " : "", r == ctx->rank ? msg.c_str() : "");
  for (int r = 0; r < ctx->nranks; ++r)
    if (all[(size_t) r * W + 2] != layout)
      fail(EVQGPU_ERR_ARG, "evqgpu_query_prepare: ranks disagree on the aggregate state layout (partitions differ in which columns are optional)");
  bool dense = nk > 0;
  std::vector<std::vector<uint64_t>> per_rank;
  for (int r = 0; r < ctx->nranks; ++r) {
    dense = dense && all[(size_t) r * W + 1] == 1;
    per_rank.emplace_back(all.begin() + (size_t) r * W + 3, all.begin() + (size_t) (r + 1) * W);
  }
  DenseMap dm;
  if (dense) {
    combine_key_bounds(q, per_rank, bounds);
    dense = dense_map_from_bounds(q, bounds, dm);
  }
  q.dense_cache_valid = true;
  q.dense_cache_ok = dense;
  q.dense_cache_map = dm;
  q.dense_cache_uids = uids;
  q.prepared = true;
  q.prepared_uids = uids;
}

static void execute_groupby(evqgpu_query& q, std::vector<TablePlan>& plans, std::vector<evqgpu_table*>& tables, bool sync) {
  evqgpu_ctx* ctx = q.ctx;
  if (is_multi_rank_partial(q)) {
    std::vector<uint64_t> uids;
    for (auto* t : tables) uids.push_back(t->uid);
    if (!q.prepared || uids != q.prepared_uids)
      fail(EVQGPU_ERR_ARG, "multi-rank job: call evqgpu_query_prepare (on every rank) for this set of tables before "
                           "evqgpu_query_enqueue; evqgpu_query_execute does it implicitly");
  }
  KernelShape s = shape_of_plans(q, plans);
  layout_states(q, s);
  uint64_t total_rows = 0;
  for (auto* t : tables) total_rows += t->num_rows;

  // ---- tier decision
  const size_t nk = q.group.size();
  DenseMap dm;
  s.tier = 1;
  s.g1 = 1;
  if (nk > 0) {
    bool dense = false;
    std::vector<uint64_t> uids;
    for (auto* t : tables) uids.push_back(t->uid);
    if (q.dense_cache_valid && uids == q.dense_cache_uids) {
      dense = q.dense_cache_ok;
      dm = q.dense_cache_map;
    } else if (q.expected_groups <= (1ull << 24)) {   // (0 = unknown)
      // (a multi-rank partial plan never gets here: evqgpu_query_prepare agreed on the map, see prepare_query)
      std::vector<uint64_t> bounds;
      dense = local_key_bounds(q, tables, bounds) == 1 && dense_map_from_bounds(q, bounds, dm);
      q.dense_cache_valid = true;
      q.dense_cache_ok = dense;
      q.dense_cache_map = dm;
      q.dense_cache_uids = uids;
    }
    bool dense_global = false;
    if (dense) {
      s.g1 = g1_for(dm.slots);
      // thread-private accumulators must fit next to the pipeline stages
      const size_t acc = (size_t) s.g1 * q.nstate_smem * 128 * 8;   // (both kernels can run 128 consumer threads)
      if (dm.slots > 64 || acc > 96 * 1024) {
        dense = false;
        // too many groups for thread-private state, but the key tuples span a small box: a direct-addressed group
        // array in global memory (no keys, no probing), which stays L2-resident up to a few million groups.  A
        // multi-rank job keeps the hash tier: its merge moves groups, not boxes.
        // In a multi-rank job the array is merged with ONE ncclAllReduce(sum) over all its words (merge.cu), which is
        // exact when every word is a wrapping 64-bit sum / count: no min / max / double words, no carry words, no first-row pairs.
        bool all_add = true;
        for (size_t w = 0; w < q.state_ops.size(); ++w)
          all_add = all_add && q.state_ops[w] == OP_ADD_U64 && q.state_carry_of[w] < 0 && q.state_keys[w].compare(0, 3, "cd:") != 0;
        const bool multi = (q.flags & EVQGPU_QUERY_PARTIAL) && ctx->nranks > 1;
        dense_global = dm.slots * q.state_ops.size() * 8 <= (1ull << 30) && (!multi || (all_add && !getenv("EVQGPU_NO_DENSE_GLOBAL_MERGE"))) &&
                       !getenv("EVQGPU_NO_DENSE_GLOBAL");
      }
    }
    if (!dense) { s.tier = 2; s.g1 = 1; s.dense_global = dense_global; }
  }
  if ((s.tier == 1 && s.g1 > 1) || s.dense_global) s.dense = dm;
  layout_narrow(q, s);
  // hash tier with a group table far beyond L2: partitioned aggregation.  Pass 1 (the scan) writes the rows that pass WHERE
  // as records into 2^part_bits partitions by the top bits of their group's home slot; the passes behind it aggregate one
  // table slice at a time - in shared memory after a second partitioning level (evq_repart + evq_agg_smem, the default), or
  // L2-resident (evq_agg_part) - sequential record traffic instead of a random HBM sector pair (and its write-back) per
  // row.  A partition that overflows (heavily skewed keys, for which the direct tier is the right
  // one anyway: hot groups live in L2) makes the query fall back to the direct tier for good.
  uint64_t ht_want = 0, part_cap = 0;
  if (s.tier == 2 && !s.dense_global) {
    ht_want = q.ht_cap;
    if (ht_want == 0) {
      uint64_t est = q.expected_groups ? q.expected_groups : std::min<uint64_t>(total_rows, 1ull << 26);
      ht_want = next_pow2(std::max<uint64_t>(1024, est * 2));
    }
    const uint32_t stride = (uint32_t) round_up(1 + nk + q.state_ops.size(), 4);
    const uint64_t table_bytes = ht_want * 8 * stride;
    uint64_t slice = 32ull << 20, min_bytes = 128ull << 20;
    if (const char* e = getenv("EVQGPU_PART_SLICE_MB")) slice = std::max<uint64_t>(1, strtoull(e, nullptr, 10)) << 20;
    if (const char* e = getenv("EVQGPU_PART_MIN_MB")) min_bytes = strtoull(e, nullptr, 10) << 20;
    if (s.fast && table_bytes >= min_bytes && q.distinct_args.empty() && !q.has_first && !q.no_partition && nk > 0 &&
        !getenv("EVQGPU_NO_PARTITION")) {
      int bits = 1;
      while ((table_bytes >> bits) > slice && bits < 8) ++bits;
      // table slices in shared memory (the default): slices of <= 100 KB of slot words, reached over two partitioning levels of
      // at most 8 bits each; a table beyond that (or EVQGPU_NO_SMEM_SLICES) keeps the L2-resident form of pass 2
      {
        const uint64_t words = 1 + nk + q.state_ops.size();
        uint64_t slots = 64;
        while (slots * 2 * words * 8 <= 100 * 1024) slots *= 2;   // (two buffers of <= 100 KB: one CTA of 1024 threads per SM)
        int need = 0;
        while ((ht_want >> need) > slots) ++need;
        if (need <= 16 && !getenv("EVQGPU_NO_SMEM_SLICES")) {
          s.slice_slots = (int) slots;
          bits = std::max(std::min(bits, need), need - 8);
          if (bits < 1) bits = 1;
        }
      }
      s.part_bits = bits;
      std::vector<bool> used(q.input_columns.size(), false);
      for (const auto& g : q.group) collect_columns(g.get(), used);
      for (const auto& item : q.select)
        if (item.agg) for (const auto& a : item.agg->args) collect_columns(a.get(), used);
      for (size_t i = 0; i < used.size(); ++i)
        if (used[i]) s.rec_cols.push_back((int) i);
      part_cap = 1;   // (the partitions' capacity follows from every table's row count, see below)
    }
  }
  fit_shape(q, s, plans);
  q.shape = s;
  q.dense = dm;
  q.stats.strategy = s.dense_global ? 3u : s.part_bits > 0 ? 4u : (uint32_t) s.tier;

  // ---- kernel text: everything the generated text depends on is summarised in a short signature, so that a repeated
  // execution (the common case: same plan, same partitions) skips spelling and hashing ~300 KB of source
  float ms = 0;
  {
    std::string sig;
    char buf[160];
    snprintf(buf, sizeof(buf), "t%d g%d n%d s%d c%d f%d x%d N%d K%d P%d|", s.tier, s.g1, s.ncons, s.nstages, s.min_ctas, (int) s.fast,
             (int) s.use_subidx, q.nnarrow, s.kt * 1000 + (s.filter_stream + 1) * 10 + (int) s.dense_global, s.part_bits + 100 * s.slice_slots);
    sig += buf;
    for (const auto& c : s.cols) {
      snprintf(buf, sizeof(buf), "%d.%u.%u.%d.%u.%u.%u%s.%d.%d.%d.%llu.%llu.%llu.%d;", (int) c.used, c.sql_type, c.kind, (int) c.nullable, c.dmax, c.bits,
               c.leb_len, c.leb_uniform ? "u" : "", c.gen_slot, c.sub_stream, c.data_stream, (unsigned long long) c.vmax, (unsigned long long) c.vmin,
               (unsigned long long) c.vmin_present, c.nv_slot);
      sig += buf;
    }
    for (size_t i = 0; i < nk && !s.dense_global; ++i) {   // (the direct-addressed array takes its bounds as parameters)
      snprintf(buf, sizeof(buf), "k%llu.%llu.%llu.%llu;", (unsigned long long) s.dense.key_min[i], (unsigned long long) s.dense.key_stride[i],
               (unsigned long long) s.dense.key_null_idx[i], (unsigned long long) s.dense.key_range[i]);
      sig += buf;
    }
    for (size_t i = 0; i < q.state_keys.size(); ++i) sig += q.state_keys[i] + "#" + std::to_string(q.state_narrow[i]) + ";";
    for (int c : q.narrow_col) sig += "p" + std::to_string(c);
    sig += q.plane_sig;
    if (!q.module || sig != q.module_sig) {
      q.kernel_source = generate_source(q, s);
      std::vector<std::string> names = {"evq_scan", "evq_init", "evq_emit"};
      if (s.tier == 1 && !s.dense_global) names.push_back("evq_tail");
      if (s.part_bits > 0) names.push_back("evq_agg_part");
      if (s.part_bits > 0 && s.slice_slots > 0) { names.push_back("evq_repart"); names.push_back("evq_agg_smem"); }
      q.module = jit_compile(ctx, q.kernel_source, names, &ms);
      q.module_sig = sig;
      if (ms > 0 && q.module->from_disk) q.stats.jit_disk_hits++;
    }
  }
  q.jit_ms_total += ms;

  // ---- state
  const size_t nstate = q.state_ops.size();
  EvqScanParams base;
  memset(&base, 0, sizeof(base));
  base.ord_base = (u64) ctx->rank << 44;   // rank-major row ordinals (first-row items of a multi-rank job)
  // dense tier: the execution ends in ONE tail kernel (merge over NVLink + emit + re-arm, codegen.cc evq_tail) that leaves
  // the state words at their identities and the control block zeroed, so that the next execution of the same kernel
  // starts without an init kernel and without memsets: a step is the scan launches + the tail
  const bool tail = s.tier == 1 && !s.dense_global && q.distinct_args.empty() && !getenv("EVQGPU_NO_TAIL");
  q.use_tail = tail;
  q.tail_done = false;
  bool armed = false;
  if (tail) {
    if (!q.ctl.p) { q.ctl.alloc(128); q.armed = false; }
    base.status = q.ctl.as<u32>();
    base.counters = q.ctl.as<u64>() + 2;
    armed = q.armed && q.armed_sig == q.module_sig && !getenv("EVQGPU_NO_REARM");
    if (!armed) EVQ_CUDA(cudaMemsetAsync(q.ctl.p, 0, 128, ctx->stream));
  } else {
    base.status = q.status.as<u32>();
    base.counters = q.counters.as<u64>();
    EVQ_CUDA(cudaMemsetAsync(q.status.p, 0, 16, ctx->stream));
    EVQ_CUDA(cudaMemsetAsync(q.counters.p, 0, 32, ctx->stream));
  }
  q.armed = false;   // (until this execution's tail has been enqueued)
  InitParams ip;
  memset(&ip, 0, sizeof(ip));
  uint64_t emit_slots = 0;
  if (s.tier == 1 || s.dense_global) {
    const uint64_t slots = s.dense_global ? dm.slots : s.g1 > 1 ? (uint64_t) s.g1 : 1;
    if (q.dense_state.bytes < slots * nstate * 8 + 16) { q.dense_state.alloc(slots * nstate * 8 + 16); armed = false; }
    // first-row pairs must be 16-byte aligned in both tiers: their word index w has (1 + nk + w) even (layout_states)
    q.dense_base = q.dense_state.as<u64>() + ((q.has_first && ((1 + nk) & 1)) ? 1 : 0);
    base.dense_state = q.dense_base;
    base.dense_slots = (s.g1 > 1 || s.dense_global) ? dm.slots : 1;
    for (size_t i = 0; i < nk; ++i) {
      base.key_min[i] = dm.key_min[i];
      base.key_stride[i] = dm.key_stride[i];
      base.key_null_idx[i] = dm.key_null_idx[i];
      base.key_span[i] = dm.key_range[i] - (dm.key_null_idx[i] != ~0ull ? 2 : 1);
    }
    ip.dense_state = base.dense_state;
    ip.slots = slots;
    emit_slots = (s.g1 > 1 || s.dense_global) ? dm.slots : 1;
  } else {
    const uint64_t want = ht_want;
    q.ht_cap = want;
    // whole 32-byte sectors, 8 words = one 64-byte DRAM atom (a probe + its updates touch one atom); a table that is only ever
    // touched slice-wise in shared memory (partitioned aggregation, evq_agg_smem) is compact: a slice is one contiguous block
    const uint32_t stride = s.part_bits > 0 && s.slice_slots > 0 ? (uint32_t) (1 + nk + nstate) : (uint32_t) round_up(1 + nk + nstate, 4);
    ensure(q.ht_slots, want * 8 * stride);
    base.ht.slots = q.ht_slots.as<u64>();
    base.ht.stride = stride;
    base.ht.nkeys = (u32) nk;
    base.ht.cap = want;
    ip.ht = base.ht;
    ip.slots = want;
    emit_slots = want;
  }
  if (!armed) {
    void* args[] = {&ip};
    launch(ctx, q.module->kernels.at("evq_init"), dim3((unsigned) ((ip.slots + 255) / 256)), dim3(256), 0, args);
    q.stats.kernel_launches++;
  }
  // count_distinct: one empty (group, value) set per distinct argument; grown x4 with the group table when it fills up
  if (!q.distinct_args.empty()) {
    // across ranks the sets are merged in the dense tier (merge.cu merge_distinct_dense: the slot index names the group on
    // every rank) and in the hash tier (merge_hash: the members follow their group's keys to its owner); the
    // direct-addressed array is merged by an all-reduce of sums, which a set count is not
    if ((q.flags & EVQGPU_QUERY_PARTIAL) && ctx->nranks > 1 && s.dense_global)
      fail(EVQGPU_ERR_UNSUPPORTED, "count_distinct in the direct-addressed tier is not merged across ranks");
    if (q.dt_cap == 0) {
      q.dt_cap = next_pow2(std::max<uint64_t>(1ull << 20, std::min<uint64_t>(total_rows, 1ull << 24) * 2));
      if (const char* e = getenv("EVQGPU_DT_CAP")) q.dt_cap = next_pow2(std::max<uint64_t>(1024, strtoull(e, nullptr, 10)));   // (tests: force growth)
    }
    q.dt_slots.resize(q.distinct_args.size());
    for (size_t d = 0; d < q.distinct_args.size(); ++d) {
      ensure(q.dt_slots[d], q.dt_cap * 8 * 4);
      EVQ_CUDA(cudaMemsetAsync(q.dt_slots[d].p, 0, q.dt_cap * 8 * 4, ctx->stream));
      base.dt[d].slots = q.dt_slots[d].as<u64>();
      base.dt[d].cap = q.dt_cap;
      base.dt[d].stride = 4;      // fingerprint, group id, value, pad: one 32-byte sector
      base.dt[d].nkeys = 2;
    }
  }

  if (s.part_bits > 0) {
    const uint64_t parts = 1ull << s.part_bits;
    const uint64_t nrec = record_layout(s).nwords;
    base.part_bits = (u32) s.part_bits;
    u32 lg = 0;
    while ((1ull << lg) < base.ht.cap) ++lg;
    base.part_shift = lg - (u32) s.part_bits;
    cudaKernel_t agg = q.module->kernels.at("evq_agg_part");
    uint64_t seg_cap = 0;
    // every partition is one flat array of records; it takes its share of the table's rows (+ 5 % and a constant: the
    // counts are binomial around rows / partitions); more is an overflow (heavily skewed keys) -> fallback
    // a partition's capacity: its share of the rows + 25 % + 8 standard deviations of a binomial count + a constant (the
    // rows of one group go together: with few groups per partition the counts spread far more than independent rows would)
    auto cap_of = [](uint64_t rows, uint64_t nparts) {
      const uint64_t avg = rows / nparts + 1;
      return avg * 5 / 4 + 8 * (uint64_t) sqrt((double) avg) + 1024;
    };
    uint64_t table_rows = 0;
    auto before = [&](unsigned grid, uint64_t rows, EvqScanParams& P) {
      (void) grid;
      table_rows = rows;
      seg_cap = cap_of(rows, parts);
      ensure(q.part_buf, parts * seg_cap * nrec * 8 + 256);
      ensure(q.part_cursor, round_up(parts * 4, 256) + 256);   // (+ the pass-2 progress counter behind the partition cursors)
      EVQ_CUDA(cudaMemsetAsync(q.part_cursor.p, 0, round_up(parts * 4, 256) + 256, ctx->stream));
      P.part_buf = q.part_buf.as<u64>();
      P.part_cursor = q.part_cursor.as<u32>();
      P.part_cap = seg_cap;
    };
    auto after_smem = [&](unsigned grid) {
      (void) grid;
      // the slice of a sub-partition: slice_slots slots, or the whole first-level slice when that is smaller already
      u32 lgs = 0;
      while ((1u << lgs) < (u32) s.slice_slots) ++lgs;
      int sub_bits = (int) lg - (int) s.part_bits - (int) lgs;
      u32 slice_slots = (u32) s.slice_slots;
      if (sub_bits < 0) { sub_bits = 0; slice_slots = (u32) (base.ht.cap >> s.part_bits); }
      const uint64_t nsub_total = parts << sub_bits;
      const u64* buf = q.part_buf.as<u64>();
      const u32* cursor = q.part_cursor.as<u32>();
      uint64_t cap = seg_cap;
      if (sub_bits > 0) {
        cudaKernel_t rp = q.module->kernels.at("evq_repart");
        const uint64_t cap2 = cap_of(table_rows, nsub_total);
        ensure(q.part_buf2, nsub_total * cap2 * nrec * 8 + 256);
        ensure(q.part_cursor2, nsub_total * 4 + 256);
        EVQ_CUDA(cudaMemsetAsync(q.part_cursor2.p, 0, nsub_total * 4, ctx->stream));
        RepartParams rpp;
        memset(&rpp, 0, sizeof(rpp));
        rpp.ht = base.ht;
        rpp.in = buf;
        rpp.in_cursor = cursor;
        rpp.in_cap = seg_cap;
        rpp.out = q.part_buf2.as<u64>();
        rpp.out_cursor = q.part_cursor2.as<u32>();
        rpp.out_cap = cap2;
        rpp.status = base.status;
        rpp.nparts = (u32) parts;
        rpp.sub_bits = (u32) sub_bits;
        rpp.sub_shift = lgs;
        const size_t smem = 2048 * nrec * 8 + (3 * 256 + 8 + 260) * 4 + 2048 + 128;
        EVQ_CUDA(cudaFuncSetAttribute((const void*) rp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        int per_sm = 0;
        EVQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*) rp, 256, smem));
        void* args[] = {&rpp};
        launch(ctx, rp, dim3((unsigned) ctx->sm_count * (unsigned) std::max(1, per_sm)), dim3(256), smem, args);
        q.stats.kernel_launches++;
        buf = q.part_buf2.as<u64>();
        cursor = q.part_cursor2.as<u32>();
        cap = cap2;
      }
      cudaKernel_t ag = q.module->kernels.at("evq_agg_smem");
      AggSmemParams ap;
      memset(&ap, 0, sizeof(ap));
      ap.ht = base.ht;
      ap.buf = buf;
      ap.cursor = cursor;
      ap.cap = cap;
      ap.status = base.status;
      ap.nsub_total = (u32) nsub_total;
      ap.slice_slots = slice_slots;
      int threads = 1024;
      if (const char* e = getenv("EVQGPU_AGG_THREADS")) threads = atoi(e);
      const size_t smem = 2 * (size_t) slice_slots * (1 + nk + nstate) * 8 + 16;   // two buffers (the next slice is copied in meanwhile) + 2 mbarriers
      EVQ_CUDA(cudaFuncSetAttribute((const void*) ag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      int per_sm = 0;
      EVQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*) ag, threads, smem));
      void* args[] = {&ap};
      const unsigned g2 = (unsigned) std::min<uint64_t>(nsub_total, (uint64_t) ctx->sm_count * (uint64_t) std::max(1, per_sm));
      launch(ctx, ag, dim3(g2), dim3((unsigned) threads), smem, args);
      q.stats.kernel_launches++;
    };
    auto after = [&](unsigned grid) {
      if (s.slice_slots > 0) return after_smem(grid);
      AggParams ap;
      memset(&ap, 0, sizeof(ap));
      ap.ht = base.ht;
      ap.part_buf = q.part_buf.as<u64>();
      ap.part_cursor = q.part_cursor.as<u32>();
      ap.part_cap = seg_cap;
      ap.nparts = (u32) parts;
      ap.nseg = 0;
      ap.status = base.status;
      ap.counters = base.counters;
      ap.bar = (u32*) ((uint8_t*) q.part_cursor.p + round_up(parts * 4, 256));
      ap.window = 2;
      if (const char* e = getenv("EVQGPU_AGG_WINDOW")) ap.window = (u32) std::max(1, atoi(e));
      // all CTAs must be resident (they wait for each other): a cooperative launch of at most what the device holds
      int per_sm = 0;
      EVQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*) agg, 256, 0));
      unsigned g2 = (unsigned) ctx->sm_count * (unsigned) std::max(1, std::min(per_sm, 8));
      if (const char* e = getenv("EVQGPU_AGG_CTAS")) g2 = (unsigned) ctx->sm_count * (unsigned) std::max(1, std::min(per_sm, atoi(e)));
      void* args[] = {&ap};
      (void) grid;
      EVQ_CUDA(cudaLaunchCooperativeKernel((const void*) agg, dim3(g2), dim3(256), args, 0, ctx->stream));
      ctx->kernel_launches++;
      q.stats.kernel_launches++;
    };
    run_scan(q, s, plans, base, before, after);
  } else {
    run_scan(q, s, plans, base);
  }

  // ---- emit (deferred to evqgpu_query_merge for partial plans of a multi-rank job)
  q.emit_total_rows = total_rows;
  q.emit = EmitParams();
  memset(&q.emit, 0, sizeof(q.emit));
  q.emit.dense_state = base.dense_state;
  q.emit.ht = base.ht;
  for (size_t i = 0; i < nk; ++i) {
    q.emit.key_min[i] = dm.key_min[i];
    q.emit.key_stride[i] = dm.key_stride[i];
    q.emit.key_null_idx[i] = dm.key_null_idx[i];
    q.emit.key_range[i] = dm.key_range[i];
  }
  q.emit.slots = emit_slots;
  q.reordered = false;
  q.emitted = false;
  if (!((q.flags & EVQGPU_QUERY_PARTIAL) && ctx->nranks > 1)) {
    if (tail) launch_tail(q, false);
    else emit_results(q);
  }
  (void) sync;
}

static void execute_scan_only(evqgpu_query& q, std::vector<TablePlan>& plans, std::vector<evqgpu_table*>& tables) {
  evqgpu_ctx* ctx = q.ctx;
  KernelShape s = shape_of_plans(q, plans);
  uint64_t total_tiles = 0, total_rows = 0;
  for (auto* t : tables) { total_tiles += t->num_tiles; total_rows += t->num_rows; }
  s.tier = 0;
  fit_shape(q, s, plans);
  q.stats.strategy = 0;
  EvqScanParams base;
  memset(&base, 0, sizeof(base));
  base.status = q.status.as<u32>();
  base.counters = q.counters.as<u64>();
  EVQ_CUDA(cudaMemsetAsync(q.status.p, 0, 16, ctx->stream));
  EVQ_CUDA(cudaMemsetAsync(q.counters.p, 0, 32, ctx->stream));
  ensure(q.tile_counts, (total_tiles + 1) * 8);
  ensure(q.tile_base, (total_tiles + 1) * 8);
  EVQ_CUDA(cudaMemsetAsync(q.tile_counts.p, 0, (total_tiles + 1) * 8, ctx->stream));
  base.tile_counts = q.tile_counts.as<u64>();

  // pass 1: rows passing WHERE per tile
  q.shape = s;
  std::string src0 = generate_source(q, s);
  float ms = 0;
  q.module = jit_compile(ctx, src0, {"evq_scan"}, &ms);
  q.jit_ms_total += ms;
  run_scan(q, s, plans, base);

  // exclusive prefix -> output row of every tile's first passing row
  {
    size_t tmp_bytes = 0;
    EVQ_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, q.tile_counts.as<u64>(), q.tile_base.as<u64>(),
                                           (int64_t) (total_tiles + 1), ctx->stream));
    DevBuf tmp;
    tmp.alloc(tmp_bytes);
    EVQ_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, q.tile_counts.as<u64>(), q.tile_base.as<u64>(),
                                           (int64_t) (total_tiles + 1), ctx->stream));
    u64 total = 0;
    EVQ_CUDA(cudaMemcpyAsync(&total, q.tile_base.as<u64>() + total_tiles, 8, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->kernel_launches += 2;
    q.stats.kernel_launches += 2;
    q.num_rows_out = total;
  }
  // pass 2: projection of the passing rows at their final position
  q.out_cols.resize(q.select.size());
  for (size_t i = 0; i < q.select.size(); ++i)
    ensure(q.out_cols[i], q.num_rows_out * (q.select[i].expr->type == EVQ_BOOL ? 2 : 9) + 16);
  q.out_capacity = q.num_rows_out;
  s.tier = 3;
  q.shape = s;
  q.kernel_source = generate_source(q, s);
  q.module = jit_compile(ctx, q.kernel_source, {"evq_scan"}, &ms);
  q.jit_ms_total += ms;
  EVQ_CUDA(cudaMemsetAsync(q.counters.p, 0, 32, ctx->stream));
  base.tile_out_base = q.tile_base.as<u64>();
  for (size_t i = 0; i < q.select.size(); ++i) base.out_cols[i] = q.out_cols[i].as<u8>();
  run_scan(q, s, plans, base);
  (void) total_rows;
}

void finish_query(evqgpu_query& q) {
  evqgpu_ctx* ctx = q.ctx;
  use_device(ctx);
  struct { u32 status[4]; u64 counters[4]; u64 out_count; } host;
  if (q.use_tail && (q.flags & EVQGPU_QUERY_GROUPBY)) {
    // dense tier: one control block.  After the tail: the values it published ([8..14]); before it (a partial plan whose
    // merge has not been called yet): the live words, no rows yet
    u64 c[16];
    EVQ_CUDA(cudaMemcpyAsync(c, q.ctl.p, 128, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
    const u64* src = q.tail_done ? c + 8 : c;
    memcpy(host.status, src, 16);
    memcpy(host.counters, src + 2, 32);
    host.out_count = q.tail_done ? c[14] : 0;
  } else {
    EVQ_CUDA(cudaMemcpyAsync(host.status, q.status.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaMemcpyAsync(host.counters, q.counters.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaMemcpyAsync(&host.out_count, q.out_count.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  q.pending = false;
  // (accumulated over the finishes of one execution: the hash merge finishes the scan before it ships the table and
  // once more after the merged groups are emitted)
  q.stats.scan_launches += (uint32_t) q.prof_events.size();
  for (auto& e : q.prof_events) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess) q.stats.scan_ms += ms;
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  q.prof_events.clear();
  const u32 err = host.status[0];
  q.stats.rows_passed = host.counters[0];
  if (err & EVQ_ERR_DIV_ZERO) fail(EVQGPU_ERR_RUNTIME, "division by zero");
  if (err & EVQ_ERR_MOD_ZERO) fail(EVQGPU_ERR_RUNTIME, "modulo by zero");
  if (err & EVQ_ERR_STAGE_OVERFLOW) fail(EVQGPU_ERR_RUNTIME, "internal error: row tile larger than its pipeline stage");
  if (err & EVQ_ERR_SLOT_RANGE) fail(EVQGPU_ERR_RUNTIME, "internal error: group key outside the dense slot range");
  if (err & EVQ_ERR_PEER_TIMEOUT) fail(EVQGPU_ERR_RUNTIME, "merge: a peer rank did not deliver its partial aggregates (did every rank call evqgpu_query_merge?)");
  if ((q.flags & EVQGPU_QUERY_GROUPBY) && (err & EVQ_ERR_PART_FULL)) {
    // a record partition overflowed: the keys are heavily skewed, which the direct hash tier handles well (hot groups are
    // L2-resident anyway).  Run again without partitioning - a retry local to this rank, no collective.
    q.no_partition = true;
    std::vector<evqgpu_table*> tables = q.tables;
    int rc = evqgpu_query_enqueue(&q, tables.data(), (uint32_t) tables.size());
    if (rc == EVQGPU_OK) rc = evqgpu_query_finish(&q);
    if (rc != EVQGPU_OK) throw Error{rc, last_error()};
    return;
  }
  if (q.flags & EVQGPU_QUERY_GROUPBY) {
    if (err & EVQ_ERR_TABLE_FULL) {
      // grow the group table and run again (resize policy: double until it fits)
      if (q.ht_cap >= (1ull << 31)) fail(EVQGPU_ERR_NOMEM, "group table exceeds 2^31 slots");
      q.ht_cap *= 4;
      if (!q.distinct_args.empty()) {
        if (q.dt_cap >= (1ull << 33)) fail(EVQGPU_ERR_NOMEM, "count_distinct set exceeds 2^33 slots");
        q.dt_cap *= 4;
      }
      std::vector<evqgpu_table*> tables = q.tables;
      const uint64_t hint = q.ht_cap;
      (void) hint;
      // (enqueue + finish, not execute: the retry is local to this rank and must not enter a collective)
      int rc = evqgpu_query_enqueue(&q, tables.data(), (uint32_t) tables.size());
      if (rc == EVQGPU_OK) rc = evqgpu_query_finish(&q);
      if (rc != EVQGPU_OK) throw Error{rc, last_error()};
      return;
    }
    q.num_rows_out = std::min<uint64_t>(host.out_count, q.out_capacity);
    q.stats.num_groups = q.num_rows_out;
  } else {
    q.stats.num_groups = 0;
  }
}

}  // namespace evq

extern "C" {

int evqgpu_query_enqueue(evqgpu_query* q, evqgpu_table* const* tables, uint32_t ntables) {
  return guarded([&] {
    if (!q || (!tables && ntables)) fail(EVQGPU_ERR_ARG, "evqgpu_query_enqueue: null argument");
    if (ntables == 0) fail(EVQGPU_ERR_ARG, "evqgpu_query_enqueue: no tables");
    if (q->coordinator) fail(EVQGPU_ERR_ARG, "a coordinator query scans nothing: feed it with evqgpu_query_merge_rows");
    use_device(q->ctx);
    q->tables.assign(tables, tables + ntables);
    const uint64_t keep_cap = q->ht_cap;
    q->stats = evqgpu_query_stats();
    q->jit_ms_total = 0;
    q->num_rows_out = 0;
    q->merged = false;
    std::vector<TablePlan> plans;
    std::vector<evqgpu_table*> tv(tables, tables + ntables);
    for (auto* t : tv) {
      if (!t) fail(EVQGPU_ERR_ARG, "null table");
      if (t->ctx != q->ctx) fail(EVQGPU_ERR_ARG, "table belongs to another context");
      TablePlan p;
      p.table = t;
      p.binding = bind_table(*q, t);
      plans.push_back(std::move(p));
      q->stats.rows_scanned += t->num_rows;
    }
    q->ht_cap = keep_cap;
    q->stats.algorithmic_bytes = algorithmic_bytes(*q, plans);
    if (q->flags & EVQGPU_QUERY_GROUPBY) execute_groupby(*q, plans, tv, false);
    else execute_scan_only(*q, plans, tv);
    q->stats.jit_ms = q->jit_ms_total;
    q->pending = true;
  });
}

int evqgpu_query_finish(evqgpu_query* q) {
  return guarded([&] {
    if (!q) fail(EVQGPU_ERR_ARG, "evqgpu_query_finish: null query");
    if (q->pending) finish_query(*q);
  });
}

int evqgpu_query_prepare(evqgpu_query* q, evqgpu_table* const* tables, uint32_t ntables) {
  return guarded([&] {
    if (!q || (!tables && ntables)) fail(EVQGPU_ERR_ARG, "evqgpu_query_prepare: null argument");
    use_device(q->ctx);
    std::vector<evqgpu_table*> tv(tables, tables + ntables);
    prepare_query(*q, tv);
  });
}

int evqgpu_query_execute(evqgpu_query* q, evqgpu_table* const* tables, uint32_t ntables) {
  // in a multi-rank job execute is collective: the ranks agree on slot assignment and state layout first
  int rc = evqgpu_query_prepare(q, tables, ntables);
  if (rc != EVQGPU_OK) return rc;
  rc = evqgpu_query_enqueue(q, tables, ntables);
  if (rc != EVQGPU_OK) return rc;
  return evqgpu_query_finish(q);
}

int evqgpu_query_num_rows(evqgpu_query* q, uint64_t* out) {
  return guarded([&] {
    if (!q || !out) fail(EVQGPU_ERR_ARG, "evqgpu_query_num_rows: null argument");
    if (q->pending) finish_query(*q);
    *out = q->num_rows_out;
  });
}

int evqgpu_query_fetch(evqgpu_query* q, uint64_t row0, uint64_t max_rows, void* const* columns, uint64_t* nrows_out) {
  return guarded([&] {
    if (!q || !columns || !nrows_out) fail(EVQGPU_ERR_ARG, "evqgpu_query_fetch: null argument");
    if (q->pending) finish_query(*q);
    use_device(q->ctx);
    uint64_t n = 0;
    if (row0 < q->num_rows_out) n = std::min<uint64_t>(max_rows, q->num_rows_out - row0);
    for (size_t i = 0; n && i < q->select.size(); ++i) {
      const uint64_t w = q->select[i].expr->type == EVQ_BOOL ? 2 : 9;
      if (!columns[i]) fail(EVQGPU_ERR_ARG, "evqgpu_query_fetch: null column buffer");
      EVQ_CUDA(cudaMemcpyAsync(columns[i], q->out_cols[i].as<u8>() + row0 * w, n * w, cudaMemcpyDeviceToHost, q->ctx->stream));
    }
    if (n) EVQ_CUDA(cudaStreamSynchronize(q->ctx->stream));
    *nrows_out = n;
  });
}

int evqgpu_query_fetch_strings(evqgpu_query* q, uint32_t column, uint64_t row0, uint64_t max_rows, void* dst, uint64_t cap,
                               uint64_t* nrows_out, uint64_t* nbytes_out) {
  return guarded([&] {
    if (!q || !nrows_out || !nbytes_out) fail(EVQGPU_ERR_ARG, "evqgpu_query_fetch_strings: null argument");
    if (column >= q->select.size() || !q->select[column].is_string)
      fail(EVQGPU_ERR_ARG, "evqgpu_query_fetch_strings: result column %u is not a string column", column);
    if (q->pending) finish_query(*q);
    use_device(q->ctx);
    uint64_t n = 0;
    if (row0 < q->num_rows_out) n = std::min<uint64_t>(max_rows, q->num_rows_out - row0);
    *nrows_out = n;
    *nbytes_out = 0;
    if (!n) return;
    // the device result holds [code u64][tag]; the values come from the context's dictionary (result rows only)
    std::vector<uint8_t> packed(n * 9);
    EVQ_CUDA(cudaMemcpyAsync(packed.data(), q->out_cols[column].as<u8>() + row0 * 9, n * 9, cudaMemcpyDeviceToHost, q->ctx->stream));
    EVQ_CUDA(cudaStreamSynchronize(q->ctx->stream));
    const auto& dict = q->ctx->code_strings;
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; ++i) {
      uint64_t code;
      memcpy(&code, &packed[i * 9], 8);
      if (code >= dict.size()) fail(EVQGPU_ERR_RUNTIME, "string code %llu outside the dictionary", (unsigned long long) code);
      total += 5 + ((packed[i * 9 + 8] & EVQ_STAG_NULL) ? 0 : dict[code].size());
    }
    *nbytes_out = total;
    if (!dst || cap < total) { *nrows_out = 0; return; }
    uint8_t* o = (uint8_t*) dst;
    for (uint64_t i = 0; i < n; ++i) {   // the packed STRING element: [u32 length][bytes][tag] (svalue.cc:533-549)
      uint64_t code;
      memcpy(&code, &packed[i * 9], 8);
      const bool null = packed[i * 9 + 8] & EVQ_STAG_NULL;
      const uint32_t len = null ? 0u : (uint32_t) dict[code].size();
      memcpy(o, &len, 4);
      if (len) memcpy(o + 4, dict[code].data(), len);
      o[4 + len] = null ? EVQ_STAG_NULL : 0;
      o += 5 + len;
    }
  });
}

int evqgpu_query_get_stats(evqgpu_query* q, evqgpu_query_stats* out) {
  return guarded([&] {
    if (!q || !out) fail(EVQGPU_ERR_ARG, "evqgpu_query_get_stats: null argument");
    if (q->pending) finish_query(*q);
    *out = q->stats;
  });
}

const char* evqgpu_query_kernel_source(evqgpu_query* q) { return q ? q->kernel_source.c_str() : ""; }

}  // extern "C"
