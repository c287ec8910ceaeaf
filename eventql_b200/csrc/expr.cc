#include "expr.h"
#include <string.h>
#include <map>
#include <sstream>

namespace evq {

static const char* kTypeNames[] = {"nil", "uint64", "int64", "float64", "bool", "string", "timestamp64"};

static std::string make_symbol(const std::string& name, int ret, const std::vector<int>& args) {
  std::string s = name + "#" + kTypeNames[ret] + "/";
  for (int a : args) { s += kTypeNames[a]; s += ";"; }
  return s;
}

// The registry follows sql/defaults.cc:38-171 for the functions the device path implements, plus the typed
// extension aggregates (min/max/mean/sum<float64>).  String functions and now()/time_at()/date_add are absent:
// looking them up yields -1 and a plan using them is refused (EVQGPU_ERR_UNSUPPORTED), never run on the CPU.
const std::vector<FnInfo>& function_table() {
  static const std::vector<FnInfo> table = [] {
    std::vector<FnInfo> t;
    auto reg = [&](const char* name, Fn fn, int ret, std::vector<int> args, bool agg = false) {
      t.push_back({make_symbol(name, ret, args), fn, ret, args, agg});
    };
    const int U = EVQ_UINT64, I = EVQ_INT64, F = EVQ_FLOAT64, B = EVQ_BOOL, T = EVQ_TIMESTAMP64, N = EVQ_NIL,
              S = EVQ_STRING;
    reg("count", Fn::COUNT, U, {N}, true);
    reg("sum", Fn::SUM, I, {I}, true);
    reg("sum", Fn::SUM, U, {U}, true);
    reg("sum", Fn::SUM, F, {F}, true);
    reg("count_distinct", Fn::COUNT_DISTINCT, U, {U}, true);   // sql/defaults.cc:50, aggregate.cc:80-137
    for (int ty : {U, I, F}) {
      reg("min", Fn::MIN, ty, {ty}, true);
      reg("max", Fn::MAX, ty, {ty}, true);
      reg("mean", Fn::MEAN, F, {ty}, true);
    }
    reg("logical_and", Fn::LOGICAL_AND, B, {B, B});
    reg("logical_or", Fn::LOGICAL_OR, B, {B, B});
    reg("neg", Fn::NEG, B, {B});
    for (int ty : {U, I, F, T}) reg("cmp", Fn::CMP, I, {ty, ty});
    for (int ty : {U, I, F, B, S, T}) {   // string eq / neq: lowered to dictionary codes at intake (query.cu: lower_strings)
      reg("eq", Fn::EQ, B, {ty, ty});
      reg("neq", Fn::NEQ, B, {ty, ty});
    }
    // string ordering comparisons (boolean.cc:439-710) and startswith / endswith (expressions/string.cc:52-74) between a
    // string column and a literal: evaluated once per dictionary entry, the rows read the verdict of their value's code
    reg("startswith", Fn::STARTSWITH, B, {S, S});
    reg("endswith", Fn::ENDSWITH, B, {S, S});
    for (int ty : {U, I, F, T, S}) {
      reg("lt", Fn::LT, B, {ty, ty});
      reg("lte", Fn::LTE, B, {ty, ty});
      reg("gt", Fn::GT, B, {ty, ty});
      reg("gte", Fn::GTE, B, {ty, ty});
    }
    for (int ty : {U, I, F, B, T}) reg("to_nil", Fn::TO_NIL, N, {ty});
    for (int ty : {U, F, B, T}) reg("to_int64", Fn::TO_INT64, I, {ty});
    reg("to_timestamp64", Fn::TO_TIMESTAMP64, T, {I});
    reg("to_timestamp64", Fn::TO_TIMESTAMP64, T, {F});
    reg("from_timestamp", Fn::FROM_TIMESTAMP, T, {I});
    reg("from_timestamp", Fn::FROM_TIMESTAMP, T, {F});
    reg("date_trunc", Fn::DATE_TRUNC, T, {S, T});
    for (int ty : {U, I, F}) {
      reg("add", Fn::ADD, ty, {ty, ty});
      reg("sub", Fn::SUB, ty, {ty, ty});
      reg("mul", Fn::MUL, ty, {ty, ty});
      reg("div", Fn::DIV, ty, {ty, ty});
      reg("mod", Fn::MOD, ty, {ty, ty});
      reg("pow", Fn::POW, ty, {ty, ty});
    }
    return t;
  }();
  return table;
}

int function_lookup(const std::string& symbol) {
  static const std::map<std::string, int> index = [] {
    std::map<std::string, int> m;
    const auto& t = function_table();
    for (size_t i = 0; i < t.size(); ++i) m[t[i].symbol] = (int) i;
    return m;
  }();
  std::string lower = symbol;
  for (auto& ch : lower) ch = (char) tolower(ch);
  auto it = index.find(lower);
  return it == index.end() ? -1 : it->second;
}

bool Expr::is_aggregate_call() const { return op == EVQ_X_CALL && fn >= 0 && info().aggregate; }

ExprPtr Expr::clone() const {
  ExprPtr e(new Expr());
  e->op = op; e->type = type; e->fn = fn; e->col = col; e->imm = imm; e->str = str;
  for (const auto& a : args) e->args.push_back(a->clone());
  return e;
}

std::string Expr::signature() const {
  std::ostringstream os;
  switch (op) {
    case EVQ_X_INPUT: os << "c" << col << ":" << type; break;
    case EVQ_X_LITERAL:
      if (type == EVQ_STRING) os << "s'" << str << "'";
      else os << "l" << type << ":" << imm;
      break;
    case EVQ_X_CALL:
      os << info().symbol << "(";
      for (const auto& a : args) os << a->signature() << ",";
      os << ")";
      break;
    case EVQ_X_IF:
      os << "if(" << args[0]->signature() << "," << args[1]->signature() << "," << args[2]->signature() << ")";
      break;
  }
  return os.str();
}

ExprPtr parse_program(const evqgpu_expr& prog) {
  if (prog.len == 0) return nullptr;
  if (!prog.code) fail(EVQGPU_ERR_ARG, "expression program has no code");
  std::vector<ExprPtr> stack;
  for (uint32_t i = 0; i < prog.len; ++i) {
    const evqgpu_insn& in = prog.code[i];
    ExprPtr e(new Expr());
    e->op = in.op;
    e->type = in.type;
    if (in.type > EVQ_TIMESTAMP64) fail(EVQGPU_ERR_ARG, "expression: invalid type %u", in.type);
    switch (in.op) {
      case EVQ_X_INPUT:
        e->col = in.arg;
        break;
      case EVQ_X_LITERAL:
        e->imm = in.imm;
        if (in.type == EVQ_STRING) {
          const uint32_t off = (uint32_t) (in.imm >> 32), len = (uint32_t) in.imm;
          if (len && (!prog.strings || (uint64_t) off + len > prog.strings_len))
            fail(EVQGPU_ERR_ARG, "expression: string literal out of range");
          if (len) e->str.assign(prog.strings + off, len);
        }
        break;
      case EVQ_X_CALL: {
        const auto& table = function_table();
        if (in.arg >= table.size()) fail(EVQGPU_ERR_ARG, "expression: unknown function id %u", in.arg);
        e->fn = (int) in.arg;
        const FnInfo& fi = table[in.arg];
        if (in.nargs != fi.args.size() || stack.size() < in.nargs)
          fail(EVQGPU_ERR_ARG, "expression: wrong argument count for %s", fi.symbol.c_str());
        e->args.resize(in.nargs);
        for (int k = (int) in.nargs - 1; k >= 0; --k) {
          e->args[k] = std::move(stack.back());
          stack.pop_back();
        }
        for (size_t k = 0; k < fi.args.size(); ++k) {
          const int have = e->args[k]->type, want = fi.args[k];
          const bool same = have == want || (want == EVQ_NIL);
          if (!same) fail(EVQGPU_ERR_ARG, "expression: argument %zu of %s has type %s", k, fi.symbol.c_str(), kTypeNames[have]);
        }
        e->type = fi.ret;
        break;
      }
      case EVQ_X_IF: {
        if (stack.size() < 3) fail(EVQGPU_ERR_ARG, "expression: if needs three operands");
        e->args.resize(3);
        for (int k = 2; k >= 0; --k) {
          e->args[k] = std::move(stack.back());
          stack.pop_back();
        }
        if (e->args[0]->type != EVQ_BOOL) fail(EVQGPU_ERR_ARG, "expression: if condition must be bool");
        if (e->args[1]->type != e->args[2]->type) fail(EVQGPU_ERR_ARG, "expression: if branches differ in type");
        e->type = e->args[1]->type;
        break;
      }
      default:
        fail(EVQGPU_ERR_ARG, "expression: unknown opcode %u", in.op);
    }
    stack.push_back(std::move(e));
  }
  if (stack.size() != 1) fail(EVQGPU_ERR_ARG, "expression: program leaves %zu values on the stack", stack.size());
  return std::move(stack.back());
}

const Expr* find_aggregate(const Expr* e) {
  if (!e) return nullptr;
  if (e->is_aggregate_call()) return e;
  for (const auto& a : e->args) {
    const Expr* r = find_aggregate(a.get());
    if (r) return r;
  }
  return nullptr;
}

void collect_columns(const Expr* e, std::vector<bool>& used) {
  if (!e) return;
  if (e->op == EVQ_X_INPUT) {
    if (e->col >= used.size()) fail(EVQGPU_ERR_ARG, "expression references input column %u of %zu", e->col, used.size());
    used[e->col] = true;
  }
  for (const auto& a : e->args) collect_columns(a.get(), used);
}

uint64_t date_trunc_window(const std::string& w) {
  static const std::map<std::string, uint64_t> units = {
      {"ms", 1000ull}, {"msec", 1000ull}, {"millisecond", 1000ull}, {"milliseconds", 1000ull},
      {"s", 1000000ull}, {"sec", 1000000ull}, {"second", 1000000ull}, {"seconds", 1000000ull},
      {"min", 60000000ull}, {"minute", 60000000ull}, {"minutes", 60000000ull},
      {"h", 3600000000ull}, {"hour", 3600000000ull}, {"hours", 3600000000ull},
      {"d", 86400000000ull}, {"day", 86400000000ull}, {"days", 86400000000ull},
      {"w", 604800000000ull}, {"week", 604800000000ull}, {"weeks", 604800000000ull},
      {"month", 2592000000000ull}, {"months", 2592000000000ull},
      {"y", 31536000000000ull}, {"year", 31536000000000ull}, {"years", 31536000000000ull}};
  size_t i = 0;
  while (i < w.size() && isspace((unsigned char) w[i])) ++i;
  size_t j = i;
  while (j < w.size() && isdigit((unsigned char) w[j])) ++j;
  uint64_t mult = 1;
  std::string unit = w;
  if (j > i) {
    mult = strtoull(w.substr(i, j - i).c_str(), nullptr, 10);
    unit = w.substr(j);
  }
  auto it = units.find(unit);
  if (it == units.end()) fail(EVQGPU_ERR_RUNTIME, "unknown time window %s", w.c_str());
  return it->second * mult;
}

const char* ctype_of(int type) {
  switch (type) {
    case EVQ_INT64: return "i64";
    case EVQ_FLOAT64: return "f64";
    case EVQ_BOOL: return "u32";
    default: return "u64";
  }
}

std::string as_bits(const Code& c, int type) {
  switch (type) {
    case EVQ_FLOAT64: return "evq_bits(" + c.value + ")";
    case EVQ_NIL: return "0ull";
    default: return "((u64) (" + c.value + "))";
  }
}

static std::string hex64(uint64_t v) {
  char buf[32];
  snprintf(buf, sizeof(buf), "0x%llxull", (unsigned long long) v);
  return buf;
}

// Value range [lo, hi] of a uint64-valued expression given the upper bounds of the input columns (column statistics).
// {0, ~0} = unknown / may wrap.
struct Range { uint64_t lo, hi; };
static const Range kAny = {0, ~0ull};

static Range expr_range(const Expr* e, const CodegenEnv& env) {
  switch (e->op) {
    case EVQ_X_INPUT:
      if (e->type == EVQ_BOOL) return {0, 1};
      if (e->type != EVQ_UINT64 && e->type != EVQ_TIMESTAMP64) return kAny;
      return {e->col < env.col_min.size() ? env.col_min[e->col] : 0ull, e->col < env.col_max.size() ? env.col_max[e->col] : ~0ull};
    case EVQ_X_LITERAL:
      if (e->type == EVQ_BOOL) return {e->imm ? 1ull : 0ull, e->imm ? 1ull : 0ull};
      if (e->type != EVQ_UINT64 && e->type != EVQ_TIMESTAMP64) return kAny;
      return {e->imm, e->imm};
    case EVQ_X_IF: {
      const Range a = expr_range(e->args[1].get(), env), b = expr_range(e->args[2].get(), env);
      return {std::min(a.lo, b.lo), std::max(a.hi, b.hi)};
    }
    case EVQ_X_CALL: break;
    default: return kAny;
  }
  const FnInfo& fi = e->info();
  if (fi.ret == EVQ_BOOL) return {0, 1};
  if (fi.ret != EVQ_UINT64 && fi.ret != EVQ_TIMESTAMP64) return kAny;
  if (fi.args.empty() || (fi.args[0] != EVQ_UINT64 && fi.args[0] != EVQ_TIMESTAMP64)) return kAny;
  if (fi.fn == Fn::DATE_TRUNC) return {0, expr_range(e->args[1].get(), env).hi};
  if (e->args.size() != 2) return kAny;
  const Range a = expr_range(e->args[0].get(), env), b = expr_range(e->args[1].get(), env);
  switch (fi.fn) {
    case Fn::ADD:
      if (a.hi > ~0ull - b.hi) return kAny;
      return {a.lo + b.lo, a.hi + b.hi};
    case Fn::SUB:
      if (a.lo < b.hi) return kAny;   // may wrap below zero
      return {a.lo - b.hi, a.hi - b.lo};
    case Fn::MUL:
      if (b.hi != 0 && a.hi > ~0ull / b.hi) return kAny;
      return {a.lo * b.lo, a.hi * b.hi};
    case Fn::DIV: return {0, a.hi};
    case Fn::MOD: return {0, b.hi ? std::min(a.hi, b.hi - 1) : a.hi};
    default: return kAny;
  }
}

uint32_t expr_value_bits(const Expr* e, const CodegenEnv& env) {
  uint32_t b = 0;
  for (uint64_t v = expr_range(e, env).hi; v; v >>= 1) ++b;
  return b ? b : 1;
}

uint64_t expr_value_max(const Expr* e, const CodegenEnv& env) { return expr_range(e, env).hi; }

static Code gen_expr_impl(const Expr* e, const CodegenEnv& env);

// Can evaluating `e` raise (integer division / modulo by zero, math.cc:136-216)?  Such subtrees are never folded away.
bool expr_may_raise(const Expr* e) {
  if (e->op == EVQ_X_CALL) {
    const Fn fn = e->info().fn;
    if (fn == Fn::DIV || fn == Fn::MOD || fn == Fn::POW) return true;
  }
  for (const auto& a : e->args)
    if (expr_may_raise(a.get())) return true;
  return false;
}

// A comparison of two uint64-valued expressions whose value ranges (column statistics) decide it: 1 / 0, else -1.
static int fold_compare(Fn fn, const Expr* x, const Expr* y, const CodegenEnv& env) {
  if (env.col_max.empty() || expr_may_raise(x) || expr_may_raise(y)) return -1;
  const Range a = expr_range(x, env), b = expr_range(y, env);
  switch (fn) {
    case Fn::GT: return a.lo > b.hi ? 1 : a.hi <= b.lo ? 0 : -1;
    case Fn::GTE: return a.lo >= b.hi ? 1 : a.hi < b.lo ? 0 : -1;
    case Fn::LT: return a.hi < b.lo ? 1 : a.lo >= b.hi ? 0 : -1;
    case Fn::LTE: return a.hi <= b.lo ? 1 : a.lo > b.hi ? 0 : -1;
    case Fn::EQ: return (a.hi < b.lo || a.lo > b.hi) ? 0 : (a.lo == a.hi && b.lo == b.hi && a.lo == b.lo) ? 1 : -1;
    case Fn::NEQ: return (a.hi < b.lo || a.lo > b.hi) ? 1 : (a.lo == a.hi && b.lo == b.hi && a.lo == b.lo) ? 0 : -1;
    default: return -1;
  }
}

// Sub-expressions whose value provably fits 32 bits are re-typed through u32: same value, but the compiler can then use
// 32-bit compares and 32x32->64 / 64x32 multiplies instead of full 64-bit arithmetic.
Code gen_expr(const Expr* e, const CodegenEnv& env) {
  Code c = gen_expr_impl(e, env);
  if (!env.col_max.empty() && e->op == EVQ_X_CALL && (e->type == EVQ_UINT64 || e->type == EVQ_TIMESTAMP64) &&
      !e->info().aggregate && expr_value_bits(e, env) <= 32)
    c.value = "((u64) (u32) (" + c.value + "))";
  return c;
}

static Code gen_expr_impl(const Expr* e, const CodegenEnv& env) {
  const std::string sig = e->signature();
  for (const auto& s : env.subst)
    if (s.first == sig) return {s.second.first, s.second.second};
  switch (e->op) {
    case EVQ_X_INPUT: {
      if (e->col >= env.col_value.size() || env.col_value[e->col].empty())
        fail(EVQGPU_ERR_UNSUPPORTED,
             "a column is referenced where only GROUP BY expressions and aggregates are available "
             "(non-aggregate select items must be functions of the GROUP BY key)");
      return {env.col_value[e->col], env.col_tag[e->col]};
    }
    case EVQ_X_LITERAL:
      switch (e->type) {
        case EVQ_FLOAT64: return {"evq_f64(" + hex64(e->imm) + ")", "0u"};
        case EVQ_INT64: return {"((i64) " + hex64(e->imm) + ")", "0u"};
        case EVQ_BOOL: return {e->imm ? "1u" : "0u", "0u"};
        case EVQ_NIL: return {"0ull", "1u"};
        case EVQ_STRING: fail(EVQGPU_ERR_UNSUPPORTED, "string values are outside the numeric device path");
        default: return {hex64(e->imm), "0u"};
      }
    case EVQ_X_IF: {
      Code c = gen_expr(e->args[0].get(), env), t = gen_expr(e->args[1].get(), env), f = gen_expr(e->args[2].get(), env);
      Code r;
      r.value = "((" + c.value + ") ? (" + t.value + ") : (" + f.value + "))";
      if (t.tag == "0u" && f.tag == "0u") r.tag = "0u";
      else r.tag = "((" + c.value + ") ? (" + t.tag + ") : (" + f.tag + "))";
      return r;
    }
    case EVQ_X_CALL: break;
    default: fail(EVQGPU_ERR_ARG, "bad expression node");
  }
  const FnInfo& fi = e->info();
  if (fi.aggregate) fail(EVQGPU_ERR_UNSUPPORTED, "aggregate call %s in a pure context (one aggregate per select item, SURVEY H6)", fi.symbol.c_str());
  if (fi.fn == Fn::DATE_TRUNC) {
    const Expr* w = e->args[0].get();
    if (w->op != EVQ_X_LITERAL || w->type != EVQ_STRING)
      fail(EVQGPU_ERR_UNSUPPORTED, "date_trunc: the window must be a string literal");
    const uint64_t t = date_trunc_window(w->str);
    Code ts = gen_expr(e->args[1].get(), env);
    return {"(((u64) (" + ts.value + ")) / " + hex64(t) + " * " + hex64(t) + ")", "0u"};
  }
  std::vector<Code> a;
  for (const auto& x : e->args) a.push_back(gen_expr(x.get(), env));
  const int T = fi.args.empty() ? EVQ_NIL : fi.args[0];
  if ((T == EVQ_UINT64 || T == EVQ_TIMESTAMP64) && e->args.size() == 2) {
    // range checks the statistics of the scanned columns already answer (e.g. `price > 0` over a column whose minimum is 90000)
    const int folded = fold_compare(fi.fn, e->args[0].get(), e->args[1].get(), env);
    if (folded >= 0) return {folded ? "1u" : "0u", "0u"};
  }
  auto bin = [&](const char* op) { return "((" + a[0].value + ") " + op + " (" + a[1].value + "))"; };
  std::string v;
  switch (fi.fn) {
    // boolean.cc:38-77 - both operands are always evaluated (no short circuit): '&' / '|' on 0/1 values
    case Fn::LOGICAL_AND: v = "((u32) ((" + a[0].value + ") != 0) & (u32) ((" + a[1].value + ") != 0))"; break;
    case Fn::LOGICAL_OR: v = "((u32) ((" + a[0].value + ") != 0) | (u32) ((" + a[1].value + ") != 0))"; break;
    case Fn::NEG: v = "((u32) !(" + a[0].value + "))"; break;
    case Fn::CMP: v = "((i64) ((" + a[0].value + ") > (" + a[1].value + ")) - (i64) ((" + a[0].value + ") < (" + a[1].value + ")))"; break;
    case Fn::EQ: v = "((u32) " + bin("==") + ")"; break;
    case Fn::NEQ: v = "((u32) " + bin("!=") + ")"; break;
    case Fn::LT: v = "((u32) " + bin("<") + ")"; break;
    case Fn::LTE: v = "((u32) " + bin("<=") + ")"; break;
    case Fn::GT: v = "((u32) " + bin(">") + ")"; break;
    case Fn::GTE: v = "((u32) " + bin(">=") + ")"; break;
    case Fn::ADD:
    case Fn::SUB:
    case Fn::MUL: {
      const char* op = fi.fn == Fn::ADD ? "+" : fi.fn == Fn::SUB ? "-" : "*";
      if (T == EVQ_INT64) v = "((i64) ((u64) (" + a[0].value + ") " + op + " (u64) (" + a[1].value + ")))";   // wraps like the reference's -O2 build
      else v = bin(op);
      break;
    }
    case Fn::DIV:
      if (T == EVQ_FLOAT64) v = bin("/");
      else v = std::string(T == EVQ_INT64 ? "evq_div_i64(" : "evq_div_u64(") + a[0].value + ", " + a[1].value + ", " + env.err + ")";
      break;
    case Fn::MOD:
      if (T == EVQ_FLOAT64) v = "fmod(" + a[0].value + ", " + a[1].value + ")";
      else v = std::string(T == EVQ_INT64 ? "evq_mod_i64(" : "evq_mod_u64(") + a[0].value + ", " + a[1].value + ", " + env.err + ")";
      break;
    case Fn::POW:
      if (T == EVQ_FLOAT64) v = "pow(" + a[0].value + ", " + a[1].value + ")";
      else v = std::string("((") + ctype_of(T) + ") pow((f64) (" + a[0].value + "), (f64) (" + a[1].value + ")))";
      break;
    case Fn::TO_NIL: v = "((void) (" + a[0].value + "), 0ull)"; break;
    case Fn::TO_INT64: v = "((i64) (" + a[0].value + "))"; break;
    case Fn::TO_TIMESTAMP64: v = "((u64) (" + a[0].value + "))"; break;
    case Fn::FROM_TIMESTAMP:
      if (T == EVQ_FLOAT64) v = "((u64) ((" + a[0].value + ") * 1000000.0))";
      else v = "((u64) ((i64) ((u64) (" + a[0].value + ") * 1000000ull)))";
      break;
    default: fail(EVQGPU_ERR_UNSUPPORTED, "function %s is not implemented on the device path", fi.symbol.c_str());
  }
  // every pure function pushes tag 0 (sql/svalue.cc:950-958)
  return {v, "0u"};
}

}  // namespace evq

// ---- C ABI ----
extern "C" {

int evqgpu_function_lookup(const char* symbol) { return symbol ? evq::function_lookup(symbol) : -1; }

const char* evqgpu_function_symbol(int id) {
  const auto& t = evq::function_table();
  if (id < 0 || (size_t) id >= t.size()) return nullptr;
  return t[id].symbol.c_str();
}

int evqgpu_function_is_aggregate(int id) {
  const auto& t = evq::function_table();
  if (id < 0 || (size_t) id >= t.size()) return 0;
  return t[id].aggregate ? 1 : 0;
}

}  // extern "C"
