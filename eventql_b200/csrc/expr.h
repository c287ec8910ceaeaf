// expr.h - expression trees rebuilt from the postfix evqgpu_insn programs, the function table, and the
// translation of a typed expression into CUDA C (the device-side replacement of the csql stack VM,
// sql/runtime/vm.cc:107-157, specialised per query instead of interpreted per row).
#pragma once
#include <memory>
#include <string>
#include <vector>
#include "util.h"

namespace evq {

enum class Fn {
  // pure functions (sql/expressions/{boolean,math,conversion,datetime}.cc)
  LOGICAL_AND, LOGICAL_OR, NEG, CMP, EQ, NEQ, LT, LTE, GT, GTE,
  STARTSWITH, ENDSWITH,   // (string predicates: evaluated once per dictionary entry, query.cu lower_strings)
  ADD, SUB, MUL, DIV, MOD, POW,
  TO_NIL, TO_INT64, TO_TIMESTAMP64, FROM_TIMESTAMP, DATE_TRUNC,
  // aggregates (sql/expressions/aggregate.cc + oracle/ref_tools/ext_aggregates.cc)
  COUNT, SUM, MIN, MAX, MEAN, COUNT_DISTINCT
};

struct FnInfo {
  std::string symbol;     // name#ret/arg;arg;
  Fn fn;
  int ret;                // SType
  std::vector<int> args;  // STypes
  bool aggregate;
};

const std::vector<FnInfo>& function_table();
int function_lookup(const std::string& symbol);

struct Expr {
  int op = 0;         // EVQ_X_*
  int type = 0;       // result SType
  int fn = -1;        // index into function_table() for EVQ_X_CALL
  uint32_t col = 0;   // EVQ_X_INPUT
  uint64_t imm = 0;   // EVQ_X_LITERAL
  std::string str;    // string literal
  std::vector<std::unique_ptr<Expr>> args;

  bool is_aggregate_call() const;
  const FnInfo& info() const { return function_table()[fn]; }
  std::unique_ptr<Expr> clone() const;
  // canonical text, used for structural equality (scalar select items vs GROUP BY expressions) and cache keys
  std::string signature() const;
};

using ExprPtr = std::unique_ptr<Expr>;

// rebuild the tree from a postfix program; throws EVQGPU_ERR_ARG on malformed programs
ExprPtr parse_program(const evqgpu_expr& prog);

// first aggregate call, depth first (QueryTreeUtil::findAggregateExpression, qtree/QueryTreeUtil.cc:209-224)
const Expr* find_aggregate(const Expr* e);
void collect_columns(const Expr* e, std::vector<bool>& used);

// ---- code generation ----
// How a leaf is spelled in the generated code.
struct CodegenEnv {
  // value / tag text of input column i (nullptr-like empty => column references are illegal here)
  std::vector<std::string> col_value, col_tag;
  // substitution of whole subtrees by signature (group key i -> stored key value; the aggregate call -> its result)
  std::vector<std::pair<std::string, std::pair<std::string, std::string>>> subst;   // signature -> (value, tag)
  std::string err = "err";   // name of the u32 error accumulator in scope
  // upper bound of the values of input column i (column statistics); empty = unknown
  std::vector<uint64_t> col_max;
  // lower bound of the values of input column i; empty = 0
  std::vector<uint64_t> col_min;
};

struct Code {
  std::string value;   // C expression of the value in its natural C type (u64, i64, f64, bool as u32 0/1)
  std::string tag;     // C expression of the STag (0 / 1)
};

Code gen_expr(const Expr* e, const CodegenEnv& env);
uint32_t expr_value_bits(const Expr* e, const CodegenEnv& env);   // bit length bound of a uint64-valued expression (64 = unknown)
uint64_t expr_value_max(const Expr* e, const CodegenEnv& env);
bool expr_may_raise(const Expr* e);   // integer division / modulo / pow somewhere in the tree
// the value as raw 64-bit pattern (u64), e.g. for key tuples and aggregate state words
std::string as_bits(const Code& c, int type);
const char* ctype_of(int type);

uint64_t date_trunc_window(const std::string& w);   // datetime.cc:58-84,115-137

}  // namespace evq
