// csql_mirror.h - the slice of the reference's csql operator surface the GPU path plugs into, restated as a
// dependency-free header (same names, argument meaning and error behaviour; the reference's own headers pull in its
// whole util/ + protobuf tree and are not vendored here).
//
//   csql::SType / STag / SVector          sql/svalue.h:41-162, svalue.cc:410-549   (packed, unaligned element encoding)
//   csql::ReturnCode                      util/return_code.h
//   csql::TableExpression                 sql/table_expression.h:35-50
//   csql::ValueExpressionNode family      sql/qtree/{ColumnReferenceNode,LiteralExpressionNode,CallExpressionNode,IfExpressionNode}.h
//   csql::SelectListNode, SequentialScanNode, GroupByNode   sql/qtree/*.h (only what the operators read)
//   csql::TableProvider                   sql/table_provider.h:42-48
//
// In a build of the reference these declarations are replaced by the reference's own headers; the
// operator implementations in gpu_operators.{h,cc} only use the members declared here.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace csql {

enum class SType : uint8_t { NIL = 0, UINT64 = 1, INT64 = 2, FLOAT64 = 3, BOOL = 4, STRING = 5, TIMESTAMP64 = 6 };
typedef uint8_t STag;
static const STag STAG_NULL = 1;

// sql_sizeof of a fixed-width element (svalue.cc:533-549): 8 B value + 1 B tag; BOOL 1 + 1; NIL 1
inline size_t sql_sizeof_fixed(SType t) {
  switch (t) {
    case SType::NIL: return 1;
    case SType::BOOL: return 2;
    default: return 9;
  }
}

class ReturnCode {
public:
  static ReturnCode success() { return ReturnCode(true, ""); }
  static ReturnCode error(const std::string& code, const std::string& msg) { return ReturnCode(false, code + ": " + msg); }
  bool isSuccess() const { return ok_; }
  const std::string& getMessage() const { return msg_; }
private:
  ReturnCode(bool ok, std::string msg) : ok_(ok), msg_(std::move(msg)) {}
  bool ok_;
  std::string msg_;
};

// growable byte buffer of packed elements (svalue.cc:410-517: malloc/realloc, append, clear keeps the capacity)
class SVector {
public:
  explicit SVector(SType type) : type_(type), data_(nullptr), capacity_(0), size_(0) {}
  SVector(const SVector&) = delete;
  SVector& operator=(const SVector&) = delete;
  SVector(SVector&& o) : type_(o.type_), data_(o.data_), capacity_(o.capacity_), size_(o.size_) { o.data_ = nullptr; o.capacity_ = o.size_ = 0; }
  ~SVector() { free(data_); }
  SType getType() const { return type_; }
  const void* getData() const { return data_; }
  void* getMutableData() { return data_; }
  size_t getSize() const { return size_; }
  void setSize(size_t n) { size_ = n; }
  void clear() { size_ = 0; }
  size_t getCapacity() const { return capacity_; }
  void increaseCapacity(size_t min_capacity) {
    if (min_capacity <= capacity_) return;
    void* p = realloc(data_, min_capacity);
    if (!p) throw std::bad_alloc();
    data_ = p;
    capacity_ = min_capacity;
  }
  void append(const void* data, size_t size) {
    if (size_ + size > capacity_) increaseCapacity(size_ + size);
    memcpy((char*) data_ + size_, data, size);
    size_ += size;
  }
private:
  SType type_;
  void* data_;
  size_t capacity_, size_;
};

class TableExpression {
public:
  virtual ~TableExpression() = default;
  virtual ReturnCode execute() = 0;
  // appends up to ~1024 rows to the caller-owned vectors; *len = rows appended, 0 = EOF (may be called again)
  virtual ReturnCode nextBatch(SVector* columns, size_t* len) = 0;
  virtual size_t getColumnCount() const = 0;
  virtual SType getColumnType(size_t idx) const = 0;
};

// ---- query tree (after planning: static types, implicit conversions already explicit to_<type> calls) ----------------

class ValueExpressionNode {
public:
  virtual ~ValueExpressionNode() = default;
  virtual SType getReturnType() const = 0;
  virtual std::vector<std::shared_ptr<ValueExpressionNode>> arguments() const { return {}; }
};
typedef std::shared_ptr<ValueExpressionNode> ExprRef;

class ColumnReferenceNode : public ValueExpressionNode {
public:
  ColumnReferenceNode(size_t column_index, SType type) : index_(column_index), type_(type) {}
  size_t columnIndex() const { return index_; }
  SType getReturnType() const override { return type_; }
private:
  size_t index_;
  SType type_;
};

class LiteralExpressionNode : public ValueExpressionNode {
public:
  static ExprRef u64(uint64_t v) { return ExprRef(new LiteralExpressionNode(SType::UINT64, v, "")); }
  static ExprRef i64(int64_t v) { return ExprRef(new LiteralExpressionNode(SType::INT64, (uint64_t) v, "")); }
  static ExprRef f64(double v) { uint64_t b; memcpy(&b, &v, 8); return ExprRef(new LiteralExpressionNode(SType::FLOAT64, b, "")); }
  static ExprRef boolean(bool v) { return ExprRef(new LiteralExpressionNode(SType::BOOL, v ? 1 : 0, "")); }
  static ExprRef string(const std::string& s) { return ExprRef(new LiteralExpressionNode(SType::STRING, 0, s)); }
  SType getReturnType() const override { return type_; }
  uint64_t bits() const { return bits_; }
  const std::string& str() const { return str_; }
private:
  LiteralExpressionNode(SType t, uint64_t bits, std::string s) : type_(t), bits_(bits), str_(std::move(s)) {}
  SType type_;
  uint64_t bits_;
  std::string str_;
};

class CallExpressionNode : public ValueExpressionNode {
public:
  // `symbol` is the resolved symbol string "name#ret/arg;arg;" (sql/runtime/symboltable.cc:33-39)
  CallExpressionNode(std::string symbol, SType return_type, std::vector<ExprRef> args)
      : symbol_(std::move(symbol)), type_(return_type), args_(std::move(args)) {}
  const std::string& getSymbol() const { return symbol_; }
  SType getReturnType() const override { return type_; }
  std::vector<ExprRef> arguments() const override { return args_; }
private:
  std::string symbol_;
  SType type_;
  std::vector<ExprRef> args_;
};

class IfExpressionNode : public ValueExpressionNode {
public:
  IfExpressionNode(ExprRef cond, ExprRef t, ExprRef f) : cond_(std::move(cond)), true_(std::move(t)), false_(std::move(f)) {}
  ExprRef conditional() const { return cond_; }
  ExprRef trueBranch() const { return true_; }
  ExprRef falseBranch() const { return false_; }
  SType getReturnType() const override { return true_->getReturnType(); }
  std::vector<ExprRef> arguments() const override { return {cond_, true_, false_}; }
private:
  ExprRef cond_, true_, false_;
};

class SelectListNode {
public:
  explicit SelectListNode(ExprRef e, std::string alias = "") : expr_(std::move(e)), alias_(std::move(alias)) {}
  ExprRef expression() const { return expr_; }
  const std::string& columnName() const { return alias_; }
private:
  ExprRef expr_;
  std::string alias_;
};
typedef std::shared_ptr<SelectListNode> SelectRef;

// sql/qtree/SequentialScanNode.h:100-191 - column references inside index selectedColumns()
class SequentialScanNode {
public:
  SequentialScanNode(std::string table_name, std::vector<std::pair<std::string, SType>> input_columns,
                     std::vector<SelectRef> select_list, ExprRef where_expr)
      : table_name_(std::move(table_name)), input_columns_(std::move(input_columns)), select_list_(std::move(select_list)),
        where_expr_(std::move(where_expr)) {}
  const std::string& tableName() const { return table_name_; }
  std::vector<SelectRef> selectList() const { return select_list_; }
  std::vector<std::string> selectedColumns() const {
    std::vector<std::string> v;
    for (const auto& c : input_columns_) v.push_back(c.first);
    return v;
  }
  SType getInputColumnType(size_t idx) const { return input_columns_.at(idx).second; }
  ExprRef whereExpression() const { return where_expr_; }   // null when absent (the reference uses Option<>)
private:
  std::string table_name_;
  std::vector<std::pair<std::string, SType>> input_columns_;
  std::vector<SelectRef> select_list_;
  ExprRef where_expr_;
};

// sql/qtree/GroupByNode.h:35-80 - column references inside index the input table's select list
class GroupByNode {
public:
  GroupByNode(std::vector<SelectRef> select_list, std::vector<ExprRef> group_exprs, std::shared_ptr<SequentialScanNode> input)
      : select_list_(std::move(select_list)), group_exprs_(std::move(group_exprs)), input_(std::move(input)) {}
  std::vector<SelectRef> selectList() const { return select_list_; }
  std::vector<ExprRef> groupExpressions() const { return group_exprs_; }
  std::shared_ptr<SequentialScanNode> inputTable() const { return input_; }
  bool isPartialAggregation() const { return partial_; }
  void setIsPartialAggreagtion(bool p) { partial_ = p; }   // (sic) the reference's spelling
private:
  std::vector<SelectRef> select_list_;
  std::vector<ExprRef> group_exprs_;
  std::shared_ptr<SequentialScanNode> input_;
  bool partial_ = false;
};

}  // namespace csql
