// gpu_binding.h - the GPU path bound to the reference's REAL classes.
//
// This header is compiled against the reference's own headers (-I /root/reference/src and the include paths of its build,
// tests/refbind/build_refsql.py); it is what a maintainer of the reference adds to the tree.  Nothing
// here is a restatement: every base class is the reference's.
//
//   evql_b200::refbind::GpuScheduler            : csql::DefaultScheduler      sql/scheduler.h:84-171 - overrides the virtual
//                                                 protected buildGroupByExpression / buildOrderByExpression / buildLimit
//                                                 (sql/scheduler.cc:153-182, :89-132, :57-68), the way eventql::Scheduler does
//                                                 for the cluster (server/sql/scheduler.h:50-68, scheduler.cc:55-77);
//                                                 installed with csql::Runtime::setScheduler (sql/runtime/runtime.h:71)
//   evql_b200::refbind::GpuCSTableScanProvider  : csql::TableProvider         sql/table_provider.h:42-48 - the GPU twin of
//                                                 csql::CSTableScanProvider (sql/CSTableScanProvider.cc:38-113); registered with
//                                                 TableRepository::addProvider (sql/runtime/tablerepository.cc:30-32)
//   evql_b200::refbind::GpuCSTableScan          : csql::AbstractCSTableScan   sql/CSTableScan.h:36-41,126-189 (FastCSTableScan)
//   evql_b200::refbind::GpuGroupByExpression    : csql::TableExpression       sql/statements/select/groupby.h:34-66 - a GROUP BY
//                                                 whose input is a sequential scan (or a SubqueryNode over one, cf. isPipelineable,
//                                                 server/sql/scheduler.cc:266-282) fused into one device pass
//   evql_b200::refbind::GpuPartialGroupByExpression                           groupby.h:66-100 (rows = 20-byte key, saved states)
//   evql_b200::refbind::GpuOrderByExpression / GpuLimitExpression             orderby.h:34-66, limit.h - over the device result
//
// Operator protocol kept: ReturnCode from execute / nextBatch, RAISE for what the reference raises, batches
// of <= 1024 rows appended to caller-owned SVectors, Option<SHA1Hash> getCacheKey(), txn->triggerHeartbeat() around the
// device pass and per fetched batch (CSTableScan.cc:193-198, groupby.cc:100-105), ExecutionContext::incrementNumTasks*
// (groupby.cc:54,70,211).  There is no CPU fallback: a plan over a GPU table that the device path cannot run raises.
#pragma once
#include <map>
#include <string>
#include <vector>
#include <eventql/sql/scheduler.h>
#include <eventql/sql/table_provider.h>
#include <eventql/sql/CSTableScan.h>
#include <eventql/sql/CSTableScanProvider.h>
#include <eventql/sql/transaction.h>
#include <eventql/sql/qtree/SequentialScanNode.h>
#include <eventql/sql/qtree/GroupByNode.h>
#include <eventql/sql/qtree/SubqueryNode.h>
#include <eventql/sql/qtree/OrderByNode.h>
#include <eventql/sql/qtree/LimitNode.h>
#include "../../../include/evqgpu.h"

namespace evql_b200 {
namespace refbind {

// One CUDA device: the evqgpu context, the resident tables (the analogue of the page cache: a cstable file is mapped and
// its referenced columns uploaded once) and the table name -> partition files registry of the providers.
class GpuDevice : public RefCounted {
public:
  explicit GpuDevice(int device);   // RAISEs without a device
  ~GpuDevice();
  evqgpu_ctx* handle() const { return ctx_; }
  // keyed by (path, inode, size, mtime): a file rewritten at the same path is opened anew
  evqgpu_table* openTable(const String& filename);
  SHA1Hash fileIdentity(const String& filename) const;
  void registerTable(const String& table_name, const Vector<String>& files);
  const Vector<String>* filesOf(const String& table_name) const;
private:
  struct Mapped { void* addr; size_t len; evqgpu_table* table; uint64_t ino, size, mtime_ns; };
  evqgpu_ctx* ctx_;
  std::map<String, Mapped> tables_;
  std::map<String, Vector<String>> registry_;
};

// an evqgpu postfix program (the shape of csql::vm::Program, sql/runtime/vm.h:44-75) spelled from a qtree expression
struct Program {
  std::vector<evqgpu_insn> code;
  std::string strings;
  evqgpu_expr view() const;
};
// column_map: what a ColumnReferenceNode of index i stands for (the input table's select list - the second of the two
// column-index spaces, sql/qtree/SequentialScanNode.cc:216-245); empty = the scan's own input columns
void translate(const RefPtr<csql::ValueExpressionNode>& expr, const Vector<RefPtr<csql::ValueExpressionNode>>& column_map,
               Program* out);

// pull protocol over a finished device query
class GpuTableExpression : public csql::TableExpression {
public:
  static const size_t kOutputBatchSize = 1024;   // sql/CSTableScan.h:142, groupby.h:36
  ~GpuTableExpression() override;
  ReturnCode nextBatch(csql::SVector* columns, size_t* len) override;
  size_t getColumnCount() const override;
  csql::SType getColumnType(size_t idx) const override;
  Option<SHA1Hash> getCacheKey() const override;
  evqgpu_query* handle() const { return query_; }
  ReturnCode refresh();   // after ORDER BY / LIMIT rewrote the device result
protected:
  GpuTableExpression(csql::Transaction* txn, csql::ExecutionContext* ectx, RefPtr<GpuDevice> gpu, Vector<String> files);
  ReturnCode run(const evqgpu_query_desc& desc);
  csql::Transaction* txn_;
  csql::ExecutionContext* execution_context_;
  RefPtr<GpuDevice> gpu_;
  Vector<String> filenames_;
  String plan_text_;        // qtree text: part of the cache key
  std::vector<csql::SType> types_;   // result column types, known from the plan before execute()
  evqgpu_query* query_;
  uint64_t cursor_, num_rows_;
  bool completed_;
  std::vector<std::vector<uint8_t>> staging_;
};

class GpuCSTableScan : public csql::AbstractCSTableScan {
public:
  GpuCSTableScan(csql::Transaction* txn, csql::ExecutionContext* execution_context, RefPtr<GpuDevice> gpu,
                 RefPtr<csql::SequentialScanNode> stmt, const String& cstable_filename);
  ~GpuCSTableScan() override;
  ReturnCode execute() override;
  ReturnCode nextBatch(csql::SVector* columns, size_t* len) override;
  size_t getColumnCount() const override;
  csql::SType getColumnType(size_t idx) const override;
  Option<SHA1Hash> getCacheKey() const override;
  void setFilter(std::vector<bool>&& filter) override;
  evqgpu_query* handle() const;
  ReturnCode refresh();
private:
  class Impl;
  Impl* impl_;
};

class GpuGroupByExpression : public GpuTableExpression {
public:
  // `scan`: the sequential scan under the GROUP BY; `through`: the SubqueryNode between them, or null
  GpuGroupByExpression(csql::Transaction* txn, csql::ExecutionContext* execution_context, RefPtr<GpuDevice> gpu,
                       RefPtr<csql::GroupByNode> node, RefPtr<csql::SequentialScanNode> scan, RefPtr<csql::SubqueryNode> through,
                       Vector<String> partition_files, uint32_t extra_flags = 0);
  ReturnCode execute() override;
protected:
  RefPtr<csql::GroupByNode> node_;
  RefPtr<csql::SequentialScanNode> scan_;
  RefPtr<csql::SubqueryNode> through_;
  uint32_t extra_flags_;
};

// the shard side of a cluster GROUP BY: rows are (STRING 20-byte SHA-1 group key, STRING saved states), groupby.cc:411-445
class GpuPartialGroupByExpression : public GpuGroupByExpression {
public:
  GpuPartialGroupByExpression(csql::Transaction* txn, csql::ExecutionContext* execution_context, RefPtr<GpuDevice> gpu,
                              RefPtr<csql::GroupByNode> node, RefPtr<csql::SequentialScanNode> scan,
                              RefPtr<csql::SubqueryNode> through, Vector<String> partition_files);
  ReturnCode nextBatch(csql::SVector* columns, size_t* len) override;
  size_t getColumnCount() const override { return 2; }
  csql::SType getColumnType(size_t) const override { return csql::SType::STRING; }
};

// ORDER BY / LIMIT on the device-resident result of a GPU operator (sort expressions that are result columns)
class GpuOrderByExpression : public csql::TableExpression {
public:
  GpuOrderByExpression(csql::Transaction* txn, csql::ExecutionContext* execution_context, std::vector<evqgpu_sort_spec> specs,
                       ScopedPtr<csql::TableExpression> input);
  ReturnCode execute() override;
  ReturnCode nextBatch(csql::SVector* columns, size_t* len) override { return input_->nextBatch(columns, len); }
  size_t getColumnCount() const override { return input_->getColumnCount(); }
  csql::SType getColumnType(size_t idx) const override { return input_->getColumnType(idx); }
  csql::TableExpression* input() const { return input_.get(); }
private:
  csql::Transaction* txn_;
  csql::ExecutionContext* execution_context_;
  std::vector<evqgpu_sort_spec> specs_;
  ScopedPtr<csql::TableExpression> input_;
};

class GpuLimitExpression : public csql::TableExpression {
public:
  GpuLimitExpression(size_t limit, size_t offset, ScopedPtr<csql::TableExpression> input);
  ReturnCode execute() override;
  ReturnCode nextBatch(csql::SVector* columns, size_t* len) override { return input_->nextBatch(columns, len); }
  size_t getColumnCount() const override { return input_->getColumnCount(); }
  csql::SType getColumnType(size_t idx) const override { return input_->getColumnType(idx); }
private:
  size_t limit_, offset_;
  ScopedPtr<csql::TableExpression> input_;
};

// csql::CSTableScanProvider's twin: a table name served from one cstable file or from several partition files
class GpuCSTableScanProvider : public csql::TableProvider {
public:
  GpuCSTableScanProvider(RefPtr<GpuDevice> gpu, const String& table_name, const Vector<String>& cstable_files);
  Option<ScopedPtr<csql::TableExpression>> buildSequentialScan(csql::Transaction* txn, csql::ExecutionContext* execution_context,
                                                               RefPtr<csql::SequentialScanNode> seqscan) const override;
  void listTables(Function<void (const csql::TableInfo& table)> fn) const override;
  Option<csql::TableInfo> describe(const String& table_name) const override;
private:
  RefPtr<GpuDevice> gpu_;
  String table_name_;
  Vector<String> files_;
  csql::CSTableScanProvider schema_;   // header parse + cstable -> SQL type mapping (CSTableScanProvider.cc:69-113)
};

class GpuScheduler : public csql::DefaultScheduler {
public:
  explicit GpuScheduler(RefPtr<GpuDevice> gpu) : gpu_(gpu) {}
  // counters for tests: how many operators of each kind the scheduler put on the device
  size_t fusedGroupBys() const { return fused_groupbys_; }
  size_t deviceSorts() const { return device_sorts_; }
protected:
  ScopedPtr<csql::TableExpression> buildGroupByExpression(csql::Transaction* txn, csql::ExecutionContext* execution_context,
                                                          RefPtr<csql::GroupByNode> node) override;
  ScopedPtr<csql::TableExpression> buildOrderByExpression(csql::Transaction* txn, csql::ExecutionContext* execution_context,
                                                          RefPtr<csql::OrderByNode> node) override;
  ScopedPtr<csql::TableExpression> buildLimit(csql::Transaction* txn, csql::ExecutionContext* execution_context,
                                              RefPtr<csql::LimitNode> node) override;
private:
  RefPtr<GpuDevice> gpu_;
  size_t fused_groupbys_ = 0, device_sorts_ = 0;
};

}  // namespace refbind
}  // namespace evql_b200
