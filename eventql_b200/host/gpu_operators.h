// gpu_operators.h - the drop-in operators: csql::TableExpression implementations that run on a B200 through the C ABI
// of include/evqgpu.h.  They replace, with the same constructor shape and pull protocol:
//
//   evql_b200::GpuCSTableScan          csql::FastCSTableScan        sql/CSTableScan.h:126-189, CSTableScan.cc:691-995
//   evql_b200::GpuGroupByExpression    csql::GroupByExpression over a FastCSTableScan input, fused into one device pass
//                                      sql/statements/select/groupby.h:34-66, groupby.cc:69-220
//   evql_b200::GpuPartitionCursor      eventql::PartitionCursor      server/sql/partition_cursor.cc:56-225 (visibility filters)
//   evql_b200::GpuTableProvider        csql::CSTableScanProvider     sql/CSTableScanProvider.cc:38-113
//   evql_b200::buildGroupByExpression  what a DefaultScheduler subclass returns from its virtual buildGroupByExpression
//                                      (sql/scheduler.cc:153-182), cf. eventql::Scheduler (server/sql/scheduler.cc:55-77)
//
// There is no CPU fallback: a plan the device path cannot run makes execute() return ReturnCode::error (the reference's
// convention, util/return_code.h) carrying evqgpu_last_error().
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "../../include/evqgpu.h"
#include "csql_mirror.h"

namespace evql_b200 {

// one CUDA device; owns the evqgpu context and the resident tables (the analogue of the page cache)
class GpuContext {
public:
  explicit GpuContext(int device);   // throws std::runtime_error without a device
  ~GpuContext();
  GpuContext(const GpuContext&) = delete;
  evqgpu_ctx* handle() const { return ctx_; }
  // open (once) a cstable file: mmap + evqgpu_table_open; columns are loaded lazily by the first query that reads them
  evqgpu_table* openTable(const std::string& filename);
private:
  struct Mapped { void* addr; size_t len; evqgpu_table* table; };
  evqgpu_ctx* ctx_;
  std::map<std::string, Mapped> tables_;
};

// qtree expression -> evqgpu postfix program (the shape of csql::vm::Program, sql/runtime/vm.h:44-75)
struct Program {
  std::vector<evqgpu_insn> code;
  std::string strings;
  evqgpu_expr view() const;
};
// column_map: rewrites ColumnReferenceNode indexes (GroupByNode space -> scan input columns); nullptr = identity
Program translate(const csql::ExprRef& expr, const std::vector<csql::ExprRef>* column_map = nullptr);

// common pull protocol over a finished evqgpu_query
class GpuQueryExpression : public csql::TableExpression {
public:
  ~GpuQueryExpression() override;
  csql::ReturnCode nextBatch(csql::SVector* columns, size_t* len) override;
  size_t getColumnCount() const override;
  csql::SType getColumnType(size_t idx) const override;
  static const size_t kOutputBatchSize = 1024;   // sql/CSTableScan.h:142, groupby.h:36
  // for the operators stacked on top (ORDER BY / LIMIT run on the device-resident result)
  evqgpu_query* handle() const { return query_; }
  csql::ReturnCode refresh();   // re-read the row count after the result was reordered / cut; rewinds the cursor
protected:
  GpuQueryExpression(GpuContext* gpu, std::vector<std::string> filenames) : gpu_(gpu), filenames_(std::move(filenames)) {}
  csql::ReturnCode run(const evqgpu_query_desc& desc);
  GpuContext* gpu_;
  std::vector<std::string> filenames_;
  evqgpu_query* query_ = nullptr;
  uint64_t cursor_ = 0, num_rows_ = 0;
  std::vector<std::vector<uint8_t>> staging_;
};

// FastCSTableScan: WHERE + projection, rows in table order
class GpuCSTableScan : public GpuQueryExpression {
public:
  GpuCSTableScan(GpuContext* gpu, std::shared_ptr<csql::SequentialScanNode> stmt, const std::string& cstable_filename);
  csql::ReturnCode execute() override;
  // FastCSTableScan::setFilter (sql/CSTableScan.h:36-41): the LSM visibility bitmap of the segment, ANDed with WHERE
  void setFilter(std::vector<bool>&& filter);
private:
  std::shared_ptr<csql::SequentialScanNode> stmt_;
  std::vector<bool> filter_;
  bool filter_enabled_ = false;
};

// eventql::PartitionCursor (server/sql/partition_cursor.h:38-70, partition_cursor.cc:56-225): the scan of ONE partition,
// i.e. of its segments - head arena, compacting arena, then the on-disk LSM tables newest first - each behind its
// visibility filter.  The reference opens the segments one after the other and builds every filter with a row loop over
// __lsm_id / __lsm_is_update / __lsm_skip and a std::set<SHA1Hash>; here the filters of all segments are built in one device
// pass (evqgpu_lsm_build_filters) and the segments are scanned as one table, in the cursor's order.
struct GpuPartitionSegment {
  std::string cstable_filename;
  bool is_arena = false;             // head / compacting arena: always filtered, skiplist from the arena
  std::vector<bool> arena_skiplist;  // PartitionArena::SkiplistReader, one flag per row (arena segments)
  bool has_skiplist = false;         // LSMTableRef::has_skiplist: the table carries a __lsm_skip column
  bool has_updates = false;          // LSMTableRef::has_updates
};
class GpuPartitionCursor : public GpuQueryExpression {
public:
  // segments in scan order (partition_cursor.cc:92-140): arenas first, then lsm_tables() from the back; the last
  // non-arena segment is the partition's oldest table (tblidx == 0)
  GpuPartitionCursor(GpuContext* gpu, std::shared_ptr<csql::SequentialScanNode> stmt, std::vector<GpuPartitionSegment> segments);
  csql::ReturnCode execute() override;
  // rows each segment's filter kept / whether it was filtered at all (needs_filter, partition_cursor.cc:149-155)
  const std::vector<uint64_t>& visibleRows() const { return visible_rows_; }
  const std::vector<bool>& filtered() const { return filtered_; }
private:
  std::shared_ptr<csql::SequentialScanNode> stmt_;
  std::vector<GpuPartitionSegment> segments_;
  std::vector<uint64_t> visible_rows_;
  std::vector<bool> filtered_;
};

// GroupByExpression whose input is a sequential scan of cstable partitions: scan + filter + aggregate in one device pass
class GpuGroupByExpression : public GpuQueryExpression {
public:
  GpuGroupByExpression(GpuContext* gpu, std::shared_ptr<csql::GroupByNode> node, std::vector<std::string> partition_files);
  csql::ReturnCode execute() override;
protected:
  std::shared_ptr<csql::GroupByNode> node_;
  uint32_t extra_flags_ = 0;
};

// PartialGroupByExpression (sql/statements/select/groupby.h:66-100, groupby.cc:223-472): the shard side of a cluster
// GROUP BY.  Same fused device pass, but the rows are (STRING 20-byte SHA-1 group key, STRING saved states) - what
// GroupByMergeExpression (groupby.cc:553-615) on a coordinator, CPU or GPU, loads and merges.
class GpuPartialGroupByExpression : public GpuGroupByExpression {
public:
  GpuPartialGroupByExpression(GpuContext* gpu, std::shared_ptr<csql::GroupByNode> node, std::vector<std::string> partition_files);
  csql::ReturnCode nextBatch(csql::SVector* columns, size_t* len) override;
  size_t getColumnCount() const override { return 2; }
  csql::SType getColumnType(size_t) const override { return csql::SType::STRING; }
  // The "store cache" step of PartialGroupByExpression::execute (groupby.cc:411-432): write the groups as the query cache
  // entry <cache_dir>/<SHA1(hex(input key) + hex(fingerprint))>.qc (groupby.cc:474-483, runtime/query_cache.cc:58-75).  Both
  // keys are 20-byte SHA-1 values (the input's getCacheKey() and the operator's expression fingerprint).  Call after execute().
  csql::ReturnCode storeCacheEntry(const std::string& cache_dir, const uint8_t input_cache_key[20], const uint8_t expression_fingerprint[20]);
};

// OrderByExpression (sql/statements/select/orderby.h:34-66, orderby.cc:58-160) over a device-resident result: the sort
// expressions are columns of the input (the planner appends hidden select items for anything else); execute() sorts on the
// device (evqgpu_query_order_by), nextBatch() streams the input's rows in the new order.
struct GpuSortSpec { size_t column; bool descending; };
class GpuOrderByExpression : public csql::TableExpression {
public:
  GpuOrderByExpression(std::vector<GpuSortSpec> sort_specs, std::unique_ptr<GpuQueryExpression> input);
  csql::ReturnCode execute() override;
  csql::ReturnCode nextBatch(csql::SVector* columns, size_t* len) override { return input_->nextBatch(columns, len); }
  size_t getColumnCount() const override { return input_->getColumnCount(); }
  csql::SType getColumnType(size_t idx) const override { return input_->getColumnType(idx); }
  GpuQueryExpression* input() const { return input_.get(); }
private:
  std::vector<GpuSortSpec> sort_specs_;
  std::unique_ptr<GpuQueryExpression> input_;
};

// LimitExpression (sql/statements/select/limit.h, limit.cc:43-112): rows [offset, offset + limit) of its input, which is
// either a device query or an ORDER BY over one
class GpuLimitExpression : public csql::TableExpression {
public:
  GpuLimitExpression(size_t limit, size_t offset, std::unique_ptr<csql::TableExpression> input, GpuQueryExpression* query);
  csql::ReturnCode execute() override;
  csql::ReturnCode nextBatch(csql::SVector* columns, size_t* len) override { return input_->nextBatch(columns, len); }
  size_t getColumnCount() const override { return input_->getColumnCount(); }
  csql::SType getColumnType(size_t idx) const override { return input_->getColumnType(idx); }
private:
  size_t limit_, offset_;
  std::unique_ptr<csql::TableExpression> input_;
  GpuQueryExpression* query_;   // the device query underneath input_ (owned by it)
};

// CSTableScanProvider: table name -> cstable file(s)
class GpuTableProvider {
public:
  GpuTableProvider(GpuContext* gpu, std::string table_name, std::vector<std::string> partition_files)
      : gpu_(gpu), table_name_(std::move(table_name)), files_(std::move(partition_files)) {}
  // TableProvider::buildSequentialScan: nullptr (the reference's None) when the table is not ours
  std::unique_ptr<csql::TableExpression> buildSequentialScan(std::shared_ptr<csql::SequentialScanNode> seqscan) const;
  // DefaultScheduler::buildGroupByExpression override: fuse when the input is a scan of our table, else nullptr
  std::unique_ptr<csql::TableExpression> buildGroupByExpression(std::shared_ptr<csql::GroupByNode> node) const;
private:
  GpuContext* gpu_;
  std::string table_name_;
  std::vector<std::string> files_;
};

}  // namespace evql_b200
