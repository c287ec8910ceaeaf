// gpu_binding.cc - see gpu_binding.h.  Compiled against the reference's own headers; host logic only (plan
// translation from the reference's qtree, file mapping, the TableExpression pull protocol); all compute is behind the
// C ABI of include/evqgpu.h.
#include "gpu_binding.h"
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <eventql/sql/qtree/ColumnReferenceNode.h>
#include <eventql/sql/qtree/LiteralExpressionNode.h>
#include <eventql/sql/qtree/CallExpressionNode.h>
#include <eventql/sql/qtree/IfExpressionNode.h>
#include <eventql/util/exception.h>

using namespace csql;

namespace evql_b200 {
namespace refbind {

static String lastError() { return evqgpu_last_error(); }

// ---- device ----------------------------------------------------------------------------------------------------------

GpuDevice::GpuDevice(int device) : ctx_(nullptr) {
  if (evqgpu_ctx_create(device, 0, &ctx_) != EVQGPU_OK) {
    RAISEF(kRuntimeError, "evqgpu_ctx_create: $0", lastError());
  }
}

GpuDevice::~GpuDevice() {
  for (auto& t : tables_) {
    evqgpu_table_destroy(t.second.table);
    munmap(t.second.addr, t.second.len);
  }
  evqgpu_ctx_destroy(ctx_);
}

evqgpu_table* GpuDevice::openTable(const String& filename) {
  struct stat st;
  const int fd = open(filename.c_str(), O_RDONLY);
  if (fd < 0) RAISEF(kRuntimeError, "cannot open $0", filename);
  if (fstat(fd, &st) != 0 || st.st_size == 0) {
    close(fd);
    RAISEF(kRuntimeError, "cannot stat $0", filename);
  }
  const uint64_t mtime_ns = (uint64_t) st.st_mtim.tv_sec * 1000000000ull + (uint64_t) st.st_mtim.tv_nsec;
  auto it = tables_.find(filename);
  if (it != tables_.end()) {
    const Mapped& m = it->second;
    if (m.ino == (uint64_t) st.st_ino && m.size == (uint64_t) st.st_size && m.mtime_ns == mtime_ns) {
      close(fd);
      return m.table;
    }
    // rewritten at the same path: drop the stale image
    evqgpu_table_destroy(m.table);
    munmap(m.addr, m.len);
    tables_.erase(it);
  }
  void* addr = mmap(nullptr, (size_t) st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (addr == MAP_FAILED) RAISEF(kRuntimeError, "cannot map $0", filename);
  evqgpu_table* t = nullptr;
  if (evqgpu_table_open(ctx_, addr, (uint64_t) st.st_size, &t) != EVQGPU_OK) {
    munmap(addr, (size_t) st.st_size);
    RAISEF(kRuntimeError, "evqgpu_table_open($0): $1", filename, lastError());
  }
  tables_[filename] = Mapped{addr, (size_t) st.st_size, t, (uint64_t) st.st_ino, (uint64_t) st.st_size, mtime_ns};
  return t;
}

SHA1Hash GpuDevice::fileIdentity(const String& filename) const {
  struct stat st;
  memset(&st, 0, sizeof(st));
  stat(filename.c_str(), &st);
  return SHA1::compute(StringUtil::format("$0|$1|$2|$3.$4", filename, (uint64_t) st.st_ino, (uint64_t) st.st_size,
                                          (uint64_t) st.st_mtim.tv_sec, (uint64_t) st.st_mtim.tv_nsec));
}

void GpuDevice::registerTable(const String& table_name, const Vector<String>& files) { registry_[table_name] = files; }

const Vector<String>* GpuDevice::filesOf(const String& table_name) const {
  auto it = registry_.find(table_name);
  return it == registry_.end() ? nullptr : &it->second;
}

// ---- qtree -> postfix programs ---------------------------------------------------------------------------------------

evqgpu_expr Program::view() const {
  evqgpu_expr e;
  e.code = code.data();
  e.len = (uint32_t) code.size();
  e.strings = strings.empty() ? nullptr : strings.data();
  e.strings_len = (uint32_t) strings.size();
  return e;
}

void translate(const RefPtr<ValueExpressionNode>& expr, const Vector<RefPtr<ValueExpressionNode>>& column_map, Program* out) {
  evqgpu_insn in;
  memset(&in, 0, sizeof(in));
  in.type = (uint8_t) expr->getReturnType();
  if (auto* c = dynamic_cast<const ColumnReferenceNode*>(expr.get())) {
    if (!c->hasColumnIndex()) RAISEF(kRuntimeError, "unresolved column reference: $0", c->columnName());
    if (!column_map.empty()) {
      // the GroupByNode / SubqueryNode column space: index into the input table's select list
      if (c->columnIndex() >= column_map.size()) RAISE(kRuntimeError, "column reference out of range");
      translate(column_map[c->columnIndex()], Vector<RefPtr<ValueExpressionNode>>(), out);
      return;
    }
    in.op = EVQ_X_INPUT;
    in.arg = (uint32_t) c->columnIndex();
  } else if (auto* l = dynamic_cast<const LiteralExpressionNode*>(expr.get())) {
    const SValue& v = l->value();
    in.op = EVQ_X_LITERAL;
    switch (v.getType()) {
      case SType::STRING: {
        const size_t len = sql_strlen(v.getData());
        in.imm = ((uint64_t) out->strings.size() << 32) | (uint64_t) len;
        out->strings.append(sql_cstr(v.getData()), len);
        break;
      }
      case SType::BOOL: {
        uint8_t b;
        memcpy(&b, v.getData(), 1);
        in.imm = b ? 1 : 0;
        break;
      }
      case SType::NIL:
        RAISE(kRuntimeError, "NULL literals are outside the device path");
      default:
        memcpy(&in.imm, v.getData(), 8);   // u64 / i64 / double bits / timestamp64
        break;
    }
  } else if (auto* call = dynamic_cast<const CallExpressionNode*>(expr.get())) {
    for (const auto& a : call->arguments()) translate(a, column_map, out);
    const int fid = evqgpu_function_lookup(call->getSymbol().c_str());
    if (fid < 0) RAISEF(kRuntimeError, "method not available on the device path: $0", call->getSymbol());
    in.op = EVQ_X_CALL;
    in.nargs = (uint16_t) call->arguments().size();
    in.arg = (uint32_t) fid;
  } else if (auto* iff = dynamic_cast<const IfExpressionNode*>(expr.get())) {
    translate(iff->conditional(), column_map, out);
    translate(iff->trueBranch(), column_map, out);
    translate(iff->falseBranch(), column_map, out);
    in.op = EVQ_X_IF;
    in.nargs = 3;
  } else {
    RAISEF(kRuntimeError, "expression is outside the device path: $0", expr->toSQL());
  }
  out->code.push_back(in);
}

// ---- pull protocol ---------------------------------------------------------------------------------------------------

GpuTableExpression::GpuTableExpression(Transaction* txn, ExecutionContext* ectx, RefPtr<GpuDevice> gpu, Vector<String> files)
    : txn_(txn), execution_context_(ectx), gpu_(gpu), filenames_(std::move(files)), query_(nullptr), cursor_(0), num_rows_(0),
      completed_(false) {
  if (execution_context_) execution_context_->incrementNumTasks();   // groupby.cc:54
}

GpuTableExpression::~GpuTableExpression() {
  if (query_) evqgpu_query_destroy(query_);
}

static ReturnCode heartbeat(Transaction* txn) {
  return txn ? txn->triggerHeartbeat() : ReturnCode::success();
}

ReturnCode GpuTableExpression::run(const evqgpu_query_desc& desc) {
  if (execution_context_) execution_context_->incrementNumTasksRunning();   // groupby.cc:70
  {
    auto rc = heartbeat(txn_);
    if (!rc.isSuccess()) RAISE(kRuntimeError, rc.getMessage());             // groupby.cc:100-105
  }
  std::vector<evqgpu_table*> tables;
  for (const auto& f : filenames_) tables.push_back(gpu_->openTable(f));
  if (query_) { evqgpu_query_destroy(query_); query_ = nullptr; }
  if (evqgpu_query_create(gpu_->handle(), &desc, &query_) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
  // enqueue, then keep the client connection alive while the device works
  if (evqgpu_query_enqueue(query_, tables.data(), (uint32_t) tables.size()) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
  {
    auto rc = heartbeat(txn_);
    if (!rc.isSuccess()) RAISE(kRuntimeError, rc.getMessage());
  }
  if (evqgpu_query_finish(query_) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
  if (evqgpu_query_num_rows(query_, &num_rows_) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
  cursor_ = 0;
  completed_ = false;
  if (evqgpu_query_num_columns(query_) != types_.size()) return ReturnCode::error("ERUNTIME", "device query and plan disagree on the column count");
  for (size_t i = 0; i < types_.size(); ++i)
    if ((SType) evqgpu_query_column_type(query_, (uint32_t) i) != types_[i])
      return ReturnCode::error("ERUNTIME", "device query and plan disagree on the type of column %d", (int) i);
  staging_.assign(types_.size(), std::vector<uint8_t>(kOutputBatchSize * 9));
  return ReturnCode::success();
}

ReturnCode GpuTableExpression::refresh() {
  if (!query_) return ReturnCode::error("ERUNTIME", "refresh before execute");
  if (evqgpu_query_num_rows(query_, &num_rows_) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
  cursor_ = 0;
  return ReturnCode::success();
}

ReturnCode GpuTableExpression::nextBatch(SVector* columns, size_t* len) {
  *len = 0;
  if (!query_) return ReturnCode::error("ERUNTIME", "nextBatch before execute");
  if (cursor_ >= num_rows_) {
    if (!completed_ && execution_context_) execution_context_->incrementNumTasksCompleted();   // groupby.cc:211
    completed_ = true;
    return ReturnCode::success();   // EOF: *len == 0; may be called again
  }
  {
    auto rc = heartbeat(txn_);
    if (!rc.isSuccess()) RAISE(kRuntimeError, rc.getMessage());
  }
  std::vector<void*> ptrs;
  for (auto& s : staging_) ptrs.push_back(s.data());
  uint64_t got = 0;
  if (evqgpu_query_fetch(query_, cursor_, kOutputBatchSize, ptrs.data(), &got) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
  for (size_t i = 0; i < staging_.size(); ++i) {
    const SType t = (SType) evqgpu_query_column_type(query_, (uint32_t) i);
    if (t == SType::STRING) {   // a string group key / projected string column: [u32 length][bytes][tag] elements
      uint64_t rows = 0, bytes = 0;
      if (evqgpu_query_fetch_strings(query_, (uint32_t) i, cursor_, got, nullptr, 0, &rows, &bytes) != EVQGPU_OK)
        return ReturnCode::error("ERUNTIME", lastError());
      std::vector<uint8_t> buf(bytes + 1);
      if (evqgpu_query_fetch_strings(query_, (uint32_t) i, cursor_, got, buf.data(), bytes, &rows, &bytes) != EVQGPU_OK)
        return ReturnCode::error("ERUNTIME", lastError());
      columns[i].append(buf.data(), bytes);
      continue;
    }
    columns[i].append(staging_[i].data(), got * sql_sizeof_static(t));   // already in the packed SVector encoding
  }
  cursor_ += got;
  *len = (size_t) got;
  if (cursor_ >= num_rows_ && !completed_) {
    completed_ = true;
    if (execution_context_) execution_context_->incrementNumTasksCompleted();
  }
  return ReturnCode::success();
}

// (ResultCursor and the operators above size their buffers from these BEFORE execute(), result_cursor.cc:34-43: the types
// come from the plan, not from the device query)
size_t GpuTableExpression::getColumnCount() const { return types_.size(); }
SType GpuTableExpression::getColumnType(size_t idx) const { return types_.at(idx); }

// the identity of the inputs (path, inode, size, mtime of every file) + the plan: what eventql::TableScan::getCacheKey
// supplies in the server (server/sql/table_scan.cc:173); lets PartialGroupByExpression use its query cache (groupby.cc:255-295)
Option<SHA1Hash> GpuTableExpression::getCacheKey() const {
  String k = plan_text_;
  for (const auto& f : filenames_) k += "|" + gpu_->fileIdentity(f).toString();
  return Some(SHA1::compute(k));
}

// ---- FastCSTableScan ---------------------------------------------------------------------------------------------------

class GpuCSTableScan::Impl : public GpuTableExpression {
public:
  Impl(Transaction* txn, ExecutionContext* ectx, RefPtr<GpuDevice> gpu, RefPtr<SequentialScanNode> stmt, const String& file)
      : GpuTableExpression(txn, ectx, gpu, Vector<String>{file}), stmt_(stmt), filter_enabled_(false) {
    plan_text_ = stmt_->toString();
    for (const auto& s : stmt_->selectList()) types_.push_back(s->expression()->getReturnType());
  }
  ReturnCode execute() override {
    evqgpu_table* t = gpu_->openTable(filenames_[0]);
    if (filter_enabled_) {
      std::vector<uint8_t> bits((filter_.size() + 7) / 8 + 1, 0);   // packed LSB first, as evqgpu_table_set_filter takes it
      for (size_t i = 0; i < filter_.size(); ++i)
        if (filter_[i]) bits[i >> 3] |= (uint8_t) (1u << (i & 7));
      if (evqgpu_table_set_filter(t, bits.data(), filter_.size(), 0) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
    }
    const Vector<String> cols = stmt_->selectedColumns();
    std::vector<const char*> names;
    for (const auto& c : cols) names.push_back(c.c_str());
    const Vector<RefPtr<ValueExpressionNode>> direct;
    Program where;
    if (!stmt_->whereExpression().isEmpty()) translate(stmt_->whereExpression().get(), direct, &where);
    std::vector<Program> sel(stmt_->selectList().size());
    {
      size_t i = 0;
      for (const auto& s : stmt_->selectList()) translate(s->expression(), direct, &sel[i++]);
    }
    std::vector<evqgpu_expr> selv;
    for (const auto& p : sel) selv.push_back(p.view());
    evqgpu_query_desc d;
    memset(&d, 0, sizeof(d));
    d.struct_size = sizeof(d);
    d.num_input_columns = (uint32_t) names.size();
    d.input_columns = names.data();
    d.where = where.view();
    d.num_select = (uint32_t) selv.size();
    d.select = selv.data();
    const ReturnCode rc = run(d);
    if (filter_enabled_) evqgpu_table_set_filter(t, nullptr, 0, 0);   // the filter belongs to this scan, the resident table is shared
    return rc;
  }
  RefPtr<SequentialScanNode> stmt_;
  std::vector<bool> filter_;
  bool filter_enabled_;
};

GpuCSTableScan::GpuCSTableScan(Transaction* txn, ExecutionContext* ectx, RefPtr<GpuDevice> gpu, RefPtr<SequentialScanNode> stmt,
                               const String& cstable_filename)
    : impl_(new Impl(txn, ectx, gpu, stmt->deepCopyAs<SequentialScanNode>(), cstable_filename)) {}   // CSTableScan.cc:695
GpuCSTableScan::~GpuCSTableScan() { delete impl_; }
ReturnCode GpuCSTableScan::execute() { return impl_->execute(); }
ReturnCode GpuCSTableScan::nextBatch(SVector* columns, size_t* len) { return impl_->nextBatch(columns, len); }
size_t GpuCSTableScan::getColumnCount() const { return impl_->getColumnCount(); }
SType GpuCSTableScan::getColumnType(size_t idx) const { return impl_->getColumnType(idx); }
Option<SHA1Hash> GpuCSTableScan::getCacheKey() const { return impl_->getCacheKey(); }
evqgpu_query* GpuCSTableScan::handle() const { return impl_->handle(); }
ReturnCode GpuCSTableScan::refresh() { return impl_->refresh(); }
void GpuCSTableScan::setFilter(std::vector<bool>&& filter) {
  impl_->filter_ = std::move(filter);
  impl_->filter_enabled_ = true;
}

// ---- GroupByExpression over a scan (optionally through a SubqueryNode) --------------------------------------------------

GpuGroupByExpression::GpuGroupByExpression(Transaction* txn, ExecutionContext* ectx, RefPtr<GpuDevice> gpu, RefPtr<GroupByNode> node,
                                           RefPtr<SequentialScanNode> scan, RefPtr<SubqueryNode> through, Vector<String> partition_files,
                                           uint32_t extra_flags)
    : GpuTableExpression(txn, ectx, gpu, std::move(partition_files)), node_(node), scan_(scan), through_(through),
      extra_flags_(extra_flags) {
  plan_text_ = node_->toString();
  for (const auto& s : node_->selectList()) types_.push_back(s->expression()->getReturnType());
}

// e with every column reference replaced through `column_map`, spelled as a postfix program over the scan's input columns
static Program lower(const RefPtr<ValueExpressionNode>& e, const Vector<RefPtr<ValueExpressionNode>>& column_map) {
  Program p;
  translate(e, column_map, &p);
  return p;
}

ReturnCode GpuGroupByExpression::execute() {
  const Vector<String> cols = scan_->selectedColumns();
  std::vector<const char*> names;
  for (const auto& c : cols) names.push_back(c.c_str());
  const Vector<RefPtr<ValueExpressionNode>> direct;

  // the column-index spaces: GROUP BY and its select list index the input table's select list.  With
  // a subquery in between (the H5 form `select ... from (select a, b from t where p) where q group by ...`) its select
  // list indexes the scan's, and its WHERE joins the scan's: SubqueryExpression is a pure projection + filter
  // (sql/statements/select/subquery.cc:57-120), so substituting its expressions is exact.
  Vector<RefPtr<ValueExpressionNode>> scan_out;
  for (const auto& s : scan_->selectList()) scan_out.push_back(s->expression());
  Program where;
  if (!scan_->whereExpression().isEmpty()) translate(scan_->whereExpression().get(), direct, &where);
  std::vector<Program> grp, sel;
  if (through_.get() == nullptr) {
    for (const auto& g : node_->groupExpressions()) grp.push_back(lower(g, scan_out));
    for (const auto& s : node_->selectList()) sel.push_back(lower(s->expression(), scan_out));
  } else {
    // two substitutions: GroupByNode -> subquery select list -> scan select list.  translate() substitutes one level and
    // spells the substituted expression with direct column indexes, so the subquery's select list is first rewritten as
    // programs over the scan's input columns, then spliced in by hand.
    struct Splice {
      static void emit(const RefPtr<ValueExpressionNode>& e, const std::vector<Program>& sub, Program* out) {
        evqgpu_insn in;
        memset(&in, 0, sizeof(in));
        in.type = (uint8_t) e->getReturnType();
        if (auto* c = dynamic_cast<const ColumnReferenceNode*>(e.get())) {
          if (!c->hasColumnIndex() || c->columnIndex() >= sub.size()) RAISE(kRuntimeError, "column reference out of range");
          const Program& s = sub[c->columnIndex()];
          for (evqgpu_insn i2 : s.code) {
            if (i2.op == EVQ_X_LITERAL && i2.type == (uint8_t) SType::STRING) {
              const uint64_t off = i2.imm >> 32, len = i2.imm & 0xffffffffull;
              i2.imm = ((uint64_t) out->strings.size() << 32) | len;
              out->strings.append(s.strings.data() + off, len);
            }
            out->code.push_back(i2);
          }
          return;
        }
        if (auto* call = dynamic_cast<const CallExpressionNode*>(e.get())) {
          for (const auto& a : call->arguments()) emit(a, sub, out);
          const int fid = evqgpu_function_lookup(call->getSymbol().c_str());
          if (fid < 0) RAISEF(kRuntimeError, "method not available on the device path: $0", call->getSymbol());
          in.op = EVQ_X_CALL;
          in.nargs = (uint16_t) call->arguments().size();
          in.arg = (uint32_t) fid;
          out->code.push_back(in);
          return;
        }
        if (auto* iff = dynamic_cast<const IfExpressionNode*>(e.get())) {
          emit(iff->conditional(), sub, out);
          emit(iff->trueBranch(), sub, out);
          emit(iff->falseBranch(), sub, out);
          in.op = EVQ_X_IF;
          in.nargs = 3;
          out->code.push_back(in);
          return;
        }
        // literals
        translate(e, Vector<RefPtr<ValueExpressionNode>>(), out);
      }
    };
    std::vector<Program> sub;
    for (const auto& s : through_->selectList()) sub.push_back(lower(s->expression(), scan_out));
    for (const auto& g : node_->groupExpressions()) { Program p; Splice::emit(g, sub, &p); grp.push_back(std::move(p)); }
    for (const auto& s : node_->selectList()) { Program p; Splice::emit(s->expression(), sub, &p); sel.push_back(std::move(p)); }
    if (!through_->whereExpression().isEmpty()) {
      // scan WHERE, then subquery WHERE on the rows it kept: if(scan_where, subquery_where, false) - the device `if` is
      // lazy like the reference's jumps (compiler.cc:174-209), so a raising subquery predicate (division by zero) is only
      // evaluated on rows the scan's WHERE let through, exactly like two stacked filters
      Program outer;
      Splice::emit(through_->whereExpression().get(), sub, &outer);
      if (where.code.empty()) {
        where = std::move(outer);
      } else {
        for (evqgpu_insn i2 : outer.code) {
          if (i2.op == EVQ_X_LITERAL && i2.type == (uint8_t) SType::STRING) {
            const uint64_t off = i2.imm >> 32, len = i2.imm & 0xffffffffull;
            i2.imm = ((uint64_t) where.strings.size() << 32) | len;
            where.strings.append(outer.strings.data() + off, len);
          }
          where.code.push_back(i2);
        }
        evqgpu_insn lit;
        memset(&lit, 0, sizeof(lit));
        lit.op = EVQ_X_LITERAL;
        lit.type = (uint8_t) SType::BOOL;
        lit.imm = 0;
        where.code.push_back(lit);
        evqgpu_insn iff;
        memset(&iff, 0, sizeof(iff));
        iff.op = EVQ_X_IF;
        iff.type = (uint8_t) SType::BOOL;
        iff.nargs = 3;
        where.code.push_back(iff);
      }
    }
  }
  std::vector<evqgpu_expr> grpv, selv;
  for (const auto& p : grp) grpv.push_back(p.view());
  for (const auto& p : sel) selv.push_back(p.view());
  evqgpu_query_desc d;
  memset(&d, 0, sizeof(d));
  d.struct_size = sizeof(d);
  d.flags = EVQGPU_QUERY_GROUPBY | extra_flags_;
  d.num_input_columns = (uint32_t) names.size();
  d.input_columns = names.data();
  d.where = where.view();
  d.num_group = (uint32_t) grpv.size();
  d.group = grpv.data();
  d.num_select = (uint32_t) selv.size();
  d.select = selv.data();
  return run(d);
}

// ---- PartialGroupByExpression rows ----------------------------------------------------------------------------------------

GpuPartialGroupByExpression::GpuPartialGroupByExpression(Transaction* txn, ExecutionContext* ectx, RefPtr<GpuDevice> gpu,
                                                         RefPtr<GroupByNode> node, RefPtr<SequentialScanNode> scan,
                                                         RefPtr<SubqueryNode> through, Vector<String> partition_files)
    : GpuGroupByExpression(txn, ectx, gpu, node, scan, through, std::move(partition_files), EVQGPU_QUERY_WIRE) {}

ReturnCode GpuPartialGroupByExpression::nextBatch(SVector* columns, size_t* len) {
  *len = 0;
  if (!query_) return ReturnCode::error("ERUNTIME", "nextBatch before execute");
  if (cursor_ >= num_rows_) return ReturnCode::success();
  std::vector<uint8_t> keys(kOutputBatchSize * 20), data(kOutputBatchSize * 64);
  std::vector<uint64_t> offs(kOutputBatchSize + 1);
  uint64_t got = 0, need = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (evqgpu_query_fetch_partial(query_, cursor_, kOutputBatchSize, keys.data(), data.data(), data.size(), offs.data(), &got, &need) != EVQGPU_OK)
      return ReturnCode::error("ERUNTIME", lastError());
    if (need <= data.size()) break;
    data.resize(need);
  }
  for (uint64_t r = 0; r < got; ++r) {   // groupby.cc:434-441: copyString(key), copyString(states)
    copyString((const char*) &keys[r * 20], 20, columns + 0);
    copyString((const char*) &data[offs[r]], (uint32_t) (offs[r + 1] - offs[r]), columns + 1);
  }
  cursor_ += got;
  *len = (size_t) got;
  return ReturnCode::success();
}

// ---- ORDER BY / LIMIT over a device-resident result ------------------------------------------------------------------------

static evqgpu_query* deviceQueryOf(TableExpression* e) {
  if (auto* g = dynamic_cast<GpuTableExpression*>(e)) return g->handle();
  if (auto* s = dynamic_cast<GpuCSTableScan*>(e)) return s->handle();
  if (auto* o = dynamic_cast<GpuOrderByExpression*>(e)) return deviceQueryOf(o->input());
  return nullptr;
}

static ReturnCode refreshOf(TableExpression* e) {
  if (auto* g = dynamic_cast<GpuTableExpression*>(e)) return g->refresh();
  if (auto* s = dynamic_cast<GpuCSTableScan*>(e)) return s->refresh();
  if (auto* o = dynamic_cast<GpuOrderByExpression*>(e)) return refreshOf(o->input());
  return ReturnCode::error("ERUNTIME", "not a device operator");
}

static bool isDeviceOperator(TableExpression* e) {
  return dynamic_cast<GpuTableExpression*>(e) || dynamic_cast<GpuCSTableScan*>(e) || dynamic_cast<GpuOrderByExpression*>(e);
}

GpuOrderByExpression::GpuOrderByExpression(Transaction* txn, ExecutionContext* ectx, std::vector<evqgpu_sort_spec> specs,
                                           ScopedPtr<TableExpression> input)
    : txn_(txn), execution_context_(ectx), specs_(std::move(specs)), input_(std::move(input)) {
  if (specs_.empty()) RAISE(kIllegalArgumentError, "can't execute ORDER BY: no sort specs");   // orderby.cc:52-54
  if (execution_context_) execution_context_->incrementNumTasks();
}

ReturnCode GpuOrderByExpression::execute() {
  auto rc = input_->execute();
  if (!rc.isSuccess()) return rc;
  if (execution_context_) execution_context_->incrementNumTasksRunning();
  if (evqgpu_query_order_by(deviceQueryOf(input_.get()), specs_.data(), (uint32_t) specs_.size()) != EVQGPU_OK)
    return ReturnCode::error("ERUNTIME", lastError());
  rc = refreshOf(input_.get());
  if (execution_context_) execution_context_->incrementNumTasksCompleted();
  return rc;
}

GpuLimitExpression::GpuLimitExpression(size_t limit, size_t offset, ScopedPtr<TableExpression> input)
    : limit_(limit), offset_(offset), input_(std::move(input)) {}

ReturnCode GpuLimitExpression::execute() {
  auto rc = input_->execute();
  if (!rc.isSuccess()) return rc;
  if (evqgpu_query_limit(deviceQueryOf(input_.get()), limit_, offset_) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
  return refreshOf(input_.get());
}

// ---- provider ------------------------------------------------------------------------------------------------------------

GpuCSTableScanProvider::GpuCSTableScanProvider(RefPtr<GpuDevice> gpu, const String& table_name, const Vector<String>& cstable_files)
    : gpu_(gpu), table_name_(table_name), files_(cstable_files), schema_(table_name, cstable_files.at(0)) {
  gpu_->registerTable(table_name_, files_);
}

Option<ScopedPtr<TableExpression>> GpuCSTableScanProvider::buildSequentialScan(Transaction* txn, ExecutionContext* execution_context,
                                                                              RefPtr<SequentialScanNode> seqscan) const {
  if (seqscan->tableName() != table_name_) return None<ScopedPtr<TableExpression>>();
  if (files_.size() != 1) RAISE(kRuntimeError, "a scan-only plan reads one cstable file (PartitionCursor concatenates segments)");
  return Option<ScopedPtr<TableExpression>>(
      ScopedPtr<TableExpression>(new GpuCSTableScan(txn, execution_context, gpu_, seqscan, files_[0])));
}

void GpuCSTableScanProvider::listTables(Function<void (const TableInfo& table)> fn) const { schema_.listTables(fn); }
Option<TableInfo> GpuCSTableScanProvider::describe(const String& table_name) const { return schema_.describe(table_name); }

// ---- scheduler -------------------------------------------------------------------------------------------------------------

ScopedPtr<TableExpression> GpuScheduler::buildGroupByExpression(Transaction* txn, ExecutionContext* execution_context,
                                                               RefPtr<GroupByNode> node) {
  // fuse scan + filter + aggregate when the input is a sequential scan of a GPU table, directly or through a subquery
  RefPtr<QueryTreeNode> input = node->inputTable();
  RefPtr<SequentialScanNode> scan;
  RefPtr<SubqueryNode> through;
  if (dynamic_cast<SequentialScanNode*>(input.get())) {
    scan = input.asInstanceOf<SequentialScanNode>();
  } else if (auto* sq = dynamic_cast<SubqueryNode*>(input.get())) {
    if (dynamic_cast<SequentialScanNode*>(sq->subquery().get())) {
      through = input.asInstanceOf<SubqueryNode>();
      scan = sq->subquery().asInstanceOf<SequentialScanNode>();
    }
  }
  const Vector<String>* files = scan.get() ? gpu_->filesOf(scan->tableName()) : nullptr;
  if (!files) {
    // not a GPU table (a CSV table, a join ...): the reference's own operator over whatever the input builds to
    return DefaultScheduler::buildGroupByExpression(txn, execution_context, node);
  }
  ++fused_groupbys_;
  if (node->isPartialAggregation()) {
    return mkScoped<TableExpression>(new GpuPartialGroupByExpression(txn, execution_context, gpu_, node, scan, through, *files));
  }
  return mkScoped<TableExpression>(new GpuGroupByExpression(txn, execution_context, gpu_, node, scan, through, *files));
}

ScopedPtr<TableExpression> GpuScheduler::buildOrderByExpression(Transaction* txn, ExecutionContext* execution_context,
                                                               RefPtr<OrderByNode> node) {
  // device sort when every sort expression is a column of the input's result and the input lives on the device;
  // otherwise the reference's OrderByExpression pulls from whatever the input builds to (our operators included)
  std::vector<evqgpu_sort_spec> specs;
  bool plain = !getenv("EVQGPU_HOST_ORDERBY");
  for (const auto& ss : node->sortSpecs()) {
    auto* c = dynamic_cast<const ColumnReferenceNode*>(ss.expr.get());
    if (!c || !c->hasColumnIndex() || c->getReturnType() == SType::STRING) { plain = false; break; }
    specs.push_back(evqgpu_sort_spec{(uint32_t) c->columnIndex(), ss.descending ? 1u : 0u});
  }
  if (!plain) return DefaultScheduler::buildOrderByExpression(txn, execution_context, node);
  auto input = buildTableExpression(txn, execution_context, node->inputTable().asInstanceOf<TableExpressionNode>());
  if (!isDeviceOperator(input.get()) || dynamic_cast<GpuPartialGroupByExpression*>(input.get())) {
    // (rebuilding is cheap: operators do no work before execute())
    return DefaultScheduler::buildOrderByExpression(txn, execution_context, node);
  }
  ++device_sorts_;
  return mkScoped<TableExpression>(new GpuOrderByExpression(txn, execution_context, std::move(specs), std::move(input)));
}

ScopedPtr<TableExpression> GpuScheduler::buildLimit(Transaction* txn, ExecutionContext* execution_context, RefPtr<LimitNode> node) {
  auto input = buildTableExpression(txn, execution_context, node->inputTable().asInstanceOf<TableExpressionNode>());
  if (!isDeviceOperator(input.get()) || dynamic_cast<GpuPartialGroupByExpression*>(input.get()) || getenv("EVQGPU_HOST_ORDERBY")) {
    return DefaultScheduler::buildLimit(txn, execution_context, node);
  }
  return mkScoped<TableExpression>(new GpuLimitExpression(node->limit(), node->offset(), std::move(input)));
}

}  // namespace refbind
}  // namespace evql_b200
