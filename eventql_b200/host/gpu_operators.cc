// gpu_operators.cc - see gpu_operators.h.  Host logic only: plan translation, file mapping, the pull protocol.
#include "gpu_operators.h"
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <stdexcept>

using namespace csql;

namespace evql_b200 {

static std::string lastError() { return evqgpu_last_error(); }

GpuContext::GpuContext(int device) : ctx_(nullptr) {
  if (evqgpu_ctx_create(device, 0, &ctx_) != EVQGPU_OK) throw std::runtime_error("evqgpu_ctx_create: " + lastError());
}

GpuContext::~GpuContext() {
  for (auto& t : tables_) {
    evqgpu_table_destroy(t.second.table);
    munmap(t.second.addr, t.second.len);
  }
  evqgpu_ctx_destroy(ctx_);
}

evqgpu_table* GpuContext::openTable(const std::string& filename) {
  auto it = tables_.find(filename);
  if (it != tables_.end()) return it->second.table;
  const int fd = open(filename.c_str(), O_RDONLY);
  if (fd < 0) throw std::runtime_error("cannot open " + filename);
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size == 0) { close(fd); throw std::runtime_error("cannot stat " + filename); }
  void* addr = mmap(nullptr, (size_t) st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (addr == MAP_FAILED) throw std::runtime_error("cannot map " + filename);
  evqgpu_table* t = nullptr;
  if (evqgpu_table_open(ctx_, addr, (uint64_t) st.st_size, &t) != EVQGPU_OK) {
    munmap(addr, (size_t) st.st_size);
    throw std::runtime_error("evqgpu_table_open(" + filename + "): " + lastError());
  }
  tables_[filename] = Mapped{addr, (size_t) st.st_size, t};
  return t;
}

evqgpu_expr Program::view() const {
  evqgpu_expr e;
  e.code = code.data();
  e.len = (uint32_t) code.size();
  e.strings = strings.empty() ? nullptr : strings.data();
  e.strings_len = (uint32_t) strings.size();
  return e;
}

static void emit(const ExprRef& e, const std::vector<ExprRef>* column_map, Program* out) {
  evqgpu_insn in;
  memset(&in, 0, sizeof(in));
  in.type = (uint8_t) e->getReturnType();
  if (auto* c = dynamic_cast<const ColumnReferenceNode*>(e.get())) {
    if (column_map) {
      // GroupByNode column space: index into the scan's select list (sql/qtree/SequentialScanNode.cc:216-245)
      if (c->columnIndex() >= column_map->size()) throw std::runtime_error("column reference out of range");
      emit((*column_map)[c->columnIndex()], nullptr, out);
      return;
    }
    in.op = EVQ_X_INPUT;
    in.arg = (uint32_t) c->columnIndex();
  } else if (auto* l = dynamic_cast<const LiteralExpressionNode*>(e.get())) {
    in.op = EVQ_X_LITERAL;
    if (l->getReturnType() == SType::STRING) {
      in.imm = ((uint64_t) out->strings.size() << 32) | (uint64_t) l->str().size();
      out->strings += l->str();
    } else {
      in.imm = l->bits();
    }
  } else if (auto* call = dynamic_cast<const CallExpressionNode*>(e.get())) {
    for (const auto& a : call->arguments()) emit(a, column_map, out);
    const int fid = evqgpu_function_lookup(call->getSymbol().c_str());
    if (fid < 0) throw std::runtime_error("method not available on the device path: " + call->getSymbol());
    in.op = EVQ_X_CALL;
    in.nargs = (uint16_t) call->arguments().size();
    in.arg = (uint32_t) fid;
  } else if (auto* iff = dynamic_cast<const IfExpressionNode*>(e.get())) {
    emit(iff->conditional(), column_map, out);
    emit(iff->trueBranch(), column_map, out);
    emit(iff->falseBranch(), column_map, out);
    in.op = EVQ_X_IF;
    in.nargs = 3;
  } else {
    throw std::runtime_error("unsupported expression node");
  }
  out->code.push_back(in);
}

Program translate(const ExprRef& expr, const std::vector<ExprRef>* column_map) {
  Program p;
  if (expr) emit(expr, column_map, &p);
  return p;
}

// ---- pull protocol ---------------------------------------------------------------------------------------------------

GpuQueryExpression::~GpuQueryExpression() {
  if (query_) evqgpu_query_destroy(query_);
}

ReturnCode GpuQueryExpression::run(const evqgpu_query_desc& desc) {
  try {
    std::vector<evqgpu_table*> tables;
    for (const auto& f : filenames_) tables.push_back(gpu_->openTable(f));
    if (query_) { evqgpu_query_destroy(query_); query_ = nullptr; }
    if (evqgpu_query_create(gpu_->handle(), &desc, &query_) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
    if (evqgpu_query_execute(query_, tables.data(), (uint32_t) tables.size()) != EVQGPU_OK)
      return ReturnCode::error("ERUNTIME", lastError());
    if (evqgpu_query_num_rows(query_, &num_rows_) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
    cursor_ = 0;
    staging_.assign(getColumnCount(), std::vector<uint8_t>(kOutputBatchSize * 9));
    return ReturnCode::success();
  } catch (const std::exception& e) {
    return ReturnCode::error("ERUNTIME", e.what());
  }
}

ReturnCode GpuQueryExpression::refresh() {
  if (!query_) return ReturnCode::error("ERUNTIME", "refresh before execute");
  if (evqgpu_query_num_rows(query_, &num_rows_) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
  cursor_ = 0;
  return ReturnCode::success();
}

// ---- PartitionCursor: the segments of one partition behind their visibility filters -------------------------------------

GpuPartitionCursor::GpuPartitionCursor(GpuContext* gpu, std::shared_ptr<SequentialScanNode> stmt, std::vector<GpuPartitionSegment> segments)
    : GpuQueryExpression(gpu, {}), stmt_(std::move(stmt)), segments_(std::move(segments)) {
  for (const auto& s : segments_) filenames_.push_back(s.cstable_filename);
}

ReturnCode GpuPartitionCursor::execute() {
  try {
    // (1) the filters (partition_cursor.cc:157-194), all segments in one device pass
    std::vector<evqgpu_lsm_segment> segs(segments_.size());
    std::vector<std::vector<uint8_t>> skipbits(segments_.size());
    int oldest = -1;
    for (size_t i = 0; i < segments_.size(); ++i)
      if (!segments_[i].is_arena) oldest = (int) i;
    for (size_t i = 0; i < segments_.size(); ++i) {
      const GpuPartitionSegment& s = segments_[i];
      memset(&segs[i], 0, sizeof(segs[i]));
      segs[i].table = gpu_->openTable(s.cstable_filename);
      if (s.is_arena) {
        // arenas are always filtered, rows are skipped by the arena's skiplist (partition_cursor.cc:92-127, :181-183)
        const uint64_t n = evqgpu_table_num_rows(segs[i].table);
        if (s.arena_skiplist.size() != n) return ReturnCode::error("ERUNTIME", "arena skiplist does not match the arena's row count");
        skipbits[i].assign((n + 7) / 8 + 1, 0);
        for (uint64_t r = 0; r < n; ++r)
          if (s.arena_skiplist[r]) skipbits[i][r >> 3] |= (uint8_t) (1u << (r & 7));
        segs[i].skiplist = skipbits[i].data();
      } else {
        segs[i].flags = EVQGPU_LSM_AUTO | (s.has_skiplist ? EVQGPU_LSM_SKIP_COLUMN : 0u) | (s.has_updates ? EVQGPU_LSM_HAS_UPDATES : 0u) |
                        ((int) i == oldest ? EVQGPU_LSM_OLDEST : 0u);
      }
    }
    if (evqgpu_lsm_build_filters(gpu_->handle(), segs.data(), (uint32_t) segs.size()) != EVQGPU_OK)
      return ReturnCode::error("ERUNTIME", lastError());
    visible_rows_.clear();
    filtered_.clear();
    for (const auto& s : segs) {
      visible_rows_.push_back(s.visible_rows);
      filtered_.push_back(s.filtered != 0);
    }
    // (2) one scan over the segments in the cursor's order (FastCSTableScan per segment in the reference, :196-203)
    const std::vector<std::string> cols = stmt_->selectedColumns();
    std::vector<const char*> names;
    for (const auto& c : cols) names.push_back(c.c_str());
    Program where = translate(stmt_->whereExpression());
    std::vector<Program> sel;
    for (const auto& s : stmt_->selectList()) sel.push_back(translate(s->expression()));
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
    // the filters belong to this cursor's snapshot, the resident tables are shared
    for (auto& s : segs) evqgpu_table_set_filter(s.table, nullptr, 0, 0);
    return rc;
  } catch (const std::exception& e) {
    return ReturnCode::error("ERUNTIME", e.what());
  }
}

// ---- ORDER BY / LIMIT over a device-resident result ----------------------------------------------------------------------

GpuOrderByExpression::GpuOrderByExpression(std::vector<GpuSortSpec> sort_specs, std::unique_ptr<GpuQueryExpression> input)
    : sort_specs_(std::move(sort_specs)), input_(std::move(input)) {
  if (sort_specs_.empty()) throw std::runtime_error("can't execute ORDER BY: no sort specs");   // orderby.cc:52-54
}

ReturnCode GpuOrderByExpression::execute() {
  ReturnCode rc = input_->execute();
  if (!rc.isSuccess()) return rc;
  std::vector<evqgpu_sort_spec> specs;
  for (const auto& s : sort_specs_) specs.push_back({(uint32_t) s.column, s.descending ? 1u : 0u});
  if (evqgpu_query_order_by(input_->handle(), specs.data(), (uint32_t) specs.size()) != EVQGPU_OK)
    return ReturnCode::error("ERUNTIME", lastError());
  return input_->refresh();
}

GpuLimitExpression::GpuLimitExpression(size_t limit, size_t offset, std::unique_ptr<TableExpression> input, GpuQueryExpression* query)
    : limit_(limit), offset_(offset), input_(std::move(input)), query_(query) {}

ReturnCode GpuLimitExpression::execute() {
  ReturnCode rc = input_->execute();
  if (!rc.isSuccess()) return rc;
  if (evqgpu_query_limit(query_->handle(), limit_, offset_) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
  return query_->refresh();
}

ReturnCode GpuQueryExpression::nextBatch(SVector* columns, size_t* len) {
  *len = 0;
  if (!query_) return ReturnCode::error("ERUNTIME", "nextBatch before execute");
  if (cursor_ >= num_rows_) return ReturnCode::success();   // EOF: *len == 0
  std::vector<void*> ptrs;
  for (auto& s : staging_) ptrs.push_back(s.data());
  uint64_t got = 0;
  if (evqgpu_query_fetch(query_, cursor_, kOutputBatchSize, ptrs.data(), &got) != EVQGPU_OK)
    return ReturnCode::error("ERUNTIME", lastError());
  for (size_t i = 0; i < staging_.size(); ++i) {
    if (getColumnType(i) == SType::STRING) {   // a string group key / projected string column: [u32 length][bytes][tag] elements
      uint64_t rows = 0, bytes = 0;
      if (evqgpu_query_fetch_strings(query_, (uint32_t) i, cursor_, got, nullptr, 0, &rows, &bytes) != EVQGPU_OK)
        return ReturnCode::error("ERUNTIME", lastError());
      std::vector<uint8_t> buf(bytes + 1);
      if (evqgpu_query_fetch_strings(query_, (uint32_t) i, cursor_, got, buf.data(), bytes, &rows, &bytes) != EVQGPU_OK)
        return ReturnCode::error("ERUNTIME", lastError());
      columns[i].append(buf.data(), bytes);
      continue;
    }
    columns[i].append(staging_[i].data(), got * sql_sizeof_fixed(getColumnType(i)));   // already in the packed SVector encoding
  }
  cursor_ += got;
  *len = (size_t) got;
  return ReturnCode::success();
}

size_t GpuQueryExpression::getColumnCount() const { return query_ ? evqgpu_query_num_columns(query_) : 0; }
SType GpuQueryExpression::getColumnType(size_t idx) const { return (SType) evqgpu_query_column_type(query_, (uint32_t) idx); }

// ---- FastCSTableScan ---------------------------------------------------------------------------------------------------

GpuCSTableScan::GpuCSTableScan(GpuContext* gpu, std::shared_ptr<SequentialScanNode> stmt, const std::string& cstable_filename)
    : GpuQueryExpression(gpu, {cstable_filename}), stmt_(std::move(stmt)) {}

void GpuCSTableScan::setFilter(std::vector<bool>&& filter) {
  filter_ = std::move(filter);
  filter_enabled_ = true;
}

ReturnCode GpuCSTableScan::execute() {
  try {
    if (filter_enabled_) {
      // packed LSB first, as evqgpu_table_set_filter takes it
      std::vector<uint8_t> bits((filter_.size() + 7) / 8 + 1, 0);
      for (size_t i = 0; i < filter_.size(); ++i)
        if (filter_[i]) bits[i >> 3] |= (uint8_t) (1u << (i & 7));
      evqgpu_table* t = gpu_->openTable(filenames_[0]);
      if (evqgpu_table_set_filter(t, bits.data(), filter_.size(), 0) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
    }
    const std::vector<std::string> cols = stmt_->selectedColumns();
    std::vector<const char*> names;
    for (const auto& c : cols) names.push_back(c.c_str());
    Program where = translate(stmt_->whereExpression());
    std::vector<Program> sel;
    for (const auto& s : stmt_->selectList()) sel.push_back(translate(s->expression()));
    std::vector<evqgpu_expr> selv;
    for (const auto& p : sel) selv.push_back(p.view());
    evqgpu_query_desc d;
    memset(&d, 0, sizeof(d));
    d.struct_size = sizeof(d);
    d.flags = 0;
    d.num_input_columns = (uint32_t) names.size();
    d.input_columns = names.data();
    d.where = where.view();
    d.num_select = (uint32_t) selv.size();
    d.select = selv.data();
    const ReturnCode rc = run(d);
    // the filter belongs to this scan, the resident table is shared
    if (filter_enabled_) evqgpu_table_set_filter(gpu_->openTable(filenames_[0]), nullptr, 0, 0);
    return rc;
  } catch (const std::exception& e) {
    return ReturnCode::error("ERUNTIME", e.what());
  }
}

// ---- GroupByExpression over a scan --------------------------------------------------------------------------------------

GpuGroupByExpression::GpuGroupByExpression(GpuContext* gpu, std::shared_ptr<GroupByNode> node, std::vector<std::string> partition_files)
    : GpuQueryExpression(gpu, std::move(partition_files)), node_(std::move(node)) {}

ReturnCode GpuGroupByExpression::execute() {
  try {
    auto scan = node_->inputTable();
    const std::vector<std::string> cols = scan->selectedColumns();
    std::vector<const char*> names;
    for (const auto& c : cols) names.push_back(c.c_str());
    // the two column-index spaces: group / select expressions reference the scan's select list
    std::vector<ExprRef> column_map;
    for (const auto& s : scan->selectList()) column_map.push_back(s->expression());
    Program where = translate(scan->whereExpression());
    std::vector<Program> grp, sel;
    for (const auto& g : node_->groupExpressions()) grp.push_back(translate(g, &column_map));
    for (const auto& s : node_->selectList()) sel.push_back(translate(s->expression(), &column_map));
    std::vector<evqgpu_expr> grpv, selv;
    for (const auto& p : grp) grpv.push_back(p.view());
    for (const auto& p : sel) selv.push_back(p.view());
    evqgpu_query_desc d;
    memset(&d, 0, sizeof(d));
    d.struct_size = sizeof(d);
    d.flags = EVQGPU_QUERY_GROUPBY | (node_->isPartialAggregation() ? EVQGPU_QUERY_PARTIAL : 0) | extra_flags_;
    d.num_input_columns = (uint32_t) names.size();
    d.input_columns = names.data();
    d.where = where.view();
    d.num_group = (uint32_t) grpv.size();
    d.group = grpv.data();
    d.num_select = (uint32_t) selv.size();
    d.select = selv.data();
    return run(d);
  } catch (const std::exception& e) {
    return ReturnCode::error("ERUNTIME", e.what());
  }
}

// ---- PartialGroupByExpression rows ----------------------------------------------------------------------------------------

GpuPartialGroupByExpression::GpuPartialGroupByExpression(GpuContext* gpu, std::shared_ptr<GroupByNode> node,
                                                         std::vector<std::string> partition_files)
    : GpuGroupByExpression(gpu, std::move(node), std::move(partition_files)) {
  extra_flags_ = EVQGPU_QUERY_WIRE;
}

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
  for (uint64_t r = 0; r < got; ++r) {   // copyString (svalue.cc:1097-1102): [u32 length][bytes][tag]
    const uint8_t tag = 0;
    uint32_t n = 20;
    columns[0].append(&n, 4);
    columns[0].append(&keys[r * 20], 20);
    columns[0].append(&tag, 1);
    n = (uint32_t) (offs[r + 1] - offs[r]);
    columns[1].append(&n, 4);
    columns[1].append(&data[offs[r]], n);
    columns[1].append(&tag, 1);
  }
  cursor_ += got;
  *len = (size_t) got;
  return ReturnCode::success();
}

ReturnCode GpuPartialGroupByExpression::storeCacheEntry(const std::string& cache_dir, const uint8_t input_cache_key[20],
                                                        const uint8_t expression_fingerprint[20]) {
  if (!query_) return ReturnCode::error("ERUNTIME", "storeCacheEntry before execute");
  char name[44];
  if (evqgpu_partial_cache_filename(input_cache_key, expression_fingerprint, name, sizeof(name)) != EVQGPU_OK)
    return ReturnCode::error("ERUNTIME", lastError());
  const std::string path = cache_dir + "/" + name;
  if (evqgpu_query_store_cache(query_, path.c_str()) != EVQGPU_OK) return ReturnCode::error("ERUNTIME", lastError());
  return ReturnCode::success();
}

// ---- provider / scheduler hooks ----------------------------------------------------------------------------------------

std::unique_ptr<TableExpression> GpuTableProvider::buildSequentialScan(std::shared_ptr<SequentialScanNode> seqscan) const {
  if (seqscan->tableName() != table_name_) return nullptr;
  if (files_.size() != 1) throw std::runtime_error("a scan-only plan reads one cstable file (PartitionCursor concatenates)");
  return std::unique_ptr<TableExpression>(new GpuCSTableScan(gpu_, std::move(seqscan), files_[0]));
}

std::unique_ptr<TableExpression> GpuTableProvider::buildGroupByExpression(std::shared_ptr<GroupByNode> node) const {
  auto scan = node->inputTable();
  if (!scan || scan->tableName() != table_name_) return nullptr;
  return std::unique_ptr<TableExpression>(new GpuGroupByExpression(gpu_, std::move(node), files_));
}

}  // namespace evql_b200
