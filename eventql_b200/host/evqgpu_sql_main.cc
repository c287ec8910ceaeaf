// evqgpu_sql - drives the drop-in operators the way the reference's sql_tests runner drives its own
// (test/sql_tests.cc:232-274: provider -> plan -> execute -> pull batches -> print), for the named query shapes.
// The query trees are built by hand in the exact shape QueryPlanBuilder::buildGroupBy produces (runtime/queryplanbuilder.cc:
// 439-540): the scan's select list holds bare column references appended by getComputedColumnIndex, the GroupByNode's
// expressions index that list.
//
//   evqgpu_sql q1    <lineitem partition .cst>...     Q1-style GROUP BY flag, status (C3)
//   evqgpu_sql q6    <lineitem .cst>...               Q6-style global aggregate (C2)
//   evqgpu_sql count <file.cst> <column>              select count(1), sum(c), min(c), max(c), mean(c) ... where c > 0
//   evqgpu_sql scan  <file.cst> <column>              select <column> from t        (FastCSTableScan alone)
//   evqgpu_sql scanf <file.cst> <column> <m>          the same with setFilter(row % m == 0)   (LSM visibility filter)
//   evqgpu_sql partial <file.cst> <column>            select c, count(1), sum(c) from t where c >= 0 group by c   as
//                                                     PartialGroupByExpression rows: hex key ; hex saved states
//   evqgpu_sql top   <file.cst> <column> <limit> <offset>
//                                                     select c, count(1), sum(c) from t where c >= 0 group by c
//                                                     order by c desc limit <limit> offset <offset>
//                                                     (LimitExpression over OrderByExpression over the fused GROUP BY)
//
// Output: one line per row, ';' separated (uint64 decimal, float64 %.17g, bool true|false, NULL).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gpu_operators.h"

using namespace csql;
using namespace evql_b200;

static const char* tname(SType t) {
  static const char* n[] = {"nil", "uint64", "int64", "float64", "bool", "string", "timestamp64"};
  return n[(int) t];
}

static ExprRef call(const std::string& name, SType ret, std::vector<ExprRef> args) {
  std::string sym = name + "#" + tname(ret) + "/";
  for (const auto& a : args) sym += std::string(tname(name == "count" ? SType::NIL : a->getReturnType())) + ";";
  return ExprRef(new CallExpressionNode(sym, ret, std::move(args)));
}
static ExprRef col(size_t i, SType t = SType::UINT64) { return ExprRef(new ColumnReferenceNode(i, t)); }
static ExprRef u(uint64_t v) { return LiteralExpressionNode::u64(v); }
static ExprRef cmp(const char* op, ExprRef a, ExprRef b) { return call(op, SType::BOOL, {a, b}); }
static ExprRef land(ExprRef a, ExprRef b) { return call("logical_and", SType::BOOL, {a, b}); }
static ExprRef arith(const char* op, ExprRef a, ExprRef b) { return call(op, SType::UINT64, {a, b}); }
static ExprRef count1() { return call("count", SType::UINT64, {call("to_nil", SType::NIL, {u(1)})}); }
static ExprRef agg(const char* fn, ExprRef a, SType ret = SType::UINT64) { return call(fn, ret, {a}); }
static SelectRef sel(ExprRef e) { return SelectRef(new SelectListNode(std::move(e))); }

static int pull(TableExpression* te) {
  ReturnCode rc = te->execute();
  if (!rc.isSuccess()) { printf("ERROR!\n%s\n", rc.getMessage().c_str()); return 1; }
  const size_t nc = te->getColumnCount();
  std::vector<SVector> cols;
  for (size_t i = 0; i < nc; ++i) cols.emplace_back(te->getColumnType(i));
  for (;;) {
    for (auto& c : cols) c.clear();
    size_t n = 0;
    rc = te->nextBatch(cols.data(), &n);
    if (!rc.isSuccess()) { printf("ERROR!\n%s\n", rc.getMessage().c_str()); return 1; }
    if (n == 0) break;
    std::vector<const uint8_t*> scur(nc);   // cursors into the variable-length STRING columns
    for (size_t i = 0; i < nc; ++i) scur[i] = (const uint8_t*) cols[i].getData();
    for (size_t r = 0; r < n; ++r) {
      std::string line;
      for (size_t i = 0; i < nc; ++i) {
        const SType t = te->getColumnType(i);
        char buf[64];
        if (i) line += ";";
        if (t == SType::STRING) {   // [u32 length][bytes][tag] (svalue.cc:533-549), printed like evqlref -H
          uint32_t len; memcpy(&len, scur[i], 4);
          if (scur[i][4 + len] & STAG_NULL) line += "NULL";
          else { line += "x"; for (uint32_t b = 0; b < len; ++b) { snprintf(buf, sizeof(buf), "%02x", scur[i][4 + b]); line += buf; } }
          scur[i] += 5 + len;
          continue;
        }
        const uint8_t* p = (const uint8_t*) cols[i].getData() + r * sql_sizeof_fixed(t);
        if (t == SType::BOOL) { line += (p[1] & STAG_NULL) ? "NULL" : (p[0] ? "true" : "false"); continue; }
        if (p[8] & STAG_NULL) { line += "NULL"; continue; }
        uint64_t v; memcpy(&v, p, 8);
        if (t == SType::FLOAT64) { double d; memcpy(&d, p, 8); snprintf(buf, sizeof(buf), "%.17g", d); }
        else if (t == SType::INT64) snprintf(buf, sizeof(buf), "%lld", (long long) v);
        else snprintf(buf, sizeof(buf), "%llu", (unsigned long long) v);
        line += buf;
      }
      puts(line.c_str());
    }
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: evqgpu_sql q1|q6|count|scan <file.cst>... [column]\n"); return 2; }
  const std::string mode = argv[1];
  try {
    if (mode == "translate") {
      // evqgpu_sql translate q6|strings|ifexpr: print the postfix evqgpu_insn program translate() emits for a WHERE expression
      // (no device needed) - "op type nargs arg imm" per instruction, then the string pool as hex
      const SType U = SType::UINT64, S = SType::STRING;
      ExprRef e;
      const std::string which = argv[2];
      if (which == "q6") {
        e = land(land(land(land(land(cmp("gte", col(0), u(8766)), cmp("lt", col(0), u(9131))), cmp("gte", col(1), u(5))),
                           cmp("lte", col(1), u(7))), cmp("lt", col(2), u(24))), cmp("gt", col(3), u(0)));
      } else if (which == "strings") {
        e = land(cmp("eq", col(1, S), LiteralExpressionNode::string("google.de")), cmp("neq", LiteralExpressionNode::string(""), col(2, S)));
      } else if (which == "ifexpr") {
        e = cmp("gt", ExprRef(new IfExpressionNode(cmp("lt", col(0), u(50)), arith("mul", col(1), u(3)), col(2))), u(7));
      } else {
        fprintf(stderr, "translate: unknown expression\n");
        return 2;
      }
      (void) U;
      const Program p = translate(e);
      for (const auto& in : p.code)
        printf("%u %u %u %u %llu\n", (unsigned) in.op, (unsigned) in.type, (unsigned) in.nargs, (unsigned) in.arg, (unsigned long long) in.imm);
      std::string hex;
      char buf[4];
      for (unsigned char ch : p.strings) { snprintf(buf, sizeof(buf), "%02x", ch); hex += buf; }
      printf("strings %s\n", hex.c_str());
      return 0;
    }
    GpuContext gpu(0);
    const SType U = SType::UINT64;
    if (mode == "q1" || mode == "q6") {
      std::vector<std::string> files(argv + 2, argv + argc);
      GpuTableProvider provider(&gpu, "lineitem", files);
      std::shared_ptr<GroupByNode> node;
      if (mode == "q1") {
        // input columns in first-reference order: WHERE columns first, then those pulled in by GROUP BY / select
        std::vector<std::pair<std::string, SType>> in = {{"shipdate", U}, {"quantity", U}, {"price", U}, {"discount", U}, {"tax", U},
                                                         {"flag", U}, {"status", U}};
        ExprRef where = land(land(land(land(cmp("lte", col(0), u(10471)), cmp("gt", col(1), u(0))), cmp("gt", col(2), u(0))),
                                  cmp("gte", col(3), u(0))), cmp("gte", col(4), u(0)));
        // scan select list = bare columns in the order the GROUP BY / select resolution asked for them
        std::vector<SelectRef> scan_sel = {sel(col(5)), sel(col(6)), sel(col(1)), sel(col(2)), sel(col(3)), sel(col(4))};
        auto scan = std::make_shared<SequentialScanNode>("lineitem", in, scan_sel, where);
        // GroupByNode space: 0 flag, 1 status, 2 quantity, 3 price, 4 discount, 5 tax
        ExprRef disc = arith("mul", col(3), arith("sub", u(100), col(4)));
        std::vector<SelectRef> gsel = {sel(col(0)), sel(col(1)), sel(count1()), sel(agg("sum", col(2))), sel(agg("sum", col(3))),
                                       sel(agg("sum", disc)), sel(agg("sum", arith("mul", disc, arith("add", u(100), col(5))))),
                                       sel(agg("sum", col(4))), sel(agg("mean", col(2), SType::FLOAT64)),
                                       sel(agg("mean", col(3), SType::FLOAT64)), sel(agg("mean", col(4), SType::FLOAT64))};
        node = std::make_shared<GroupByNode>(gsel, std::vector<ExprRef>{col(0), col(1)}, scan);
      } else {
        std::vector<std::pair<std::string, SType>> in = {{"shipdate", U}, {"discount", U}, {"quantity", U}, {"price", U}};
        ExprRef where = land(land(land(land(land(cmp("gte", col(0), u(8766)), cmp("lt", col(0), u(9131))), cmp("gte", col(1), u(5))),
                                       cmp("lte", col(1), u(7))), cmp("lt", col(2), u(24))), cmp("gt", col(3), u(0)));
        std::vector<SelectRef> scan_sel = {sel(col(3)), sel(col(1))};
        auto scan = std::make_shared<SequentialScanNode>("lineitem", in, scan_sel, where);
        std::vector<SelectRef> gsel = {sel(count1()), sel(agg("sum", arith("mul", col(0), col(1))))};
        node = std::make_shared<GroupByNode>(gsel, std::vector<ExprRef>{}, scan);
      }
      auto te = provider.buildGroupByExpression(node);
      if (!te) { fprintf(stderr, "provider declined the plan\n"); return 1; }
      return pull(te.get());
    }
    if ((mode == "count" || mode == "scan") && argc >= 4) {
      GpuTableProvider provider(&gpu, "t", {argv[2]});
      std::vector<std::pair<std::string, SType>> in = {{argv[3], U}};
      if (mode == "scan") {
        auto scan = std::make_shared<SequentialScanNode>("t", in, std::vector<SelectRef>{sel(col(0))}, nullptr);
        auto te = provider.buildSequentialScan(scan);
        return pull(te.get());
      }
      auto scan = std::make_shared<SequentialScanNode>("t", in, std::vector<SelectRef>{sel(col(0))}, cmp("gt", col(0), u(0)));
      std::vector<SelectRef> gsel = {sel(count1()), sel(agg("sum", col(0))), sel(agg("min", col(0))), sel(agg("max", col(0))),
                                     sel(agg("mean", col(0), SType::FLOAT64))};
      auto node = std::make_shared<GroupByNode>(gsel, std::vector<ExprRef>{}, scan);
      auto te = provider.buildGroupByExpression(node);
      return pull(te.get());
    }
    if (mode == "scanf" && argc >= 5) {
      GpuTableProvider provider(&gpu, "t", {argv[2]});
      std::vector<std::pair<std::string, SType>> in = {{argv[3], U}};
      auto scan = std::make_shared<SequentialScanNode>("t", in, std::vector<SelectRef>{sel(col(0))}, nullptr);
      auto te = provider.buildSequentialScan(scan);
      auto* gs = dynamic_cast<GpuCSTableScan*>(te.get());
      if (!gs) { fprintf(stderr, "provider declined the plan\n"); return 1; }
      const uint64_t m = strtoull(argv[4], nullptr, 10);
      const uint64_t nrows = evqgpu_table_num_rows(gpu.openTable(argv[2]));
      std::vector<bool> keep(nrows);
      for (uint64_t i = 0; i < nrows; ++i) keep[i] = m && i % m == 0;
      gs->setFilter(std::move(keep));
      return pull(te.get());
    }
    if (mode == "partial" && argc >= 4) {
      std::vector<std::pair<std::string, SType>> in = {{argv[3], U}};
      auto scan = std::make_shared<SequentialScanNode>("t", in, std::vector<SelectRef>{sel(col(0))}, cmp("gte", col(0), u(0)));
      std::vector<SelectRef> gsel = {sel(col(0)), sel(count1()), sel(agg("sum", col(0)))};
      auto node = std::make_shared<GroupByNode>(gsel, std::vector<ExprRef>{col(0)}, scan);
      GpuPartialGroupByExpression te(&gpu, node, {argv[2]});
      ReturnCode rc = te.execute();
      if (!rc.isSuccess()) { printf("ERROR!\n%s\n", rc.getMessage().c_str()); return 1; }
      if (argc >= 5) {   // evqgpu_sql partial <file> <column> <cache dir>: also store the query cache entry (keys of all 0x11 / 0x22 bytes)
        uint8_t in_key[20], fp[20];
        memset(in_key, 0x11, 20);
        memset(fp, 0x22, 20);
        rc = te.storeCacheEntry(argv[4], in_key, fp);
        if (!rc.isSuccess()) { printf("ERROR!\n%s\n", rc.getMessage().c_str()); return 1; }
      }
      std::vector<SVector> cols;
      cols.emplace_back(SType::STRING);
      cols.emplace_back(SType::STRING);
      for (;;) {
        for (auto& c : cols) c.clear();
        size_t n = 0;
        rc = te.nextBatch(cols.data(), &n);
        if (!rc.isSuccess()) { printf("ERROR!\n%s\n", rc.getMessage().c_str()); return 1; }
        if (n == 0) break;
        const uint8_t* p[2] = {(const uint8_t*) cols[0].getData(), (const uint8_t*) cols[1].getData()};
        for (size_t r = 0; r < n; ++r) {
          std::string line;
          for (int c = 0; c < 2; ++c) {
            uint32_t len; memcpy(&len, p[c], 4);
            char buf[4];
            for (uint32_t i = 0; i < len; ++i) { snprintf(buf, sizeof(buf), "%02x", p[c][4 + i]); line += buf; }
            p[c] += 4 + len + 1;
            if (c == 0) line += ";";
          }
          puts(line.c_str());
        }
      }
      return 0;
    }
    if (mode == "strgroup" && argc >= 5) {
      // evqgpu_sql strgroup <file.cst> <string column> <numeric column>:
      //   select <s>, count(1), sum(<n>) from t where <n> >= 0 and <s> != 'x' group by <s>
      GpuTableProvider provider(&gpu, "t", {argv[2]});
      const SType S = SType::STRING;
      std::vector<std::pair<std::string, SType>> in = {{argv[4], U}, {argv[3], S}};
      ExprRef lit = LiteralExpressionNode::string("x");
      ExprRef scol = col(1, S);
      ExprRef where = land(cmp("gte", col(0), u(0)), call("neq", SType::BOOL, {scol, lit}));
      auto scan = std::make_shared<SequentialScanNode>("t", in, std::vector<SelectRef>{sel(scol), sel(col(0))}, where);
      // GroupByNode space: 0 = the string column, 1 = the numeric column
      ExprRef gkey = col(0, S);
      std::vector<SelectRef> gsel = {sel(gkey), sel(count1()), sel(agg("sum", col(1)))};
      auto node = std::make_shared<GroupByNode>(gsel, std::vector<ExprRef>{gkey}, scan);
      auto te = provider.buildGroupByExpression(node);
      if (!te) { fprintf(stderr, "provider declined the plan\n"); return 1; }
      return pull(te.get());
    }
    if (mode == "partition" && argc >= 4) {
      // evqgpu_sql partition <column> <seg.cst>[:flags]...   flags: a = arena (skiplist: every 7th row), s = has_skiplist,
      // u = has_updates; segments in the cursor's order.  `select <column> from t where <column> >= 0` through
      // GpuPartitionCursor (PartitionCursor::nextBatch, server/sql/partition_cursor.cc:56-80)
      std::vector<std::pair<std::string, SType>> in = {{argv[2], U}};
      auto scan = std::make_shared<SequentialScanNode>("t", in, std::vector<SelectRef>{sel(col(0))}, cmp("gte", col(0), u(0)));
      std::vector<GpuPartitionSegment> segs;
      for (int i = 3; i < argc; ++i) {
        std::string a = argv[i];
        GpuPartitionSegment sg;
        const size_t c = a.rfind(':');
        std::string fl;
        if (c != std::string::npos && a.find('/', c) == std::string::npos) { fl = a.substr(c + 1); a = a.substr(0, c); }
        sg.cstable_filename = a;
        sg.is_arena = fl.find('a') != std::string::npos;
        sg.has_skiplist = fl.find('s') != std::string::npos;
        sg.has_updates = fl.find('u') != std::string::npos;
        if (sg.is_arena) {
          const uint64_t n = evqgpu_table_num_rows(gpu.openTable(a));
          sg.arena_skiplist.resize(n);
          for (uint64_t r = 0; r < n; ++r) sg.arena_skiplist[r] = r % 7 == 3;
        }
        segs.push_back(std::move(sg));
      }
      GpuPartitionCursor cursor(&gpu, scan, segs);
      const int rc = pull(&cursor);
      if (rc == 0) {
        std::string line = "#visible";
        for (size_t i = 0; i < segs.size(); ++i)
          line += " " + std::to_string(cursor.visibleRows()[i]) + (cursor.filtered()[i] ? "f" : "u");
        fprintf(stderr, "%s\n", line.c_str());
      }
      return rc;
    }
    if (mode == "top" && argc >= 6) {
      GpuTableProvider provider(&gpu, "t", {argv[2]});
      std::vector<std::pair<std::string, SType>> in = {{argv[3], U}};
      auto scan = std::make_shared<SequentialScanNode>("t", in, std::vector<SelectRef>{sel(col(0))}, cmp("gte", col(0), u(0)));
      std::vector<SelectRef> gsel = {sel(col(0)), sel(count1()), sel(agg("sum", col(0)))};
      auto node = std::make_shared<GroupByNode>(gsel, std::vector<ExprRef>{col(0)}, scan);
      auto te = provider.buildGroupByExpression(node);
      auto* gq = dynamic_cast<GpuQueryExpression*>(te.get());
      if (!gq) { fprintf(stderr, "provider declined the plan\n"); return 1; }
      te.release();
      std::unique_ptr<GpuQueryExpression> input(gq);
      std::unique_ptr<TableExpression> ordered(new GpuOrderByExpression({{0, true}}, std::move(input)));
      GpuLimitExpression limited(strtoull(argv[4], nullptr, 10), strtoull(argv[5], nullptr, 10), std::move(ordered), gq);
      return pull(&limited);
    }
    fprintf(stderr, "bad arguments\n");
    return 2;
  } catch (const std::exception& e) {
    printf("ERROR!\n%s\n", e.what());
    return 1;
  }
}
