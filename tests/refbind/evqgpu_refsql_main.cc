/**
 * evqgpu_refsql - SQL TEXT through the reference's own parser and planner into the GPU operators.
 * (TEST target: built only where /root/reference and oracle/_ref/build exist; the binary travels to the GPU box.)
 *
 * The reference's Runtime (parser -> QueryPlanBuilder -> QueryPlan::execute -> ResultCursor), unmodified, with
 *   - evql_b200::refbind::GpuScheduler installed through csql::Runtime::setScheduler (sql/runtime/runtime.h:71)
 *   - one evql_b200::refbind::GpuCSTableScanProvider per table in the transaction's TableRepository
 * (eventql_b200/host/refbind/gpu_binding.{h,cc}, compiled against the reference's real headers).  Same command line
 * and output format as `evqlref sql` (oracle/ref_tools/evqlref.cc), so the stored rows of the reference engine
 * (tests/golden/ref_*.json) compare directly.  The typed extension aggregates min / max / mean / sum<float64> are
 * registered so that the planner resolves their symbols; on the GPU path their CPU bodies never run.
 *
 *   evqgpu_refsql sql [-d device] [-t name=file.cst[,file2.cst...]]... [-P] [-H] [-n reps] -q 'SQL'
 * stderr: TIMING rep=<i> ms=<t> rows_out=<n>   and   GPUPLAN fused_groupbys=<n> device_sorts=<n> heartbeats=<n> tasks=<n>/<n>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <string>
#include <vector>
#include <eventql/sql/runtime/defaultruntime.h>
#include <eventql/sql/runtime/runtime.h>
#include <eventql/sql/result_cursor.h>
#include <eventql/sql/query_plan.h>
#include <eventql/sql/svalue.h>
#include <eventql/sql/transaction.h>
#include "../../eventql_b200/host/refbind/gpu_binding.h"

namespace evqlref {
void registerExtensionAggregates(csql::SymbolTable* sym);
}

using evql_b200::refbind::GpuCSTableScanProvider;
using evql_b200::refbind::GpuDevice;
using evql_b200::refbind::GpuScheduler;

static std::string hexOf(const uint8_t* p, uint32_t len) {
  static const char* digits = "0123456789abcdef";
  std::string out;
  for (uint32_t i = 0; i < len; ++i) { out += digits[p[i] >> 4]; out += digits[p[i] & 15]; }
  return out;
}

static std::string fmtValue(csql::SType type, const void* data, bool partial, bool hexstr) {
  const uint8_t* p = (const uint8_t*) data;
  char buf[64];
  switch (type) {
    case csql::SType::UINT64:
    case csql::SType::TIMESTAMP64: {
      if (p[8] & csql::STAG_NULL) return "NULL";
      uint64_t v; memcpy(&v, p, 8);
      snprintf(buf, sizeof(buf), "%llu", (unsigned long long) v);
      return buf;
    }
    case csql::SType::INT64: {
      if (p[8] & csql::STAG_NULL) return "NULL";
      int64_t v; memcpy(&v, p, 8);
      snprintf(buf, sizeof(buf), "%lld", (long long) v);
      return buf;
    }
    case csql::SType::FLOAT64: {
      if (p[8] & csql::STAG_NULL) return "NULL";
      double v; memcpy(&v, p, 8);
      snprintf(buf, sizeof(buf), "%.17g", v);
      return buf;
    }
    case csql::SType::BOOL:
      if (p[1] & csql::STAG_NULL) return "NULL";
      return p[0] ? "true" : "false";
    case csql::SType::STRING: {
      uint32_t len; memcpy(&len, p, 4);
      if (partial) return hexOf(p + 4, len);
      if (p[4 + len] & csql::STAG_NULL) return "NULL";
      if (hexstr) return "x" + hexOf(p + 4, len);
      return std::string((const char*) p + 4, len);
    }
    case csql::SType::NIL:
      return "NULL";
  }
  return "?";
}

// `-P`: every GROUP BY runs as the shard side of a cluster query (the rows PartialGroupByExpression produces)
class PartialGpuScheduler : public GpuScheduler {
public:
  explicit PartialGpuScheduler(RefPtr<GpuDevice> gpu) : GpuScheduler(gpu) {}
protected:
  ScopedPtr<csql::TableExpression> buildGroupByExpression(csql::Transaction* txn, csql::ExecutionContext* ectx,
                                                          RefPtr<csql::GroupByNode> node) override {
    node->setIsPartialAggreagtion(true);
    return GpuScheduler::buildGroupByExpression(txn, ectx, node);
  }
};

static std::vector<std::string> split(const std::string& s, char c) {
  std::vector<std::string> out;
  size_t b = 0;
  for (;;) {
    auto e = s.find(c, b);
    if (e == std::string::npos) { out.push_back(s.substr(b)); break; }
    out.push_back(s.substr(b, e - b));
    b = e + 1;
  }
  return out;
}

static int usage() {
  fprintf(stderr, "usage: evqgpu_refsql sql [-d device] [-t name=file.cst[,file.cst...]]... [-P] [-H] [-n reps] -q 'SQL'\n");
  return 2;
}

int main(int argc, char** argv) {
  if (argc < 2 || std::string(argv[1]) != "sql") return usage();
  std::vector<std::pair<std::string, std::vector<std::string>>> tables;
  std::string query;
  int reps = 1, device = 0;
  bool partial = false, hexstr = false;
  for (int i = 2; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "-t" && i + 1 < argc) {
      std::string spec = argv[++i];
      auto eq = spec.find('=');
      if (eq == std::string::npos) return usage();
      tables.emplace_back(spec.substr(0, eq), split(spec.substr(eq + 1), ','));
    } else if (a == "-q" && i + 1 < argc) query = argv[++i];
    else if (a == "-n" && i + 1 < argc) reps = atoi(argv[++i]);
    else if (a == "-d" && i + 1 < argc) device = atoi(argv[++i]);
    else if (a == "-P") partial = true;
    else if (a == "-H") hexstr = true;
    else return usage();
  }
  if (query.empty()) return usage();
  try {
    auto runtime = csql::Runtime::getDefaultRuntime();
    evqlref::registerExtensionAggregates(runtime->symbols());
    RefPtr<GpuDevice> gpu(new GpuDevice(device));
    GpuScheduler* sched = partial ? new PartialGpuScheduler(gpu) : new GpuScheduler(gpu);
    runtime->setScheduler(ScopedPtr<csql::Scheduler>(sched));   // sql/runtime/runtime.h:71

    for (int rep = 0; rep < reps; ++rep) {
      auto txn = runtime->newTransaction();
      size_t heartbeats = 0;
      txn->setHeartbeatCallback([&heartbeats] () -> ReturnCode { ++heartbeats; return ReturnCode::success(); });
      auto repo = mkScoped(new csql::TableRepository());
      for (const auto& t : tables) repo->addProvider(new GpuCSTableScanProvider(gpu, t.first, t.second));
      txn->setTableProvider(std::move(repo));

      auto t0 = std::chrono::steady_clock::now();
      auto qplan = runtime->buildQueryPlan(txn.get(), query);
      auto cursor = qplan->execute(0);
      const size_t ncols = cursor->getColumnCount();
      const bool print = rep == reps - 1;
      if (print) {
        const auto& names = qplan->getStatementgetResultColumns(0);
        std::string hdr = "#";
        for (size_t i = 0; i < ncols; ++i) {
          if (i) hdr += ";";
          hdr += (i < names.size() ? names[i] : std::string("?"));
          hdr += ":";
          hdr += csql::sql_typename(cursor->getColumnType(i));
        }
        puts(hdr.c_str());
      }
      size_t nrows = 0;
      while (cursor->isValid()) {
        if (print) {
          std::string line;
          for (size_t i = 0; i < ncols; ++i) {
            if (i) line += ";";
            line += fmtValue(cursor->getColumnType(i), cursor->getColumnData(i), partial, hexstr);
          }
          puts(line.c_str());
        }
        ++nrows;
        auto rc = cursor->next();
        if (!rc.isSuccess()) {
          fprintf(stdout, "ERROR!\n%s\n", rc.getMessage().c_str());
          return 1;
        }
      }
      auto t1 = std::chrono::steady_clock::now();
      fprintf(stderr, "TIMING rep=%d ms=%.3f rows_out=%zu\n", rep, std::chrono::duration<double, std::milli>(t1 - t0).count(), nrows);
      if (print) {
        fprintf(stderr, "GPUPLAN fused_groupbys=%zu device_sorts=%zu heartbeats=%zu\n", sched->fusedGroupBys(), sched->deviceSorts(), heartbeats);
      }
    }
  } catch (const std::exception& e) {
    fprintf(stdout, "ERROR!\n%s\n", e.what());
    return 1;
  }
  return 0;
}
