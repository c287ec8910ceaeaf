/**
 * evqlref - a command line front-end to the UNMODIFIED reference engine.
 * (TEST INFRASTRUCTURE - oracle/_ref only; never on the product path.)
 *
 * This is our own code; it only *calls* the reference the way its own tools do:
 *   - `sql`   mirrors test/sql_tests.cc:232-274 (runTest) and the timing loop of
 *             cli/benchmarks/local_sql.cc:249 (evqlbench local-sql): default runtime,
 *             one CSTableScanProvider per table, buildQueryPlan + execute.
 *   - `write` drives cstable::CSTableWriter the way cstable_test.cc:587-650 does, so
 *             synthetic tables used by the parity tests are reference-authentic.
 *
 * Output format of `sql` (one line per row, fields separated by ';'):
 *   #<name>:<type>;...           header
 *   uint64/timestamp64 -> decimal, int64 -> decimal, float64 -> %.17g, bool -> true|false,
 *   string -> raw bytes, NULL -> NULL
 * Timing lines go to stderr:  TIMING rep=<i> ms=<t>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <string>
#include <vector>
#include <eventql/sql/runtime/defaultruntime.h>
#include <eventql/sql/runtime/runtime.h>
#include <eventql/sql/runtime/query_cache.h>
#include <eventql/transport/native/frames/query_partialaggr_result.h>
#include <eventql/server/sql/partition_cursor.h>
#include <eventql/db/partition_snapshot.h>
#include <eventql/db/database.h>
#include <eventql/db/file_tracker.h>
#include <eventql/util/io/outputstream.h>
#include <arpa/inet.h>
#include <eventql/sql/CSTableScanProvider.h>
#include <eventql/sql/result_cursor.h>
#include <eventql/sql/query_plan.h>
#include <eventql/sql/svalue.h>
#include <eventql/sql/scheduler.h>
#include <eventql/sql/transaction.h>
#include <eventql/sql/runtime/compiler.h>
#include <eventql/sql/qtree/GroupByNode.h>
#include <eventql/sql/statements/select/groupby.h>
#include <eventql/util/SHA1.h>
#include <eventql/io/cstable/cstable_writer.h>
#include <eventql/io/cstable/cstable_reader.h>
#include <eventql/io/cstable/TableSchema.h>
#include <eventql/util/io/fileutil.h>

namespace evqlref {
void registerExtensionAggregates(csql::SymbolTable* sym);
}

static int usage() {
  fprintf(stderr,
      "usage:\n"
      "  evqlref sql [-t name=file.cst]... [-n reps] [-x] [-P] [-H] -q 'SQL'\n"
      "      -H  print string values as x<hex> (NULL stays NULL)\n"
      "      -F <file> [-M <soft max bytes>]  with -P: instead of printing, write the QUERY_PARTIALAGGR_RESULT frames the server op\n"
      "                sends for the result (loop of transport/native/ops/query_partialaggr.cc:83-124 over the reference's\n"
      "                QueryPartialAggrResultFrame; 8-byte headers as TCPConnection::writeFrameHeaderAsync writes them)\n"
      "      -S <dir>:<name>[:su],<name>[:su],...  table `t` = one partition scanned by the reference's PartitionCursor; its LSM\n"
      "                tables are <dir>/<name>.cst, newest first; s = has_skiplist, u = has_updates\n"
      "      -C <dir>  with -P: install the reference's QueryCache on <dir>; the partial operator stores its .qc file there\n"
      "      -x  do NOT register the extension aggregates (min/max/mean/sum<float64>)\n"
      "  evqlref write <out.cst> <v1|v2> <nrows> <name>:<uint|datetime|float|bool|string>:<encoding>:<optional 0|1>:<datafile>[:<nullfile>] ...\n"
      "      datafile = nrows x 8 B little-endian (u64 / double bits / 0|1); nullfile = nrows x 1 B (1 = NULL)\n"
      "                 string columns: nrows x ([u32 length][bytes]) (an entry is read for NULL rows too)\n"
      "      encoding = leb128 | uint64 | uint32 | bitpacked | ieee754 | boolean | string\n"
      "  evqlref info <file.cst>\n");
  return 2;
}

static std::string fmtValue(csql::SType type, const void* data) {
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
    case csql::SType::BOOL: {
      if (p[1] & csql::STAG_NULL) return "NULL";
      return p[0] ? "true" : "false";
    }
    case csql::SType::STRING: {
      uint32_t len; memcpy(&len, p, 4);
      if (p[4 + len] & csql::STAG_NULL) return "NULL";
      return std::string((const char*) p + 4, len);
    }
    case csql::SType::NIL:
      return "NULL";
  }
  return "?";
}

// `sql -P`: every GROUP BY of the plan runs as the reference's PartialGroupByExpression (the shard side of a cluster query,
// sql/statements/select/groupby.cc:223-472) instead of GroupByExpression; its rows are (20-byte group key, saved states),
// printed as hex.  This is what a shard puts on the wire and into the query cache.
// `sql -P -C <dir>`: the reference's own query cache (sql/runtime/query_cache.cc) is installed and the partial operator's
// input is given a cache key (in the server eventql::TableScan::getCacheKey supplies one, server/sql/table_scan.cc:173;
// FastCSTableScan has none), so that PartialGroupByExpression::execute stores its `.qc` file (groupby.cc:411-432).
class CacheKeyedExpression : public csql::TableExpression {
public:
  CacheKeyedExpression(ScopedPtr<csql::TableExpression> input, SHA1Hash key) : input_(std::move(input)), key_(key) {}
  ReturnCode execute() override { return input_->execute(); }
  ReturnCode nextBatch(csql::SVector* columns, size_t* len) override { return input_->nextBatch(columns, len); }
  size_t getColumnCount() const override { return input_->getColumnCount(); }
  csql::SType getColumnType(size_t idx) const override { return input_->getColumnType(idx); }
  Option<SHA1Hash> getCacheKey() const override { return Some(key_); }
private:
  ScopedPtr<csql::TableExpression> input_;
  SHA1Hash key_;
};

static bool g_cache_keyed = false;

// `sql -S <dir>:<newest>[:flags],...`: table `t` is ONE partition whose on-disk LSM tables are the listed cstable files
// (<dir>/<name>.cst; flags s = has_skiplist, u = has_updates), newest first - the order the reference's PartitionCursor
// visits them (server/sql/partition_cursor.cc:128-155 walks PartitionState::lsm_tables back to front).  The scan of `t` is
// the reference's own eventql::PartitionCursor over a hand-built PartitionSnapshot (no arenas), i.e. its visibility filter
// loop (partition_cursor.cc:157-194) + one FastCSTableScan per table.
class PartitionProvider : public csql::TableProvider {
public:
  PartitionProvider(const std::string& dir, const std::vector<std::string>& newest_first) :
      schema_provider_("t", dir + "/" + newest_first[0].substr(0, newest_first[0].find(':')) + ".cst") {
    eventql::PartitionState state;
    state.set_partition_key(std::string(20, '\0'));
    // the snapshot counts references on its files through DatabaseContext::file_tracker (db/partition_snapshot.cc:59-68)
    memset(&dbctx_, 0, sizeof(dbctx_));
    file_tracker_.reset(new eventql::FileTracker(dir));
    dbctx_.file_tracker = file_tracker_.get();
    // lsm_tables: oldest first
    for (size_t i = newest_first.size(); i-- > 0; ) {
      std::string name = newest_first[i], flags;
      auto c = name.find(':');
      if (c != std::string::npos) { flags = name.substr(c + 1); name = name.substr(0, c); }
      auto* ref = state.add_lsm_tables();
      ref->set_filename(name);
      ref->set_first_sequence(1);
      ref->set_last_sequence(1);
      ref->set_has_skiplist(flags.find('s') != std::string::npos);
      ref->set_has_updates(flags.find('u') != std::string::npos);
    }
    snap_ = new eventql::PartitionSnapshot(state, dir, "", &dbctx_, 0);
  }
  Option<ScopedPtr<csql::TableExpression>> buildSequentialScan(
      csql::Transaction* txn,
      csql::ExecutionContext* execution_context,
      RefPtr<csql::SequentialScanNode> seqscan) const override {
    if (seqscan->tableName() != "t") return None<ScopedPtr<csql::TableExpression>>();
    return Option<ScopedPtr<csql::TableExpression>>(ScopedPtr<csql::TableExpression>(
        new eventql::PartitionCursor(txn, execution_context, nullptr, snap_, seqscan)));
  }
  void listTables(Function<void (const csql::TableInfo& table)> fn) const override { schema_provider_.listTables(fn); }
  Option<csql::TableInfo> describe(const String& table_name) const override { return schema_provider_.describe(table_name); }
private:
  csql::CSTableScanProvider schema_provider_;
  eventql::DatabaseContext dbctx_;
  ScopedPtr<eventql::FileTracker> file_tracker_;
  RefPtr<eventql::PartitionSnapshot> snap_;
};

class PartialScheduler : public csql::DefaultScheduler {
protected:
  ScopedPtr<csql::TableExpression> buildGroupByExpression(
      csql::Transaction* txn,
      csql::ExecutionContext* execution_context,
      RefPtr<csql::GroupByNode> node) override {
    Vector<csql::ValueExpression> select_expressions;
    Vector<csql::ValueExpression> group_expressions;
    for (const auto& slnode : node->selectList()) {
      select_expressions.emplace_back(txn->getCompiler()->buildValueExpression(txn, slnode->expression()));
    }
    for (const auto& e : node->groupExpressions()) {
      group_expressions.emplace_back(txn->getCompiler()->buildValueExpression(txn, e));
    }
    return mkScoped(
        new csql::PartialGroupByExpression(
            txn,
            std::move(select_expressions),
            std::move(group_expressions),
            SHA1::compute(std::string("evqlref")),
            g_cache_keyed
                ? ScopedPtr<csql::TableExpression>(new CacheKeyedExpression(
                      buildTableExpression(txn, execution_context, node->inputTable().asInstanceOf<csql::TableExpressionNode>()),
                      SHA1::compute(std::string("evqlref-input"))))
                : buildTableExpression(txn, execution_context, node->inputTable().asInstanceOf<csql::TableExpressionNode>())));
  }
};

static std::string hexString(const void* data) {
  const uint8_t* p = (const uint8_t*) data;
  uint32_t len; memcpy(&len, p, 4);
  static const char* digits = "0123456789abcdef";
  std::string out;
  for (uint32_t i = 0; i < len; ++i) { out += digits[p[4 + i] >> 4]; out += digits[p[4 + i] & 15]; }
  return out;
}

static std::vector<std::string> split(const std::string& s, char c);

static int cmdSql(int argc, char** argv) {
  std::vector<std::pair<std::string, std::string>> tables;
  std::string query;
  int reps = 1;
  bool ext = true;
  bool partial = false;
  bool hexstr = false;
  std::string cache_dir;
  std::string frames_file;
  std::string partition_spec;
  size_t frame_soft_max = 1024 * 1024 * 8;   // kPartialAggrResponseSoftMaxSize (transport/native/ops/query_partialaggr.cc:39)
  for (int i = 0; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "-t" && i + 1 < argc) {
      std::string spec = argv[++i];
      auto eq = spec.find('=');
      if (eq == std::string::npos) return usage();
      tables.emplace_back(spec.substr(0, eq), spec.substr(eq + 1));
    } else if (a == "-q" && i + 1 < argc) {
      query = argv[++i];
    } else if (a == "-n" && i + 1 < argc) {
      reps = atoi(argv[++i]);
    } else if (a == "-x") {
      ext = false;
    } else if (a == "-P") {
      partial = true;
    } else if (a == "-H") {
      hexstr = true;
    } else if (a == "-C" && i + 1 < argc) {
      cache_dir = argv[++i];
    } else if (a == "-S" && i + 1 < argc) {
      partition_spec = argv[++i];
    } else if (a == "-F" && i + 1 < argc) {
      frames_file = argv[++i];
    } else if (a == "-M" && i + 1 < argc) {
      frame_soft_max = strtoull(argv[++i], NULL, 10);
    } else {
      return usage();
    }
  }
  if (query.empty()) return usage();

  auto runtime = csql::Runtime::getDefaultRuntime();
  if (ext) {
    evqlref::registerExtensionAggregates(runtime->symbols());
  }
  if (partial) {
    runtime->setScheduler(mkScoped<csql::Scheduler>(new PartialScheduler()));
  }
  if (!cache_dir.empty()) {
    g_cache_keyed = true;
    // store on the first execution: cache_store_minhits = 0
    runtime->setQueryCache(new csql::QueryCache(cache_dir, csql::QueryCache::kDefaultAssocCacheSize, 0));
  }

  for (int rep = 0; rep < reps; ++rep) {
    try {
      auto txn = runtime->newTransaction();
      auto repo = mkScoped(new csql::TableRepository());
      for (const auto& t : tables) {
        repo->addProvider(new csql::CSTableScanProvider(t.first, t.second));
      }
      if (!partition_spec.empty()) {
        auto colon = partition_spec.find(':');
        if (colon == std::string::npos) return usage();
        repo->addProvider(new PartitionProvider(partition_spec.substr(0, colon), split(partition_spec.substr(colon + 1), ',')));
      }
      txn->setTableProvider(std::move(repo));

      auto t0 = std::chrono::steady_clock::now();
      auto qplan = runtime->buildQueryPlan(txn.get(), query);
      auto cursor = qplan->execute(0);
      size_t ncols = cursor->getColumnCount();
      bool print = (rep == reps - 1);
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
      if (!frames_file.empty()) {
        // the result loop of performOperation_QUERY_PARTIALAGGR (transport/native/ops/query_partialaggr.cc:83-124), with the
        // connection replaced by a file
        FILE* ff = fopen(frames_file.c_str(), "wb");
        if (!ff) { perror(frames_file.c_str()); return 1; }
        for (bool eof = false; !eof; ) {
          eventql::native_transport::QueryPartialAggrResultFrame result_frame;
          auto os = StringOutputStream::fromString(&result_frame.getBody());
          size_t num_rows = 0;
          while ((eof = !cursor->isValid()) == false) {
            ++num_rows;
            ++nrows;
            os->appendString(cursor->getColumnString(0));
            os->appendString(cursor->getColumnString(1));
            auto rc = cursor->next();
            if (!rc.isSuccess()) { fprintf(stdout, "ERROR!\n%s\n", rc.getMessage().c_str()); return 1; }
            if (result_frame.getBody().size() > frame_soft_max) break;
          }
          result_frame.setNumRows(num_rows);
          std::string payload;
          auto payload_os = StringOutputStream::fromString(&payload);
          result_frame.writeTo(payload_os.get());
          // TCPConnection::writeFrameHeaderAsync (transport/native/connection_tcp.cc:238-251)
          uint16_t opcode_n = htons(EVQL_OP_QUERY_PARTIALAGGR_RESULT);
          uint16_t flags_n = htons(eof ? EVQL_ENDOFREQUEST : 0);
          uint32_t len_n = htonl(payload.size());
          fwrite(&opcode_n, 2, 1, ff); fwrite(&flags_n, 2, 1, ff); fwrite(&len_n, 4, 1, ff);
          fwrite(payload.data(), 1, payload.size(), ff);
        }
        fclose(ff);
      }
      while (frames_file.empty() && cursor->isValid()) {
        if (print) {
          std::string line;
          for (size_t i = 0; i < ncols; ++i) {
            if (i) line += ";";
            if (partial && cursor->getColumnType(i) == csql::SType::STRING) line += hexString(cursor->getColumnData(i));
            else if (hexstr && cursor->getColumnType(i) == csql::SType::STRING) {
              const uint8_t* sp = (const uint8_t*) cursor->getColumnData(i);
              uint32_t slen; memcpy(&slen, sp, 4);
              line += (sp[4 + slen] & csql::STAG_NULL) ? std::string("NULL") : "x" + hexString(sp);
            }
            else line += fmtValue(cursor->getColumnType(i), cursor->getColumnData(i));
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
      fprintf(stderr, "TIMING rep=%d ms=%.3f rows_out=%zu\n", rep,
          std::chrono::duration<double, std::milli>(t1 - t0).count(), nrows);
    } catch (const std::exception& e) {
      fprintf(stdout, "ERROR!\n%s\n", e.what());
      return 1;
    }
  }
  return 0;
}

struct ColSpec {
  std::string name, type, enc, datafile, nullfile;
  bool optional;
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

static cstable::ColumnEncoding parseEnc(const std::string& e) {
  if (e == "leb128") return cstable::ColumnEncoding::UINT64_LEB128;
  if (e == "uint64") return cstable::ColumnEncoding::UINT64_PLAIN;
  if (e == "uint32") return cstable::ColumnEncoding::UINT32_PLAIN;
  if (e == "bitpacked") return cstable::ColumnEncoding::UINT32_BITPACKED;
  if (e == "ieee754") return cstable::ColumnEncoding::FLOAT_IEEE754;
  if (e == "boolean") return cstable::ColumnEncoding::BOOLEAN_BITPACKED;
  if (e == "string") return cstable::ColumnEncoding::STRING_PLAIN;
  fprintf(stderr, "bad encoding %s\n", e.c_str());
  exit(2);
}

static int cmdWrite(int argc, char** argv) {
  if (argc < 4) return usage();
  std::string out = argv[0];
  std::string ver = argv[1];
  size_t nrows = strtoull(argv[2], NULL, 10);
  std::vector<ColSpec> cols;
  for (int i = 3; i < argc; ++i) {
    auto p = split(argv[i], ':');
    if (p.size() < 5) return usage();
    ColSpec c;
    c.name = p[0]; c.type = p[1]; c.enc = p[2]; c.optional = (p[3] == "1"); c.datafile = p[4];
    if (p.size() > 5) c.nullfile = p[5];
    cols.push_back(c);
  }

  cstable::TableSchema schema;
  for (const auto& c : cols) {
    cstable::ColumnType t;
    if (c.type == "uint") t = cstable::ColumnType::UNSIGNED_INT;
    else if (c.type == "datetime") t = cstable::ColumnType::DATETIME;
    else if (c.type == "float") t = cstable::ColumnType::FLOAT;
    else if (c.type == "bool") t = cstable::ColumnType::BOOLEAN;
    else if (c.type == "string") t = cstable::ColumnType::STRING;
    else return usage();
    schema.addColumn(c.name, t, parseEnc(c.enc), false, c.optional);
  }

  FileUtil::rm(out);
  auto writer = cstable::CSTableWriter::createFile(
      out,
      ver == "v1" ? cstable::BinaryFormatVersion::v0_1_0 : cstable::BinaryFormatVersion::v0_2_0,
      schema);

  for (const auto& c : cols) {
    auto cw = writer->getColumnWriter(c.name);
    FILE* df = fopen(c.datafile.c_str(), "rb");
    if (!df) { perror(c.datafile.c_str()); return 1; }
    FILE* nf = NULL;
    if (!c.nullfile.empty()) {
      nf = fopen(c.nullfile.c_str(), "rb");
      if (!nf) { perror(c.nullfile.c_str()); return 1; }
    }
    uint64_t dmax = c.optional ? 1 : 0;
    if (c.type == "string") {
      std::string val;
      for (size_t i = 0; i < nrows; ++i) {
        uint32_t len = 0;
        if (fread(&len, 4, 1, df) != 1) { fprintf(stderr, "short read %s\n", c.datafile.c_str()); return 1; }
        val.resize(len);
        if (len && fread(&val[0], 1, len, df) != len) { fprintf(stderr, "short read %s\n", c.datafile.c_str()); return 1; }
        uint8_t isnull = 0;
        if (nf && fread(&isnull, 1, 1, nf) != 1) { fprintf(stderr, "short read %s\n", c.nullfile.c_str()); return 1; }
        if (isnull) cw->writeNull(0, 0);
        else cw->writeString(0, dmax, val);
      }
      fclose(df);
      if (nf) fclose(nf);
      continue;
    }
    const size_t CH = 1 << 16;
    std::vector<uint64_t> buf(CH);
    std::vector<uint8_t> nbuf(CH);
    for (size_t done = 0; done < nrows; ) {
      size_t n = std::min(CH, nrows - done);
      if (fread(buf.data(), 8, n, df) != n) { fprintf(stderr, "short read %s\n", c.datafile.c_str()); return 1; }
      if (nf && fread(nbuf.data(), 1, n, nf) != n) { fprintf(stderr, "short read %s\n", c.nullfile.c_str()); return 1; }
      for (size_t i = 0; i < n; ++i) {
        if (nf && nbuf[i]) {
          cw->writeNull(0, 0);
        } else if (c.type == "float") {
          double d; memcpy(&d, &buf[i], 8);
          cw->writeFloat(0, dmax, d);
        } else if (c.type == "bool") {
          cw->writeBoolean(0, dmax, buf[i] != 0);
        } else {
          cw->writeUnsignedInt(0, dmax, buf[i]);
        }
      }
      done += n;
    }
    fclose(df);
    if (nf) fclose(nf);
  }
  writer->addRows(nrows);
  writer->commit();
  return 0;
}

static int cmdInfo(int argc, char** argv) {
  if (argc < 1) return usage();
  auto reader = cstable::CSTableReader::openFile(argv[0]);
  printf("rows=%llu\n", (unsigned long long) reader->numRecords());
  for (const auto& c : reader->columns()) {
    printf("column id=%u name=%s logical=%d storage=%d rmax=%llu dmax=%llu\n",
        (unsigned) c.column_id, c.column_name.c_str(), (int) c.logical_type, (int) c.storage_type,
        (unsigned long long) c.rlevel_max, (unsigned long long) c.dlevel_max);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) return usage();
  std::string cmd = argv[1];
  try {
    if (cmd == "sql") return cmdSql(argc - 2, argv + 2);
    if (cmd == "write") return cmdWrite(argc - 2, argv + 2);
    if (cmd == "info") return cmdInfo(argc - 2, argv + 2);
  } catch (const std::exception& e) {
    fprintf(stdout, "ERROR!\n%s\n", e.what());
    return 1;
  }
  return usage();
}
