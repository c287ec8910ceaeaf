/**
 * Link stub for csql::ChartExpression.  (TEST INFRASTRUCTURE - oracle/_ref build only.)
 *
 * The reference snapshot's own sql/extensions/chartsql/ sources do not compile
 * against its own SValue, but sql/scheduler.cc:388 references the
 * class.  DRAW statements are outside the scan-filter-aggregate path, so the stub
 * only has to satisfy the linker; executing one is an error.
 */
#include <eventql/sql/extensions/chartsql/chart_expression.h>

namespace csql {

ChartExpression::ChartExpression(
    Transaction* txn,
    RefPtr<ChartStatementNode> qtree,
    Vector<Vector<ScopedPtr<TableExpression>>> input_tables,
    Vector<Vector<RefPtr<TableExpressionNode>>> input_table_qtrees) :
    txn_(txn),
    qtree_(qtree),
    input_tables_(std::move(input_tables)),
    input_table_qtrees_(input_table_qtrees),
    counter_(0) {}

ReturnCode ChartExpression::execute() {
  return ReturnCode::error("ERUNTIME", "DRAW is not available in the oracle build");
}

ReturnCode ChartExpression::nextBatch(SVector* columns, size_t* len) {
  *len = 0;
  return ReturnCode::success();
}

size_t ChartExpression::getColumnCount() const {
  return 1;
}

SType ChartExpression::getColumnType(size_t idx) const {
  return SType::STRING;
}

} // namespace csql
