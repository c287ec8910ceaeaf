// merge.cu - GroupByMergeExpression (sql/statements/select/groupby.cc:528-637) across GPUs.
#include "query.h"

namespace evq {

void merge_query(evqgpu_query& q) {
  if (q.ctx->nranks <= 1) {
    q.merged = true;
    return;
  }
  fail(EVQGPU_ERR_UNSUPPORTED, "multi-rank merge is not implemented yet");
}

}  // namespace evq
