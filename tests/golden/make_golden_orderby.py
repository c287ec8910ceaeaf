#!/usr/bin/env python3
"""Golden vectors for ORDER BY / LIMIT (tests/common.py:orderby_cases): the rows, IN ORDER, that the unmodified
reference engine (oracle/_ref/evqlref: OrderByExpression + LimitExpression over GroupByExpression / FastCSTableScan)
returns on the reference-written `mixed` table.  Asserts while generating that the oracle's order_by / limit
restatement reproduces them.  Build container only; output committed as tests/golden/ref_orderby.json.

Usage: python tests/golden/make_golden_orderby.py
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as G  # noqa: E402
from oracle import evq_oracle as O  # noqa: E402
from tests import common as T  # noqa: E402


def main():
    tmp = tempfile.mkdtemp(prefix="evqorderby")
    mk, nrows = T.GOLDEN_TABLES["mixed"]
    spec = mk()
    rp = os.path.join(tmp, "mixed.ref.cst")
    G.ref_write(rp, spec, nrows, tmp=tmp)
    f = O.read_cstable(rp)
    out = {"generator": "tests/golden/make_golden_orderby.py", "reference": "17ai/eventql v0.5.0 (oracle/_ref/evqlref)", "cases": {}}
    for name, sql, plan, specs, limit, offset, ncols in T.orderby_cases():
        types, rows = G.ref_sql([("t", rp)], sql)
        # the planner appends hidden select items for the sort expressions (their values: the group's first row,
        # groupby.cc:161-172): stored as "full_*" for the binding test, the query's own columns as "types" / "rows"
        full_types, full_rows = types, rows
        types, rows = types[:ncols], [r[:ncols] for r in rows]
        want = T.parse_ref_rows(rows, types)
        res = O.run_query([f], plan)
        if specs:
            res = O.order_by(res, specs)
        if limit is not None:
            res = O.limit(res, limit, offset)
        got = res.rows()
        assert len(got) == len(want), (name, len(got), len(want))
        for g, w in zip(got, want):
            ok, why = T.rows_equal([g], [w])
            assert ok, (name, why)
        out["cases"][name] = {"sql": sql, "types": types, "rows": rows, "full_types": full_types, "full_rows": full_rows}
        print("case %-28s rows=%d ok" % (name, len(rows)))
    with open(os.path.join(HERE, "ref_orderby.json"), "w") as fh:
        json.dump(out, fh, indent=0, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "ref_orderby.json"))


if __name__ == "__main__":
    main()
