#!/usr/bin/env python3
"""Golden vectors for the partial-aggregation row format (tests/common.py:partial_cases): the (group key, saved states)
rows the unmodified reference engine's PartialGroupByExpression returns (oracle/_ref/evqlref sql -P, which installs a
DefaultScheduler subclass whose buildGroupByExpression builds the partial operator) on the reference-written `mixed`
table.  Asserts while generating that the oracle's restatement reproduces them (keys and integer states exact, float
states 1e-9).  Build container only; output committed as tests/golden/ref_partial.json.

Usage: python tests/golden/make_golden_partial.py
"""
import json
import os
import subprocess
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
    tmp = tempfile.mkdtemp(prefix="evqpartial")
    mk, nrows = T.GOLDEN_TABLES["mixed"]
    rp = os.path.join(tmp, "mixed.ref.cst")
    G.ref_write(rp, mk(), nrows, tmp=tmp)
    f = O.read_cstable(rp)
    # a second partition (the rows behind the first one's) and the table that holds both: what two shards return, merged like
    # GroupByMergeExpression does, must be what the reference answers on the whole table
    nrows_b = 1700
    rp_b = os.path.join(tmp, "mixed_b.ref.cst")
    rp_all = os.path.join(tmp, "mixed_all.ref.cst")
    G.ref_write(rp_b, mk(), nrows_b, tmp=tmp, row_offset=nrows)
    G.ref_write(rp_all, mk(), nrows + nrows_b, tmp=tmp)
    out = {"generator": "tests/golden/make_golden_partial.py", "reference": "17ai/eventql v0.5.0 (oracle/_ref/evqlref sql -P)", "cases": {}}
    for name, sql, plan in T.partial_cases():
        cdir = tempfile.mkdtemp(prefix="evqqc")
        r = subprocess.run([G.EVQLREF, "sql", "-P", "-C", cdir, "-t", "t=" + rp, "-q", sql], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        lines = [ln for ln in r.stdout.split("\n") if ln]
        assert r.returncode == 0 and lines[0].startswith("#") and "ERROR!" not in lines, (name, r.stdout[:400], r.stderr[-400:])
        rows = sorted(ln.split(";") for ln in lines[1:])
        want = [(bytes.fromhex(k), bytes.fromhex(d)) for k, d in rows]
        ok, why = T.partial_rows_equal(plan, O.run_partial_query([f], plan), want)
        assert ok, (name, why)
        # the query cache entry the reference's own PartialGroupByExpression::execute stored (groupby.cc:411-432)
        qcs = [x for x in os.listdir(cdir) if x.endswith(".qc")]
        assert len(qcs) == 1, qcs
        qc = open(os.path.join(cdir, qcs[0]), "rb").read()
        # ... and the QUERY_PARTIALAGGR_RESULT frames of the same result, with the reference's 8 MiB soft maximum and with a
        # small one that splits the result (evqlref sql -P -F: the server op's loop over the reference's frame class)
        frames = {}
        for soft_max in (0, 4096):
            ff = os.path.join(cdir, "frames.bin")
            fr = subprocess.run([G.EVQLREF, "sql", "-P", "-F", ff, "-t", "t=" + rp, "-q", sql] + (["-M", str(soft_max)] if soft_max else []),
                                stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            assert fr.returncode == 0, fr.stdout[-300:]
            frames[str(soft_max)] = open(ff, "rb").read().hex()
        rb = subprocess.run([G.EVQLREF, "sql", "-P", "-t", "t=" + rp_b, "-q", sql], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        rows_b = sorted(ln.split(";") for ln in rb.stdout.split("\n")[1:] if ln)
        types, whole = G.ref_sql([("t", rp_all)], sql)
        assert types != "error", (name, whole)
        merged = O.merge_partial_rows(plan, [want, [(bytes.fromhex(k), bytes.fromhex(d)) for k, d in rows_b]])
        ok, why = T.rows_equal(merged, T.parse_ref_rows(whole, types))
        assert ok, (name, "merge of the two partitions' partial rows vs the reference on the whole table", why)
        out["cases"][name] = {"sql": sql, "rows": rows, "qc_file": qcs[0], "qc": qc.hex(), "frames": frames,
                              "merge": {"partition_b_rows": nrows_b, "rows_b": rows_b, "types": types, "whole_table_rows": whole}}
        print("case %-36s groups=%d ok" % (name, len(rows)))
    with open(os.path.join(HERE, "ref_partial.json"), "w") as fh:
        json.dump(out, fh, indent=0, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "ref_partial.json"))


if __name__ == "__main__":
    main()
