#!/usr/bin/env python3
"""Generate the golden vectors that pin the oracle (and, through it, the CUDA path) to the REFERENCE.

Runs in the build container only (needs /root/reference compiled into oracle/_ref/evqlref by
oracle/build_ref.py).  Nothing here runs on the GPU box; its outputs are committed:

  tests/golden/ref_results.json     for every case of tests/common.py:golden_cases() the rows the unmodified
                                    reference engine returns (FastCSTableScan + GroupByExpression through its
                                    own planner), on tables written by the reference's own CSTableWriter
  tests/golden/ref_mixed_v2.cst.gz  a small table (every numeric encoding, optional columns) written by the
  tests/golden/ref_mixed_v1.cst.gz  reference's CSTableWriter in both file format versions

While generating, three cross-checks are asserted (a failure aborts the generation):
  * the oracle's decoder reads the reference-written files back to exactly the synthetic values
  * the reference returns identical rows on the oracle-written twin of every table (pins the oracle's writer)
  * the oracle's own result for every case equals the reference's (bit-exact; floats 1e-9 relative)

Usage: python tests/golden/make_golden.py
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from eventql_b200 import plan as P  # noqa: E402
from oracle import evq_oracle as O  # noqa: E402
from tests import common as T  # noqa: E402

EVQLREF = os.path.join(ROOT, "oracle", "_ref", "evqlref")

ENC_NAMES = {P.ENC_UINT64_LEB128: "leb128", P.ENC_UINT64_PLAIN: "uint64", P.ENC_UINT32_PLAIN: "uint32",
             P.ENC_UINT32_BITPACKED: "bitpacked", P.ENC_FLOAT_IEEE754: "ieee754", P.ENC_BOOLEAN_BITPACKED: "boolean"}
TYPE_NAMES = {P.COL_UNSIGNED_INT: "uint", P.COL_DATETIME: "datetime", P.COL_FLOAT: "float", P.COL_BOOLEAN: "bool"}


def ref_write(path, spec, nrows, version="v2", tmp="/tmp", row_offset=0):
    args = [EVQLREF, "write", path, version, str(nrows)]
    for s in spec:
        v, nulls = T.synth_values(s, nrows, row_offset)
        df = os.path.join(tmp, "col_%s.bin" % s["name"])
        v.astype("<u8").tofile(df)
        a = "%s:%s:%s:%d:%s" % (s["name"], TYPE_NAMES[s.get("logical_type", P.COL_UNSIGNED_INT)], ENC_NAMES[s["encoding"]],
                                1 if s.get("null_every") else 0, df)
        if s.get("null_every"):
            nf = os.path.join(tmp, "null_%s.bin" % s["name"])
            nulls.astype(np.uint8).tofile(nf)
            a += ":" + nf
        args.append(a)
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)


def ref_sql(tables, sql):
    """-> (types, rows as lists of strings) or ('error', message)"""
    args = [EVQLREF, "sql"]
    for alias, path in tables:
        args += ["-t", "%s=%s" % (alias, path)]
    args += ["-q", sql]
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    lines = r.stdout.split("\n")
    if "ERROR!" in lines:
        i = lines.index("ERROR!")
        return "error", lines[i + 1] if i + 1 < len(lines) else ""
    if r.returncode != 0:
        raise RuntimeError("evqlref failed: %s\n%s" % (sql, r.stderr))
    hdr = lines[0]
    assert hdr.startswith("#"), hdr
    types = [h.rsplit(":", 1)[1] for h in hdr[1:].split(";")]
    rows = [ln.split(";") for ln in lines[1:] if ln != ""]
    return types, rows


def check_decode(path, spec, nrows):
    f = O.read_cstable(path)
    for s in spec:
        want, nulls = T.synth_values(s, nrows)
        d = O.decode_column(f, s["name"])
        assert np.array_equal(d.present, ~nulls), (path, s["name"], "presence")
        got = d.values.astype(np.uint64) if d.values.dtype != np.float64 else d.values.view(np.uint64)
        if s["encoding"] in (P.ENC_UINT32_BITPACKED, P.ENC_UINT32_PLAIN):
            want = want & np.uint64(0xFFFFFFFF)
        assert np.array_equal(got[~nulls], want[~nulls]), (path, s["name"], "values")


def main():
    if not os.path.exists(EVQLREF):
        raise SystemExit("build the reference first: python oracle/build_ref.py")
    tmp = tempfile.mkdtemp(prefix="evqgolden")
    out = {"generator": "tests/golden/make_golden.py", "reference": "17ai/eventql v0.5.0 (oracle/_ref/evqlref)", "cases": {}}
    paths = {}
    for tname, (mk, nrows) in T.GOLDEN_TABLES.items():
        spec = mk()
        rp = os.path.join(tmp, tname + ".ref.cst")
        op = os.path.join(tmp, tname + ".oracle.cst")
        ref_write(rp, spec, nrows, tmp=tmp)
        T.write_table(op, spec, nrows)
        check_decode(rp, spec, nrows)
        check_decode(op, spec, nrows)
        paths[tname] = (rp, op, spec, nrows)
        print("table %-16s rows=%d ref=%d B oracle=%d B" % (tname, nrows, os.path.getsize(rp), os.path.getsize(op)))
    for name, tname, alias, sql, plan in T.golden_cases():
        rp, op, spec, nrows = paths[tname]
        a = ref_sql([(alias, rp)], sql)
        b = ref_sql([(alias, op)], sql)
        if a[0] == "error":
            assert b[0] == "error", name
            out["cases"][name] = {"table": tname, "sql": sql, "error": a[1]}
            try:
                O.run_query([O.read_cstable(op)], plan)
                raise AssertionError("oracle did not raise for " + name)
            except O.OracleError as e:
                print("case %-28s ERROR ref=%r oracle=%r" % (name, a[1], str(e)))
            continue
        types, rows = a
        ordered = not plan.is_groupby
        if ordered:
            assert a == b, name
        else:
            assert types == b[0] and sorted(rows) == sorted(b[1]), name
        want = T.parse_ref_rows(rows, types)
        got = O.run_query([O.read_cstable(op)], plan).rows()
        if ordered:
            assert len(got) == len(want), name
            ok, why = all(T.rows_equal([g], [w])[0] for g, w in zip(got, want)), "ordered rows differ"
        else:
            ok, why = T.rows_equal(got, want)
        assert ok, (name, why)
        if len(rows) > 6000:
            # long results: keep a digest (of the rows in table order for projections, sorted for GROUP BY, floats
            # excluded) + the first rows; the tests recompute the digest with tests/common.py:rows_digest
            h = T.rows_digest(want, types, ordered)
            out["cases"][name] = {"table": tname, "sql": sql, "types": types, "num_rows": len(rows), "sha256": h,
                                  "rows": (rows if ordered else sorted(rows))[:50]}
        else:
            out["cases"][name] = {"table": tname, "sql": sql, "types": types, "num_rows": len(rows), "rows": rows}
        print("case %-28s rows=%d ok" % (name, len(rows)))
    # C1: the reference's own fixture
    fx = os.path.join(HERE, "testtbl.cst")
    for name, sql, plan in T.testtbl_queries():
        types, rows = ref_sql([("testtable", fx)], sql)
        want = T.parse_ref_rows(rows, types)
        got = O.run_query([O.read_cstable(fx)], plan).rows()
        ok, why = T.rows_equal(got, want)
        assert ok, (name, why)
        out["cases"][name] = {"table": "testtbl.cst", "sql": sql, "types": types, "num_rows": len(rows), "rows": rows}
        print("case %-28s rows=%d ok" % (name, len(rows)))
    with open(os.path.join(HERE, "ref_results.json"), "w") as fh:
        json.dump(out, fh, indent=0, separators=(",", ":"))
    # small reference-written fixtures (both format versions)
    spec = T.mixed_spec()
    for ver in ("v2", "v1"):
        p = os.path.join(tmp, "ref_mixed_%s.cst" % ver)
        ref_write(p, spec, 3000, version=ver, tmp=tmp)
        check_decode(p, spec, 3000)
        with open(p, "rb") as fi, gzip.GzipFile(os.path.join(HERE, "ref_mixed_%s.cst.gz" % ver), "wb", mtime=0) as fo:
            fo.write(fi.read())
        print("fixture ref_mixed_%s.cst.gz: %d B raw, %d B gz" % (ver, os.path.getsize(p), os.path.getsize(os.path.join(HERE, "ref_mixed_%s.cst.gz" % ver))))
    print("wrote", os.path.join(HERE, "ref_results.json"))


if __name__ == "__main__":
    main()
