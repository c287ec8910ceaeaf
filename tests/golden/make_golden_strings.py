#!/usr/bin/env python3
"""Golden vectors for STRING_PLAIN columns (SURVEY §8 a4 readString / a9 fetchColumnString), from the REFERENCE.

Runs in the build container only (needs oracle/_ref/evqlref, built by oracle/build_ref.py from /root/reference).
Committed outputs:

  tests/golden/ref_strings_v2.cst.gz   the string fixture table of tests/common.py:strings_table_columns written by the
  tests/golden/ref_strings_v1.cst.gz   reference's CSTableWriter (writeString / writeNull) in both file format versions
  tests/golden/ref_strings.json        per table and string column the values the unmodified reference engine returns for
                                       `select <col> from t where k >= 0` (FastCSTableScan::fetchColumnString), one entry per
                                       row: "NULL", "x<hex>" for short values, "sha1:<digest>:<length>" for long ones;
                                       also for the flat string columns of the reference's own test/sql_testdata/testtbl.cst

Cross-checks asserted while generating: the oracle's string decoder returns exactly the reference engine's rows on every
file; the reference returns identical rows on the oracle-written twin of the v0.2.0 table (pins the oracle's string writer).

Usage: python tests/golden/make_golden_strings.py
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import evq_oracle as O  # noqa: E402
from tests import common as T  # noqa: E402

EVQLREF = os.path.join(ROOT, "oracle", "_ref", "evqlref")
TESTTBL_STRING_COLUMNS = ["attr.referrer_url", "attr.referrer_name", "attr.referrer_campaign", "device_id",
                          "attr.customer_session_id", "session_id", "user_id"]


def digest(v):
    if v is None:
        return "NULL"
    if len(v) <= 48:
        return "x" + v.hex()
    return "sha1:%s:%d" % (hashlib.sha1(v).hexdigest(), len(v))


def ref_write(path, version, num_rows, tmp):
    args = [EVQLREF, "write", path, version, str(num_rows)]
    for name, kind, vals, nulls in T.strings_table_columns(num_rows):
        df = os.path.join(tmp, "col_%s.bin" % name)
        if kind == "string":
            with open(df, "wb") as f:
                for v in vals:
                    f.write(len(v).to_bytes(4, "little") + v)
            a = "%s:string:string:%d:%s" % (name, 1 if nulls is not None else 0, df)
        else:
            vals.astype("<u8").tofile(df)
            a = "%s:uint:leb128:0:%s" % (name, df)
        if nulls is not None:
            nf = os.path.join(tmp, "null_%s.bin" % name)
            nulls.astype(np.uint8).tofile(nf)
            a += ":" + nf
        args.append(a)
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)


def ref_column(path, alias, col, where):
    r = subprocess.run([EVQLREF, "sql", "-H", "-t", "%s=%s" % (alias, path), "-q", "select %s from %s where %s" % (col, alias, where)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stdout[-400:]
    lines = r.stdout.decode().split("\n")
    assert lines[0].startswith("#") and lines[0].endswith(":string"), lines[0]
    out = []
    for ln in lines[1:]:
        if ln == "":
            continue
        out.append(None if ln == "NULL" else bytes.fromhex(ln[1:]))
    return out


def main():
    if not os.path.exists(EVQLREF):
        raise SystemExit("build the reference first: python oracle/build_ref.py")
    tmp = tempfile.mkdtemp(prefix="evqstr")
    out = {"generator": "tests/golden/make_golden_strings.py", "reference": "17ai/eventql v0.5.0 (oracle/_ref/evqlref)",
           "tables": {}}
    n = T.STRINGS_ROWS
    for ver in ("v2", "v1"):
        rp = os.path.join(tmp, "strings_%s.cst" % ver)
        ref_write(rp, ver, n, tmp)
        f = O.read_cstable(rp)
        cols = {}
        for name, kind, vals, nulls in T.strings_table_columns(n):
            if kind != "string":
                continue
            got = ref_column(rp, "t", name, "k >= 0")
            want = [None if (nulls is not None and nulls[i]) else vals[i] for i in range(n)]
            assert got == want, (ver, name, "reference engine vs generator")
            assert O.decode_string_column(f, name) == got, (ver, name, "oracle decode vs reference engine")
            cols[name] = [digest(v) for v in got]
        out["tables"]["ref_strings_" + ver] = {"rows": n, "columns": cols}
        with open(rp, "rb") as fi, gzip.GzipFile(os.path.join(HERE, "ref_strings_%s.cst.gz" % ver), "wb", mtime=0) as fo:
            fo.write(fi.read())
        print("ref_strings_%s: %d B, data pages of s_req: %d" % (ver, os.path.getsize(rp),
              len(f.pages(f.columns["s_req"].column_id, 1)) if ver == "v2" else 0))
    # oracle-written twin (pins the oracle's string writer)
    op = os.path.join(tmp, "strings_oracle.cst")
    T.write_strings_table(op, n)
    for name, kind, vals, nulls in T.strings_table_columns(n):
        if kind == "string":
            assert [digest(v) for v in ref_column(op, "t", name, "k >= 0")] == out["tables"]["ref_strings_v2"]["columns"][name], name
    # the reference's own fixture (v0.1.0, optional string columns, dlevel_max 1 and 2)
    tt = os.path.join(HERE, "testtbl.cst")
    f = O.read_cstable(tt)
    cols = {}
    for name in TESTTBL_STRING_COLUMNS:
        got = ref_column(tt, "testtable", name, "time > 0")
        assert O.decode_string_column(f, name) == got, name
        cols[name] = [digest(v) for v in got]
    out["tables"]["testtbl"] = {"rows": f.num_rows, "columns": cols}
    # queries over string columns: rows of the reference engine (strings as digests), the oracle must agree
    out["queries"] = {}
    rp = os.path.join(tmp, "strings_v2.cst")
    f = O.read_cstable(rp)
    for name, sql, plan, ordered in T.string_query_cases():
        r = subprocess.run([EVQLREF, "sql", "-H", "-t", "t=" + rp, "-q", sql], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        lines = r.stdout.decode().split("\n")
        assert r.returncode == 0 and lines[0].startswith("#"), (name, lines[:3])
        types = [h.rsplit(":", 1)[1] for h in lines[0][1:].split(";")]
        rows = []
        for ln in lines[1:]:
            if ln == "":
                continue
            row = []
            for v, t in zip(ln.split(";"), types):
                if t == "string":
                    row.append("NULL" if v == "NULL" else digest(bytes.fromhex(v[1:])))
                else:
                    row.append(v)
            rows.append(row)
        out["queries"][name] = {"sql": sql, "types": types, "rows": rows}
        want = T.parse_string_query_rows(rows, types)
        got = [tuple(digest(v) if isinstance(v, bytes) else v for v in row) for row in O.run_query([f], plan).rows()]
        if ordered:
            assert got == want, (name, got[:3], want[:3])
        else:
            ok, why = T.rows_equal(got, want)
            assert ok, (name, why)
        print("query %-28s %d rows" % (name, len(rows)))
    # partial-aggregation rows with string group keys (evqlref sql -P: the reference's PartialGroupByExpression); the oracle's
    # restatement must reproduce them byte for byte (integer aggregates only)
    out["partial"] = {}
    for name, sql, plan in T.string_partial_cases():
        r = subprocess.run([EVQLREF, "sql", "-P", "-t", "t=" + rp, "-q", sql], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        lines = [ln for ln in r.stdout.split("\n") if ln]
        assert r.returncode == 0 and lines[0].startswith("#") and "ERROR!" not in lines, (name, r.stdout[:300], r.stderr[-300:])
        want = [(bytes.fromhex(kd.split(";")[0]), bytes.fromhex(kd.split(";")[1])) for kd in lines[1:]]
        got = O.run_partial_query([f], plan)
        assert sorted(got) == sorted(want), (name, len(got), len(want))
        out["partial"][name] = {"sql": sql, "rows": T.digest_partial_rows(want)}
        print("partial %-28s %d groups" % (name, len(want)))
    with open(os.path.join(HERE, "ref_strings.json"), "w") as fo:
        json.dump(out, fo, indent=0, sort_keys=True)
    print("wrote ref_strings.json")


if __name__ == "__main__":
    main()
