"""The C++ host mirror of the reference's operator surface (eventql_b200/host/): GpuTableProvider /
GpuGroupByExpression / GpuCSTableScan pulled through TableExpression::execute + nextBatch by the evqgpu_sql driver,
the way test/sql_tests.cc drives the reference's operators."""
import ctypes
import os
import subprocess

import pytest

from eventql_b200 import capi
from tests import common as T

EXE = os.path.join(T.ROOT, "eventql_b200", "evqgpu_sql")
HOSTLIB = os.path.join(T.ROOT, "eventql_b200", "libevqhost.so")
GOLD = os.path.join(T.ROOT, "tests", "golden")


def run_sql(*args):
    r = subprocess.run([EXE] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    lines = [l for l in r.stdout.split("\n") if l]
    return r.returncode, lines, r.stderr


def test_host_library_builds_and_loads(native_lib):
    assert os.path.exists(HOSTLIB) and os.path.exists(EXE)
    ctypes.CDLL(capi.LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    ctypes.CDLL(HOSTLIB)
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", HOSTLIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    for sym in ("evql_b200::GpuGroupByExpression::execute()", "evql_b200::GpuCSTableScan::execute()",
                "evql_b200::GpuQueryExpression::nextBatch(csql::SVector*, unsigned long*)",
                "evql_b200::GpuTableProvider::buildSequentialScan", "evql_b200::GpuTableProvider::buildGroupByExpression",
                "evql_b200::translate("):
        assert sym in out, sym


def test_operators_fail_loudly_without_a_device(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    rc, lines, _ = run_sql("q1", os.path.join(GOLD, "testtbl.cst"))
    assert rc == 1 and lines[0] == "ERROR!" and "no CUDA device" in lines[1]


def _parse(lines, types):
    return T.parse_ref_rows([l.split(";") for l in lines], types)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,case,table", [("q1", "q1_lineitem_leb", "lineitem_leb"), ("q6", "q6_lineitem_leb", "lineitem_leb"),
                                             ("q1", "q1_lineitem_plain", "lineitem_plain"), ("q1", "q1_lineitem_null", "lineitem_null")])
def test_fused_groupby_operator_equals_reference_rows(native_lib, mode, case, table):
    rc, lines, err = run_sql(mode, T.golden_table_path(table))
    assert rc == 0, (lines, err)
    g = T.golden()[case]
    T.check_against_golden(case, _parse(lines, g["types"]), ordered=False)


@pytest.mark.gpu
def test_partitions_through_the_provider(native_lib, tmp_path):
    """three partition files through one GpuGroupByExpression == the golden single-file result of the same rows"""
    spec = T.lineitem_spec()
    n = T.GOLDEN_TABLES["lineitem_leb"][1]
    cuts = [0, 1000, 70_001, n]
    files = []
    for i in range(3):
        p = str(tmp_path / ("part%d.cst" % i))
        T.write_table(p, spec, cuts[i + 1] - cuts[i], row_offset=cuts[i])
        files.append(p)
    rc, lines, err = run_sql("q1", *files)
    assert rc == 0, (lines, err)
    g = T.golden()["q1_lineitem_leb"]
    T.check_against_golden("q1_lineitem_leb", _parse(lines, g["types"]), ordered=False)


@pytest.mark.gpu
def test_scan_and_aggregates_on_the_reference_fixture(native_lib):
    fx = os.path.join(GOLD, "testtbl.cst")
    rc, lines, err = run_sql("scan", fx, "time")
    assert rc == 0, err
    want = [l for l in open(os.path.join(GOLD, "sql_00001.result.txt")).read().split("\n")[1:] if l.strip()]
    assert lines == want                       # test/sql/00001: 213 rows, table order, through nextBatch batches
    rc, lines, err = run_sql("count", fx, "time")
    assert rc == 0, err
    g = T.golden()["c1_global"]
    T.check_against_golden("c1_global", _parse(lines, g["types"]), ordered=False)


@pytest.mark.gpu
def test_order_by_limit_and_row_filter_operators(native_lib):
    """GpuLimitExpression over GpuOrderByExpression over the fused GROUP BY, and GpuCSTableScan::setFilter, on the
    reference's fixture: against the 213 golden `time` values of test/sql/00001"""
    fx = os.path.join(GOLD, "testtbl.cst")
    times = [int(l) for l in open(os.path.join(GOLD, "sql_00001.result.txt")).read().split("\n")[1:] if l.strip()]
    groups = {}
    for x in times:
        c, s = groups.get(x, (0, 0))
        groups[x] = (c + 1, s + x)
    want = [(k, groups[k][0], groups[k][1]) for k in sorted(groups, reverse=True)]
    rc, lines, err = run_sql("top", fx, "time", "7", "3")
    assert rc == 0, (lines, err)
    assert _parse(lines, ["uint64"] * 3) == want[3:10]
    rc, lines, err = run_sql("top", fx, "time", "100000", "0")
    assert rc == 0 and _parse(lines, ["uint64"] * 3) == want
    rc, lines, err = run_sql("scanf", fx, "time", "3")
    assert rc == 0, (lines, err)
    assert [int(l) for l in lines] == times[::3]
    rc, lines, err = run_sql("scanf", fx, "time", "0")      # nothing visible
    assert rc == 0 and lines == []


@pytest.mark.gpu
def test_partial_groupby_operator_rows(native_lib):
    """GpuPartialGroupByExpression on the reference's fixture: (SHA-1 group key, saved states) rows == the oracle's
    restatement of PartialGroupByExpression, which tests/golden/ref_partial.json pins to the reference engine"""
    from eventql_b200 import plan as P
    from oracle import evq_oracle as O
    fx = os.path.join(GOLD, "testtbl.cst")
    rc, lines, err = run_sql("partial", fx, "time")
    assert rc == 0, (lines, err)
    got = [tuple(bytes.fromhex(x) for x in l.split(";")) for l in lines]
    t = P.Col(0, P.UINT64)
    plan = P.QueryPlan(["time"], [t, P.call("count", P.lit(1)), P.call("sum", t)], where=t >= 0, group=[t],
                       flags=P.QUERY_GROUPBY | P.QUERY_WIRE)
    want = O.run_partial_query([O.read_cstable(fx)], plan)
    assert len(got) == 213
    ok, why = T.partial_rows_equal(plan, got, want)
    assert ok, why
    # ... and the operator's "store cache" step: the query cache entry under the reference's file name, the same groups
    import hashlib
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        rc, lines2, err = run_sql("partial", fx, "time", d)
        assert rc == 0, (lines2, err)
        name = hashlib.sha1((("11" * 20) + ("22" * 20)).encode()).hexdigest() + ".qc"
        assert os.listdir(d) == [name]
        qc = open(os.path.join(d, name), "rb").read()
    rows2 = [tuple(bytes.fromhex(x) for x in l.split(";")) for l in lines2]
    assert qc == capi.partial_cache_encode(rows2)


@pytest.mark.gpu
def test_partition_cursor_operator(native_lib, tmp_path):
    """GpuPartitionCursor (eventql::PartitionCursor, server/sql/partition_cursor.cc:56-225): an arena with a skiplist, then
    on-disk tables newest first with the cursor's own needs_filter rule; the rows it returns, in order, are the visible rows
    of the oracle's restatement of the filter loop."""
    import numpy as np
    from eventql_b200 import plan as P
    from oracle import evq_oracle as O
    sizes = [1500, 2100, 1200, 800]
    files, cols = [], []
    for i, n in enumerate(sizes):
        p = str(tmp_path / ("seg%d.cst" % i))
        cols.append(T.write_lsm_segment(p, i, n, key_space=350))
        files.append(p)
    flags = ["a", "su", "u", ""]          # arena; has_skiplist + has_updates; has_updates; oldest table, no updates flag
    rc, lines, err = run_sql("partition", "v", *["%s:%s" % (f, fl) if fl else f for f, fl in zip(files, flags)])
    assert rc == 0, (lines, err)
    arena_skip = (np.arange(sizes[0]) % 7) == 3
    segs = [O.LsmSegment(O.read_cstable(files[0]), arena_skip, False, True),
            O.LsmSegment(O.read_cstable(files[1]), None, True, None, True, False),
            O.LsmSegment(O.read_cstable(files[2]), None, False, None, True, False),
            O.LsmSegment(O.read_cstable(files[3]), None, False, None, False, True)]
    vis = O.lsm_visibility(segs)
    want = []
    for i in range(4):
        keep = np.ones(sizes[i], dtype=bool) if vis[i] is None else vis[i]
        want += [int(x) for x in cols[i]["v"][keep]]
    assert [int(l) for l in lines] == want
    summary = [l for l in err.split("\n") if l.startswith("#visible")][0].split()[1:]
    assert summary == ["%d%s" % (sizes[i] if vis[i] is None else int(vis[i].sum()), "u" if vis[i] is None else "f") for i in range(4)]


@pytest.mark.gpu
def test_string_group_key_through_the_host_operators(native_lib):
    """A string GROUP BY key and a string predicate through GpuGroupByExpression: nextBatch appends the packed STRING
    elements ([u32 length][bytes][tag]) the reference's operators exchange; rows equal the oracle's."""
    import gzip
    import tempfile
    from eventql_b200 import plan as P
    from oracle import evq_oracle as O
    raw = gzip.open(os.path.join(GOLD, "ref_strings_v2.cst.gz")).read()
    with tempfile.NamedTemporaryFile(suffix=".cst") as tf:
        tf.write(raw)
        tf.flush()
        rc, lines, err = run_sql("strgroup", tf.name, "s_opt", "k")
    assert rc == 0, (lines, err)
    got = []
    for l in lines:
        s, c, v = l.split(";")
        got.append((None if s == "NULL" else bytes.fromhex(s[1:]), int(c), int(v)))
    k, s_opt = P.Col(0, P.UINT64), P.Col(1, P.STRING)
    plan = P.QueryPlan(["k", "s_opt"], [s_opt, P.call("count", P.lit(1)), P.call("sum", k)],
                       where=(k >= 0) & s_opt.neq(P.lit("x")), group=[s_opt])
    ok, why = T.rows_equal(got, O.run_query([O.parse_cstable(raw)], plan).rows())
    assert ok, why


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(T.LSM_CASES))
def test_partition_cursor_operator_equals_reference_partition_cursor(native_lib, tmp_path, case):
    """GpuPartitionCursor pulled through execute / nextBatch returns the rows of the reference's eventql::PartitionCursor
    (tests/golden/ref_lsm.json), in order."""
    import json
    g = json.load(open(os.path.join(GOLD, "ref_lsm.json")))["cases"][case]
    args = []
    for i, n in enumerate(T.LSM_SIZES):
        p = str(tmp_path / ("seg%d.cst" % i))
        T.write_lsm_segment(p, i, n, key_space=T.LSM_KEY_SPACE)
        m = T.LSM_CASES[case][i]
        fl = ("s" if m[0] else "") + ("u" if m[1] else "")
        args.append(p + (":" + fl if fl else ""))
    rc, lines, err = run_sql("partition", "v", *args)
    assert rc == 0, (lines, err)
    assert [int(l) for l in lines] == g["rows"]


@pytest.mark.parametrize("which", ["q6", "strings", "ifexpr"])
def test_host_translate_emits_the_same_program_as_the_plan_module(native_lib, which):
    """evql_b200::translate (qtree -> postfix evqgpu_insn program) against eventql_b200.plan.flatten for the same expression:
    same opcodes, types, argument counts, function ids (evqgpu_function_lookup of the reference's symbol strings), literal
    bits and string pool.  No device needed."""
    from eventql_b200 import plan as P
    U, S = P.UINT64, P.STRING
    c = lambda i, t=U: P.Col(i, t)
    if which == "q6":
        e = ((c(0) >= 8766) & (c(0) < 9131) & (c(1) >= 5) & (c(1) <= 7) & (c(2) < 24) & (c(3) > 0))
    elif which == "strings":
        e = c(1, S).eq(P.lit("google.de")) & P.lit("").neq(c(2, S))
    else:
        e = P.If(c(0) < 50, c(1) * 3, c(2)) > 7
    prog = P.flatten(e, lambda sym: native_lib.evqgpu_function_lookup(sym.encode()))
    rc, lines, err = run_sql("translate", which)
    assert rc == 0, err
    got = [tuple(int(x) for x in l.split()) for l in lines[:-1]]
    assert got == [tuple(i) for i in prog.insns]
    assert lines[-1].split()[1:] == ([prog.strings.hex()] if prog.strings else [])
