"""Parity of the CUDA path (through the C ABI, include/evqgpu.h) with the reference - runs on the B200 box.

  * every golden case: CUDA result == rows the unmodified reference engine returned (tests/golden/ref_results.json)
    and == the CPU oracle on the same table
  * column decode of every encoding, bit for bit in the packed SVector layout (FastCSTableScan::fetchColumn*)
  * ragged sizes around the 128-value block and 1024-row tile boundaries, empty tables, partitions
  * BASELINE.json's full sizes through size-independent properties (partition additivity, closed forms, idempotence)

Tolerances: integers, counts, min/max, group keys, NULL tags: bit-exact.  float64 sum/mean: 1e-9 relative
(BASELINE.json north_star; summation order differs).
"""
import gzip
import os

import numpy as np
import pytest

from eventql_b200 import capi, plan as P
from oracle import evq_oracle as O
from tests import common as T

pytestmark = pytest.mark.gpu
GOLD = os.path.join(T.ROOT, "tests", "golden")


def run_gpu(ctx, tables, plan):
    q = ctx.query(plan)
    try:
        q.execute(tables)
        return q.rows(), q.stats()
    finally:
        q.close()


def compare(got, want, ordered):
    if ordered:
        assert len(got) == len(want)
        for i, (a, b) in enumerate(zip(got, want)):
            ok, why = T.rows_equal([a], [b])
            assert ok, "row %d: %s" % (i, why)
    else:
        ok, why = T.rows_equal(got, want)
        assert ok, why


_CASES = T.all_golden_cases()


@pytest.mark.parametrize("case,tname,plan,ordered", _CASES, ids=[c[0] for c in _CASES])
def test_cuda_equals_reference_engine(gpu_ctx, case, tname, plan, ordered):
    path = T.golden_table_path(tname)
    tbl = gpu_ctx.open_table_file(path)
    g = T.golden()[case]
    try:
        if "error" in g:
            with pytest.raises(capi.EvqError) as ei:
                run_gpu(gpu_ctx, [tbl], plan)
            assert ei.value.status == 4 and g["error"] in ei.value.message
            return
        got, stats = run_gpu(gpu_ctx, [tbl], plan)
        T.check_against_golden(case, got, ordered)
        res = O.run_query([O.read_cstable(path)], plan)
        compare(got, res.rows(), ordered)
        assert stats["rows_scanned"] == res.rows_scanned and stats["rows_passed"] == res.rows_passed
        assert stats["kernel_launches"] >= 1
    finally:
        tbl.close()


def _decode_parity(ctx, tbl, f, names):
    for name in names:
        d = O.decode_column(f, name)
        st = d.sql_type
        vals = d.values.view(np.float64) if st == P.FLOAT64 else (d.values.astype(bool) if st == P.BOOL else d.values)
        want = O.pack_svector(O.Vec(st, vals, np.where(d.present, 0, 1).astype(np.uint8)))
        got = tbl.decode_column(name)
        assert got == want, "column %s" % name


@pytest.mark.parametrize("ver", ["v1", "v2"])
def test_decode_reference_written_files(gpu_ctx, ver):
    """Tables written by the reference's CSTableWriter, both file format versions, every numeric encoding."""
    raw = np.frombuffer(gzip.open(os.path.join(GOLD, "ref_mixed_%s.cst.gz" % ver)).read(), dtype=np.uint8).copy()
    tbl = gpu_ctx.open_table(raw)
    f = O.parse_cstable(raw.tobytes())
    assert tbl.num_rows == 3000
    _decode_parity(gpu_ctx, tbl, f, [s["name"] for s in T.mixed_spec()])
    tbl.close()


@pytest.mark.parametrize("null_every", [3, 0], ids=["optional", "required"])
@pytest.mark.parametrize("nrows", [0, 1, 3, 127, 128, 129, 1023, 1024, 1025, 4096, 70001])
def test_decode_and_aggregate_ragged_sizes(gpu_ctx, nrows, null_every, tmp_path):
    """optional columns run the general kernel, required ones the fast kernel (4 consecutive rows per thread)"""
    spec = T.mixed_spec(null_every=null_every)
    path = str(tmp_path / "m.cst")
    T.write_table(path, spec, nrows)
    tbl = gpu_ctx.open_table_file(path)
    f = O.read_cstable(path)
    assert tbl.num_rows == nrows
    if nrows:
        _decode_parity(gpu_ctx, tbl, f, [s["name"] for s in spec])
    for name, _sql, plan in T.semantic_queries(spec):
        if name == "div_zero" and nrows > 100:
            continue
        got, _ = run_gpu(gpu_ctx, [tbl], plan)
        compare(got, O.run_query([f], plan).rows(), not plan.is_groupby)
    tbl.close()


@pytest.mark.parametrize("nrows", [300_000, 12_345, 7])
def test_column_statistics(gpu_ctx, nrows):
    """value_bits / leb_max_len / value_min / value_max (computed when a column is loaded): the fast kernel's static types,
    the comparisons it may fold and the range checks it may drop all rest on them, so they are checked value for value"""
    spec = T.lineitem_spec() + [dict(name="wide", logical_type=P.COL_UNSIGNED_INT, encoding=P.ENC_UINT64_PLAIN, seed=77, lo=5, span=1 << 40),
                                dict(name="narrow", logical_type=P.COL_UNSIGNED_INT, encoding=P.ENC_UINT32_PLAIN, seed=78, lo=1000, span=70000),
                                dict(name="big", logical_type=P.COL_UNSIGNED_INT, encoding=P.ENC_UINT64_LEB128, seed=9, lo=0,
                                     span=(1 << 64) - 1, transform=1),
                                dict(name="opt", logical_type=P.COL_UNSIGNED_INT, encoding=P.ENC_UINT64_LEB128, seed=10, lo=3, span=100,
                                     null_every=5)]
    tbl = gpu_ctx.synthesize(nrows, spec)
    info = {c["name"]: c for c in tbl.columns()}
    assert info["flag"]["leb_max_len"] == 1 and info["quantity"]["leb_max_len"] == 1 and info["shipdate"]["leb_max_len"] == 2
    if nrows > 1000:
        assert info["price"]["leb_max_len"] == 4 and info["price"]["value_bits"] == 24
        assert info["wide"]["value_bits"] == 40
        assert info["big"]["leb_max_len"] == 10 and info["big"]["value_bits"] == 64
    for s in spec:
        v, nulls = T.synth_values(s, nrows)
        i = info[s["name"]]
        assert int(v.max()) < (1 << i["value_bits"]) or i["value_bits"] == 64
        if s["name"] == "opt":      # NULLs read as 0: the minimum must cover them
            assert i["value_min"] == 0 and i["value_max"] >= int(v.max())
        else:
            assert (i["value_min"], i["value_max"]) == (int(v.min()), int(v.max())), s["name"]
    tbl.close()


def test_predicates_decided_by_statistics(gpu_ctx, tmp_path):
    """Comparisons that the value range of a column decides are folded at kernel-generation time; the result must not change"""
    spec = T.lineitem_spec()
    n = 50_000
    tbl = gpu_ctx.synthesize(n, spec)
    f = str(tmp_path / "t.cst")
    tbl.write_file(f)
    ref = O.read_cstable(f)
    c, names = T.cols_of(spec)

    def q(where):
        return P.QueryPlan(names, [c["flag"], P.call("count", P.lit(1)), P.call("sum", c["price"]), P.call("max", c["quantity"])],
                           where=where, group=[c["flag"]])

    wheres = [
        c["price"] > 89999,                                   # always true (the minimum is >= 90000)
        c["price"] > 90000,                                   # not decided by the (coarsened) range
        c["price"] < 65536,                                   # always false -> no groups
        (c["quantity"] >= 1) & (c["quantity"] <= 50),         # both true
        (c["quantity"] > 50) | c["discount"].eq(11),          # both false
        c["tax"].neq(9),                                      # always true
        (P.lit(100) - c["discount"]) >= 90,                   # range arithmetic: always true
        (c["price"] * c["quantity"]) > (1 << 40),             # always false
        c["shipdate"] <= 10471,                               # undecided
        (c["price"] / (c["quantity"] - c["quantity"])) >= 0,  # "always true" but raises: must not be folded away
    ]
    for w in wheres[:-1]:
        plan = q(w)
        got, _ = run_gpu(gpu_ctx, [tbl], plan)
        compare(got, O.run_query([ref], plan).rows(), False)
    with pytest.raises(capi.EvqError):
        run_gpu(gpu_ctx, [tbl], q(wheres[-1]))
    tbl.close()


def test_testtbl_v010_fixture_through_cuda(gpu_ctx):
    """C1: the reference's fixture (v0.1.0) + its golden file test/sql/00001 through the CUDA decode path."""
    tbl = gpu_ctx.open_table_file(os.path.join(GOLD, "testtbl.cst"))
    lines = open(os.path.join(GOLD, "sql_00001.result.txt")).read().split("\n")
    times = [int(x) for x in lines[1:] if x.strip()]
    got, _ = run_gpu(gpu_ctx, [tbl], P.QueryPlan(["time"], [P.Col(0, P.UINT64)], flags=0))
    assert [r[0] for r in got] == times
    got, _ = run_gpu(gpu_ctx, [tbl], P.QueryPlan(["time"], [P.call("count", P.lit(1))]))
    assert got == [(213,)]
    tbl.close()


def test_partitions_are_scanned_as_one_table(gpu_ctx, tmp_path):
    """PartitionCursor semantics: a scan over several segment files is the concatenation of the scans."""
    spec = T.lineitem_spec(null_every=5)
    sizes = [10_000, 1, 2_047, 33_333]
    tables, files, off = [], [], 0
    for i, n in enumerate(sizes):
        p = str(tmp_path / ("p%d.cst" % i))
        T.write_table(p, spec, n, row_offset=off)
        off += n
        tables.append(gpu_ctx.open_table_file(p))
        files.append(O.read_cstable(p))
    for qf in (T.q1, T.q6):
        _sql, plan = qf(spec)
        got, stats = run_gpu(gpu_ctx, tables, plan)
        compare(got, O.run_query(files, plan).rows(), False)
        assert stats["rows_scanned"] == sum(sizes)
    # scan-only plans keep table order across partitions
    c, names = T.cols_of(spec)
    plan = P.QueryPlan(names, [c["price"], c["shipdate"]], where=c["quantity"] < 3, flags=0)
    got, _ = run_gpu(gpu_ctx, tables, plan)
    compare(got, O.run_query(files, plan).rows(), True)
    for t in tables:
        t.close()


@pytest.mark.parametrize("sizes", [(30_000, 5_000), (2_047, 1_025, 1), (1_024, 3_000)])
def test_partitions_with_different_length_profiles(gpu_ctx, tmp_path, sizes):
    """A LEB128 column whose values all have one length carries no sub-index; another partition of the same query where
    the lengths differ does.  The kernel specialised for the set of partitions must decode both (odd tile counts exercise
    the last, half-filled pipeline stage of the several-tiles-per-stage layout)."""
    def spec(lo, span):
        return [dict(name="k", logical_type=P.COL_UNSIGNED_INT, encoding=P.ENC_UINT64_LEB128, seed=5, lo=0, span=3),
                dict(name="x", logical_type=P.COL_UNSIGNED_INT, encoding=P.ENC_UINT64_LEB128, seed=6, lo=lo, span=span),
                dict(name="y", logical_type=P.COL_UNSIGNED_INT, encoding=P.ENC_UINT64_LEB128, seed=7, lo=100_000, span=3_000_000)]
    profiles = [(200, 800), (0, 1000), (16384, 100)]      # all 2 bytes / 1-2 bytes / all 3 bytes
    tables, files = [], []
    for i, n in enumerate(sizes):
        sp = spec(*profiles[i % len(profiles)])
        t = gpu_ctx.synthesize(n, sp, row_offset=10_000 * i)
        f = str(tmp_path / ("p%d.cst" % i))
        t.write_file(f)
        tables.append(t)
        files.append(O.read_cstable(f))
    c, names = T.cols_of(spec(0, 1))
    plan = P.QueryPlan(names, [c["k"], P.call("count", P.lit(1)), P.call("sum", c["x"]), P.call("sum", c["x"] * c["k"]),
                               P.call("max", c["y"]), P.call("sum", c["y"])], where=c["x"] < 900_000, group=[c["k"]])
    for subset in ([0], [1] if len(tables) > 1 else [0], list(range(len(tables)))):
        got, _ = run_gpu(gpu_ctx, [tables[i] for i in subset], plan)
        compare(got, O.run_query([files[i] for i in subset], plan).rows(), False)
    scan = P.QueryPlan(names, [c["x"], c["y"]], where=c["k"].eq(1), flags=0)
    got, _ = run_gpu(gpu_ctx, tables, scan)
    compare(got, O.run_query(files, scan).rows(), True)
    for t in tables:
        t.close()


@pytest.mark.parametrize("null_every", [0, 5], ids=["required", "optional"])
def test_external_row_filter(gpu_ctx, tmp_path, null_every):
    """FastCSTableScan::setFilter (CSTableScan.cc:826-833, 1006-1009): the LSM visibility bitmap is ANDed with WHERE, in the
    fast and the general kernel, for scan-only plans (table order kept), the dense and the hash tier, over partitions of
    which only some carry a filter; WHERE still runs (and raises) on rows the filter drops."""
    spec = T.lineitem_spec(null_every=null_every)
    sizes = [5_000, 2_049, 1_024]
    tables, files, keeps = [], [], []
    rng = np.random.default_rng(11)
    for i, n in enumerate(sizes):
        t = gpu_ctx.synthesize(n, spec, row_offset=100_000 * i)
        f = str(tmp_path / ("f%d.cst" % i))
        t.write_file(f)
        keep = rng.random(n) < (0.6 if i == 0 else 0.1)
        if i == 2:
            keep = None                       # a partition without a filter keeps all its rows
        else:
            t.set_filter(keep)
        tables.append(t)
        files.append(O.read_cstable(f))
        keeps.append(np.ones(n, dtype=bool) if keep is None else keep)
    row_filter = np.concatenate(keeps)
    c, names = T.cols_of(spec)
    plans = [
        (T.q1(spec)[1], False),
        (T.q6(spec)[1], False),
        (P.QueryPlan(names, [c["price"], c["shipdate"], c["flag"]], where=c["quantity"] < 30, flags=0), True),
        (P.QueryPlan(names, [c["price"], P.call("count", P.lit(1)), P.call("sum", c["quantity"])], where=c["discount"] < 9,
                     group=[c["price"]], expected_groups=20_000), False),
    ]
    for plan, ordered in plans:
        got, stats = run_gpu(gpu_ctx, tables, plan)
        want = O.run_query(files, plan, row_filter=row_filter)
        compare(got, want.rows(), ordered)
    # one table, then the filter removed again
    got, _ = run_gpu(gpu_ctx, tables[:1], plans[0][0])
    compare(got, O.run_query(files[:1], plans[0][0], row_filter=keeps[0]).rows(), False)
    tables[0].set_filter(None)
    got, _ = run_gpu(gpu_ctx, tables[:1], plans[0][0])
    compare(got, O.run_query(files[:1], plans[0][0]).rows(), False)
    # WHERE runs on the rows the filter drops too: a division by zero there still raises (math.cc:136-143)
    tables[0].set_filter(np.zeros(sizes[0], dtype=bool))
    bad = P.QueryPlan(names, [P.call("count", P.lit(1))], where=(c["price"] / (c["quantity"] - c["quantity"])) > 0, group=[])
    with pytest.raises(capi.EvqError):
        run_gpu(gpu_ctx, tables[:1], bad)
    with pytest.raises(capi.EvqError):
        tables[1].set_filter(np.ones(7, dtype=bool))     # wrong length
    for t in tables:
        t.close()


@pytest.mark.parametrize("case", T.orderby_cases(), ids=[c[0] for c in T.orderby_cases()])
def test_order_by_limit_equals_reference_engine(gpu_ctx, case):
    """evqgpu_query_order_by / evqgpu_query_limit (device radix sort + gather over the packed result columns) against the
    rows, in order, of the reference's OrderByExpression / LimitExpression (tests/golden/ref_orderby.json) and the oracle"""
    import json
    name, _sql, plan, specs, limit, offset, _ncols = case
    with open(os.path.join(GOLD, "ref_orderby.json")) as fh:
        g = json.load(fh)["cases"][name]
    path = T.golden_table_path("mixed")
    tbl = gpu_ctx.open_table_file(path)
    q = gpu_ctx.query(plan)
    try:
        q.execute([tbl])
        if specs:
            q.order_by(specs)
        if limit is not None:
            q.limit(limit, offset)
        got = q.rows()
        # the packed columns the shim appends, bit for bit against the oracle's
        res = O.run_query([O.read_cstable(path)], plan)
        if specs:
            res = O.order_by(res, specs)
        if limit is not None:
            res = O.limit(res, limit, offset)
    finally:
        q.close()
        tbl.close()
    want = T.parse_ref_rows(g["rows"], g["types"])
    compare(got, want, True)
    compare(got, res.rows(), True)


@pytest.mark.parametrize("shape", ["dense", "hash", "global", "wire"])
def test_first_row_items(gpu_ctx, tmp_path, shape):
    """Non-aggregate select items that are NOT functions of the GROUP BY key - what the reference's planner appends as hidden
    columns for `ORDER BY <expression>` - take the value (and NULL tag) of the group's FIRST row in table order
    (groupby.cc:161-172): one 128-bit CAS on (row ordinal | tag, value) per improving row.  Every tier, optional columns,
    two partitions (table order = partition order)."""
    spec = T.mixed_spec()
    c, names = T.cols_of(spec)
    cnt = P.call("count", P.lit(1))
    files = []
    for i, n in enumerate((33_000, 21_000)):
        p = str(tmp_path / ("m%d.cst" % i))
        T.write_table(p, spec, n, row_offset=i * 40_000)
        files.append(p)
    if shape == "dense":
        k0, k1 = c["b"] % 7, c["d"] % 3
        plan = P.QueryPlan(names, [k0, k1, cnt, P.call("sum", c["c"]), c["b"], c["d"], c["a"] + 1, c["f"], c["bo"]],
                           where=(c["b"] >= 0), group=[k0, k1])
    elif shape == "global":
        plan = P.QueryPlan(names, [cnt, c["a"], c["k"], c["t"]], where=(c["b"] >= 3))
    elif shape == "wire":
        plan = P.QueryPlan(names, [c["b"] % 5, cnt, c["a"], c["d"]], where=(c["b"] >= 0), group=[c["b"] % 5], flags=P.QUERY_GROUPBY | P.QUERY_WIRE)
    else:
        plan = P.QueryPlan(names, [c["c"] % 5003, cnt, P.call("max", c["b"]), c["a"], c["big"], c["k"]], where=(c["b"] >= 0),
                           group=[c["c"] % 5003], expected_groups=1 << 25)
    tables = [gpu_ctx.open_table_file(p) for p in files]
    try:
        q = gpu_ctx.query(plan)
        try:
            q.execute(tables)
            got = q.rows()
            st = q.stats()
            part = q.fetch_partial() if shape == "wire" else None
        finally:
            q.close()
        fs = [O.read_cstable(p) for p in files]
        want = O.run_query(fs, plan).rows()
        compare(got, want, False)
        assert st["strategy"] == {"dense": 1, "global": 1, "wire": 1, "hash": 2}[shape]
        if part is not None:
            ok, why = T.partial_rows_equal(plan, part, O.run_partial_query(fs, plan))
            assert ok, why
    finally:
        for t in tables:
            t.close()


@pytest.mark.parametrize("form", ["smem_slices", "l2_slices"])
@pytest.mark.parametrize("variant", ["uniform", "two_keys_nullable", "odd_record_words", "one_word_record", "skewed_falls_back"])
def test_partitioned_hash_aggregation(gpu_ctx, tmp_path, monkeypatch, variant, form):
    """Hash tier with a group table far beyond L2: pass 1 writes the passing rows as records into partitions by the top bits of
    their group's home slot.  Default form: the records are partitioned once more until a sub-partition's table slice fits
    shared memory, and every slice is aggregated there (evq_repart + evq_agg_smem); the other form aggregates one first-level
    partition (one L2-resident table slice) at a time.  Forced here on small tables (EVQGPU_PART_MIN_MB / _SLICE_MB); a
    partition that overflows (every row the same key) falls back to the direct tier."""
    monkeypatch.setenv("EVQGPU_PART_MIN_MB", "0")
    monkeypatch.setenv("EVQGPU_PART_SLICE_MB", "1")
    if form == "l2_slices":
        monkeypatch.setenv("EVQGPU_NO_SMEM_SLICES", "1")
    cnt = P.call("count", P.lit(1))
    if variant == "uniform":
        spec = T.events_spec(5000)
        _sql, plan = T.q_highcard(spec, expected_groups=1 << 20)
        sizes = (70_001, 33_000)
    else:
        spec = T.mixed_spec()
        c, names = T.cols_of(spec)
        if variant == "odd_record_words":   # three required record columns: records of 3 words (the 8-byte copy paths)
            key = c["big"] / 1_000_003
            plan = P.QueryPlan(names, [key, cnt, P.call("sum", c["b"]), P.call("max", c["t"])], where=c["bo"] | (c["b"] < 50),
                               group=[key], expected_groups=1 << 20)
        elif variant == "one_word_record":   # a 40-bit NULL-able key + its tag + a 7-bit argument: the packed record is ONE word
            # (a > 0 drops the NULLs, which read as 0: one key with a seventh of the rows would - rightly - overflow its partition)
            plan = P.QueryPlan(names, [c["a"], cnt, P.call("sum", c["b"])], where=(c["b"] >= 0) & (c["a"] > 0), group=[c["a"]],
                               expected_groups=1 << 20)
        elif variant == "two_keys_nullable":
            key = c["big"] / 1_000_003   # (spans far more than a direct-addressed array takes: the hash tier)
            plan = P.QueryPlan(names, [key, c["k"], cnt, P.call("sum", c["a"]), P.call("min", c["f"]), P.call("max", c["d"]),
                                       P.call("mean", c["big"])], where=c["b"] >= 1, group=[key, c["k"]], expected_groups=1 << 20)
        else:
            key = c["big"] * (c["b"] / 99)   # 0 for 99 % of the rows, a full-range value otherwise: one partition takes nearly all records
            plan = P.QueryPlan(names, [key, cnt, P.call("sum", c["c"])], where=c["b"] >= 0, group=[key], expected_groups=1 << 20)
        sizes = (50_000, 21_000)
    files = []
    for i, n in enumerate(sizes):
        p = str(tmp_path / ("t%d.cst" % i))
        T.write_table(p, spec, n, row_offset=i * 100_000)
        files.append(p)
    tables = [gpu_ctx.open_table_file(p) for p in files]
    try:
        q = gpu_ctx.query(plan)
        try:
            for _ in range(2):
                q.execute(tables)
                got = q.rows()
                st = q.stats()
        finally:
            q.close()
        want = O.run_query([O.read_cstable(p) for p in files], plan).rows()
        compare(got, want, False)
        assert st["strategy"] == (2 if variant == "skewed_falls_back" else 4), st
        assert st["rows_passed"] == sum(r[2 if variant == "two_keys_nullable" else 1] for r in want)
    finally:
        for t in tables:
            t.close()


def test_order_by_large_result_properties(gpu_ctx):
    """ORDER BY over a 2 M-group result (hash tier): sortedness, stability across equal keys, permutation of the unsorted
    rows, LIMIT / OFFSET windows; multi-key and descending orders against numpy's lexsort."""
    n, nkeys = 4_000_000, 2_000_000
    spec = T.events_spec(nkeys)
    tbl = gpu_ctx.synthesize(n, spec)
    c, names = T.cols_of(spec)
    plan = P.QueryPlan(names, [c["ekey"], P.call("count", P.lit(1)), P.call("sum", c["v"]), c["ekey"] % 5], where=c["v"] >= 0,
                       group=[c["ekey"]], expected_groups=nkeys)
    q = gpu_ctx.query(plan)
    q.execute([tbl])
    base = q.rows()
    q.order_by([(0, False)])
    rows = q.rows()
    keys = [r[0] for r in rows]
    assert keys == sorted(keys) and sorted(rows) == sorted(base)
    # (count desc, key % 5 asc): ties keep the previous (key ascending) order - the sort is stable
    q.order_by([(1, True), (3, False)])
    rows2 = q.rows()
    want = sorted(rows, key=lambda r: (-r[1], r[3]))      # python's sort is stable too
    assert rows2 == want
    q.limit(1000, 500)
    assert q.rows() == want[500:1500]
    q.limit(10**9, 990)
    assert q.rows() == want[1490:1500]
    q.limit(5, 100)
    assert q.rows() == []
    q.close()
    tbl.close()


def test_count_distinct_set_grows(gpu_ctx, monkeypatch):
    """count_distinct whose (group, value) set starts too small (forced to 4096 slots): it fills up, is grown x4 until the
    pairs fit and the query re-run; result == numpy's unique count per group"""
    n = 300_000
    spec = [dict(name="g", logical_type=P.COL_UNSIGNED_INT, encoding=P.ENC_UINT64_LEB128, seed=31, lo=0, span=3),
            dict(name="x", logical_type=P.COL_UNSIGNED_INT, encoding=P.ENC_UINT64_PLAIN, seed=32, lo=0, span=1 << 26)]
    tbl = gpu_ctx.synthesize(n, spec)
    c, names = T.cols_of(spec)
    plan = P.QueryPlan(names, [c["g"], P.call("count_distinct", c["x"]), P.call("count", P.lit(1))], where=c["x"] >= 0, group=[c["g"]])
    monkeypatch.setenv("EVQGPU_DT_CAP", "4096")
    got, stats = run_gpu(gpu_ctx, [tbl], plan)
    g, _ = T.synth_values(spec[0], n)
    x, _ = T.synth_values(spec[1], n)
    want = []
    for k in range(3):
        m = g == k
        want.append((k, int(np.unique(x[m]).size), int(m.sum())))
    assert sorted(got) == want
    tbl.close()


@pytest.mark.parametrize("case", T.partial_cases(), ids=[c[0] for c in T.partial_cases()])
def test_partial_rows_equal_reference_engine(gpu_ctx, case):
    """evqgpu_query_fetch_partial: the groups in the reference's partial-aggregation row format (SHA-1 of the key tuple
    computed in the emit kernel + saved states) against the rows of the reference's PartialGroupByExpression
    (tests/golden/ref_partial.json) and the oracle; the ordinary rows of the same execution stay available"""
    import json
    name, _sql, plan = case
    with open(os.path.join(GOLD, "ref_partial.json")) as fh:
        g = json.load(fh)["cases"][name]
    want = [(bytes.fromhex(k), bytes.fromhex(d)) for k, d in g["rows"]]
    path = T.golden_table_path("mixed")
    tbl = gpu_ctx.open_table_file(path)
    q = gpu_ctx.query(plan)
    import tempfile
    try:
        q.execute([tbl])
        got = q.fetch_partial()
        rows = q.rows()
        # ... and as the query cache entry the reference's partial operator stores (groupby.cc:411-432)
        with tempfile.TemporaryDirectory() as d:
            qc_path = os.path.join(d, g["qc_file"])
            q.store_cache(qc_path)
            qc = open(qc_path, "rb").read()
            assert os.listdir(d) == [g["qc_file"]]
    finally:
        q.close()
        tbl.close()
    ok, why = T.partial_rows_equal(plan, got, want)
    assert ok, why
    assert qc == capi.partial_cache_encode(got)
    assert qc[0] == 1 and int.from_bytes(qc[1:9], "little") == len(want) and len(qc) == len(bytes.fromhex(g["qc"]))
    f = O.read_cstable(path)
    ok, why = T.partial_rows_equal(plan, got, O.run_partial_query([f], plan))
    assert ok, why
    compare(rows, O.run_query([f], plan).rows(), False)


@pytest.mark.parametrize("case", T.partial_cases(), ids=[c[0] for c in T.partial_cases()])
def test_coordinator_merges_cpu_shard_rows_on_the_device(gpu_ctx, case):
    """The coordinator's side (GroupByMergeExpression, groupby.cc:528-637) on the device: the partial rows the REFERENCE's CPU
    PartialGroupByExpression returned for two partitions (tests/golden/ref_partial.json) are parsed (loadInstanceState /
    SValue::decode), merged in a device hash table keyed by the 20-byte group keys and emitted - equal to the rows the
    reference engine returns on the table that holds both partitions, and to the oracle's restatement of the merge.  Also a
    mixed cluster: one shard's rows from the CPU reference, the other's from a GPU shard (evqgpu_query_fetch_partial)."""
    import json
    name, _sql, plan = case
    with open(os.path.join(GOLD, "ref_partial.json")) as fh:
        g = json.load(fh)["cases"][name]
    rows_a = [(bytes.fromhex(k), bytes.fromhex(d)) for k, d in g["rows"]]
    rows_b = [(bytes.fromhex(k), bytes.fromhex(d)) for k, d in g["merge"]["rows_b"]]
    want = T.parse_ref_rows(g["merge"]["whole_table_rows"], g["merge"]["types"])
    cplan = P.QueryPlan(plan.input_columns, plan.select, where=plan.where, group=plan.group, flags=P.QUERY_GROUPBY | P.QUERY_COORDINATOR)
    q = gpu_ctx.query(cplan)
    try:
        # (the rows arrive the way a coordinator gets them: as QUERY_PARTIALAGGR_RESULT frames, located with the plan)
        frames = capi.partial_frames_encode(rows_a, 4096)
        located = capi.partial_frames_decode(plan, frames)
        assert located[0] == rows_a
        q.merge_rows(rows_a)
        q.merge_rows(rows_b)
        q.merge_finish()
        got = q.rows()
        assert q.stats()["rows_scanned"] == len(rows_a) + len(rows_b)
        # ... and again with the same query object, shards in the other order
        q.merge_rows(rows_b)
        q.merge_rows(rows_a)
        q.merge_finish()
        again = q.rows()
    finally:
        q.close()
    compare(got, want, False)
    compare(again, want, False)
    compare(got, O.merge_partial_rows(plan, [rows_a, rows_b]), False)
    # mixed cluster: partition A scanned by a GPU shard, partition B's rows from the CPU reference
    path = T.golden_table_path("mixed")
    tbl = gpu_ctx.open_table_file(path)
    shard = gpu_ctx.query(plan)
    q = gpu_ctx.query(cplan)
    try:
        shard.execute([tbl])
        q.merge_rows(shard.fetch_partial())
        q.merge_rows(rows_b)
        q.merge_finish()
        mixed = q.rows()
    finally:
        q.close()
        shard.close()
        tbl.close()
    compare(mixed, want, False)


def test_partial_rows_need_the_wire_flag(gpu_ctx):
    spec = T.lineitem_spec()
    tbl = gpu_ctx.synthesize(5000, spec)
    _sql, plan = T.q1(spec)
    q = gpu_ctx.query(plan)
    q.execute([tbl])
    with pytest.raises(capi.EvqError):
        q.fetch_partial()
    q.close()
    plan.flags |= P.QUERY_WIRE
    q = gpu_ctx.query(plan)
    q.execute([tbl])
    assert len(q.fetch_partial()) == 4
    q.order_by([(0, True)])
    with pytest.raises(capi.EvqError):
        q.fetch_partial()          # the key hashes no longer line up with the reordered rows
    q.close()
    tbl.close()


def test_device_generator_matches_numpy(gpu_ctx):
    """The synthetic tables of bench.py are generated on the device; pin the generator to tests/common.py:synth_values
    (same splitmix64 definition) through the CUDA decode path, for every encoding, with a row offset."""
    spec = T.mixed_spec(null_every=7)
    n, off = 50_000, 123_457
    tbl = gpu_ctx.synthesize(n, [s for s in spec if not (s["encoding"] == P.ENC_UINT32_BITPACKED and s.get("null_every"))], row_offset=off)
    for s in spec:
        want, nulls = T.synth_values(s, n, row_offset=off)
        if s["encoding"] in (P.ENC_UINT32_BITPACKED, P.ENC_UINT32_PLAIN):
            want = want & np.uint64(0xFFFFFFFF)
        st = T.sql_type_of(s)
        vals = want.view(np.float64) if st == P.FLOAT64 else (want.astype(bool) if st == P.BOOL else want)
        packed = O.pack_svector(O.Vec(st, vals, nulls.astype(np.uint8)))
        assert tbl.decode_column(s["name"]) == packed, s["name"]
    tbl.close()


def test_device_written_file_is_read_by_the_oracle(gpu_ctx, tmp_path):
    """evqgpu_table_write_file produces a v0.2.0 cstable the reference format reader (oracle restatement) accepts."""
    spec = [s for s in T.mixed_spec(null_every=7) if not (s["encoding"] == P.ENC_UINT32_BITPACKED and s.get("null_every"))]
    n = 200_000
    tbl = gpu_ctx.synthesize(n, spec)
    p = str(tmp_path / "dev.cst")
    tbl.write_file(p)
    f = O.read_cstable(p)
    assert f.num_rows == n
    for s in spec:
        want, nulls = T.synth_values(s, n)
        d = O.decode_column(f, s["name"])
        assert np.array_equal(d.present, ~nulls), s["name"]
        got = d.values.view(np.uint64) if d.values.dtype == np.float64 else d.values.astype(np.uint64)
        if s["encoding"] in (P.ENC_UINT32_BITPACKED, P.ENC_UINT32_PLAIN):
            want = want & np.uint64(0xFFFFFFFF)
        assert np.array_equal(got[~nulls], want[~nulls]), s["name"]
    tbl.close()


def test_group_table_grows_when_the_hint_is_too_small(gpu_ctx):
    spec = T.events_spec(200_000)
    n = 400_000
    tbl = gpu_ctx.synthesize(n, spec)
    _sql, plan = T.q_highcard(spec, expected_groups=100)       # 100 -> table of 1024 slots, needs ~180 K
    got, stats = run_gpu(gpu_ctx, [tbl], plan)
    key, nulls = T.synth_values(spec[0], n)
    v, _ = T.synth_values(spec[1], n)
    uk, inv = np.unique(key, return_inverse=True)
    cnt = np.bincount(inv)
    sm = np.bincount(inv, weights=v.astype(np.float64)).astype(np.uint64)
    want = [(int(k), int(c), int(s), float(s) / int(c)) for k, c, s in zip(uk.tolist(), cnt.tolist(), sm.tolist())]
    compare(got, want, False)
    assert stats["strategy"] == 2 and stats["num_groups"] == len(uk)
    tbl.close()


def test_errors_are_loud(gpu_ctx, tmp_path):
    spec = T.lineitem_spec()
    tbl = gpu_ctx.synthesize(1000, spec)
    c, names = T.cols_of(spec)
    with pytest.raises(capi.EvqError) as ei:      # unknown column
        run_gpu(gpu_ctx, [tbl], P.QueryPlan(["nope"], [P.call("count", P.lit(1))], where=P.Col(0, P.UINT64) > 0))
    assert ei.value.status == 1
    with pytest.raises(capi.EvqError) as ei:      # modulo by zero raises like math.cc:176-216
        run_gpu(gpu_ctx, [tbl], P.QueryPlan(names, [P.call("sum", c["price"] % c["flag"])], where=c["price"] > 0))
    assert ei.value.status == 4 and "modulo by zero" in ei.value.message
    with pytest.raises(capi.EvqError):            # not a cstable
        gpu_ctx.open_table(np.zeros(1000, dtype=np.uint8))
    tbl.close()
    # repeated columns and queries over string columns are outside the device scan path: loud, not silently skipped
    # (flat string columns load and decode: tests/test_strings_lsm.py)
    t2 = gpu_ctx.open_table_file(os.path.join(GOLD, "testtbl.cst"))
    with pytest.raises(capi.EvqError) as ei:
        t2.load(["event.search_query.query_string"])
    assert ei.value.status == 2
    with pytest.raises(capi.EvqError) as ei:
        run_gpu(gpu_ctx, [t2], P.QueryPlan(["session_id"], [P.call("count", P.lit(1))], where=P.Col(0, P.UINT64) > 0))
    assert ei.value.status == 2
    t2.close()


def test_corrupt_cstable_files_are_refused(gpu_ctx, tmp_path):
    """Truncated files and file-supplied offsets / sizes that would wrap a 64-bit sum end in ERR_FORMAT, not in reads
    beyond the buffer (the range checks are written overflow-safe: size > n || off > n - size)."""
    import hashlib
    import struct
    p = str(tmp_path / "ok.cst")
    T.write_table(p, T.lineitem_spec(), 5000)
    good = np.fromfile(p, dtype=np.uint8)
    gpu_ctx.open_table(good).close()
    FORMAT = 5   # EVQGPU_ERR_FORMAT
    for cut in (10, 100, 250, len(good) // 2, len(good) - 1):
        with pytest.raises(capi.EvqError) as ei:
            t = gpu_ctx.open_table(good[:cut].copy())
            t.load(["price"])
        assert ei.value.status == FORMAT, (cut, ei.value.message)
    # the page index offset of both metablocks (cstable.cc:64-76: [txid u64][rows u64][index offset u64][index size u32][sha1]):
    # offset + size wraps around 2^64 and would pass a naive `off + size > nbytes`
    for ioff in (2 ** 64 - 8, 2 ** 64 - 1, len(good) + 1):
        bad = good.copy()
        for k in range(2):
            at = 14 + 48 * k
            blk = bytearray(bad[at:at + 28].tobytes())
            blk[16:24] = struct.pack("<Q", ioff)
            bad[at:at + 28] = np.frombuffer(bytes(blk), dtype=np.uint8)
            bad[at + 28:at + 48] = np.frombuffer(hashlib.sha1(bytes(blk)).digest(), dtype=np.uint8)
        with pytest.raises(capi.EvqError) as ei:
            gpu_ctx.open_table(bad)
        assert ei.value.status == FORMAT, (ioff, ei.value.message)


# ---- BASELINE.json sizes: size-independent properties ------------------------------------------------------------------

def _merge_partials(rows_list, plan):
    """Merge per-partition results of a count/sum plan by key (GroupByMergeExpression semantics, groupby.cc:577-612)."""
    nk = len(plan.group)
    acc = {}
    for rows in rows_list:
        for r in rows:
            k = r[:nk]
            if k not in acc:
                acc[k] = list(r[nk:])
            else:
                acc[k] = [a + b for a, b in zip(acc[k], r[nk:])]
    return [k + tuple(v) for k, v in acc.items()]


def test_q6_100m_rows_properties(gpu_ctx):
    """C2: 100 M-row lineitem, Q6.  (a) whole table == sum over 4 partitions generated with row offsets,
    (b) a closed-form column, (c) idempotence, (d) the first 2 M rows against the oracle."""
    spec = T.lineitem_spec()
    n = 100_000_000
    _sql, plan = T.q6(spec)
    whole = gpu_ctx.synthesize(n, spec)
    got, stats = run_gpu(gpu_ctx, [whole], plan)
    assert stats["rows_scanned"] == n and len(got) == 1
    again, _ = run_gpu(gpu_ctx, [whole], plan)
    assert again == got
    whole.close()
    parts = [gpu_ctx.synthesize(n // 4, spec, row_offset=i * (n // 4)) for i in range(4)]
    per = [run_gpu(gpu_ctx, [p], plan)[0] for p in parts]
    assert _merge_partials(per, plan) == got
    allp, _ = run_gpu(gpu_ctx, parts, plan)
    assert allp == got
    for p in parts:
        p.close()
    # selectivity of the synthetic Q6 is 1.81 % by construction
    assert abs(got[0][0] / n - 0.0181) < 0.001
    m = 2_000_000
    inputs = [O.Vec(P.UINT64, T.synth_values(s, m)[0], np.zeros(m, dtype=np.uint8)) for s in spec]
    small = gpu_ctx.synthesize(m, spec)
    got_small, _ = run_gpu(gpu_ctx, [small], plan)
    assert got_small == O.run_query_on(inputs, m, plan).rows()
    small.close()


def test_q1_250m_rows_properties(gpu_ctx):
    """C3 shape (Q1, 4 groups, 8 aggregates + 3 means): two 125 M-row partitions as in the 8-file layout."""
    spec = T.lineitem_spec()
    n = 125_000_000
    _sql, plan = T.q1(spec)
    _sql, plan_int = T.q1(spec, means=False)
    parts = [gpu_ctx.synthesize(n, spec, row_offset=i * n) for i in range(2)]
    both, stats = run_gpu(gpu_ctx, parts, plan_int)
    assert stats["rows_scanned"] == 2 * n and stats["strategy"] == 1 and len(both) == 4
    per = [run_gpu(gpu_ctx, [p], plan_int)[0] for p in parts]
    compare(_merge_partials(per, plan_int), both, False)
    # counts add up to the rows that pass WHERE (shipdate <= 10471: 2436 of 2526 values); sums are bounded by the
    # value ranges of their columns
    assert sum(r[2] for r in both) == stats["rows_passed"]
    assert abs(stats["rows_passed"] / (2 * n) - 2436 / 2526) < 1e-3
    for r in both:
        assert r[2] * 1 <= r[3] <= r[2] * 50 and r[2] * 90000 <= r[4] <= r[2] * 10089999
    # means == sum / count of the same run (1e-9)
    full, _ = run_gpu(gpu_ctx, parts, plan)
    for r in full:
        assert abs(r[8] - r[3] / r[2]) <= 1e-9 * r[8] and abs(r[9] - r[4] / r[2]) <= 1e-9 * r[9]
    for p in parts:
        p.close()
    m = 3_000_000
    inputs = [O.Vec(P.UINT64, T.synth_values(s, m)[0], np.zeros(m, dtype=np.uint8)) for s in spec]
    small = gpu_ctx.synthesize(m, spec)
    compare(run_gpu(gpu_ctx, [small], plan)[0], O.run_query_on(inputs, m, plan).rows(), False)
    small.close()


def test_highcard_10m_keys_properties(gpu_ctx):
    """C4 shape: 10 M distinct full-range u64 keys over 100 M rows (hash tier): counts add up, keys are distinct,
    every key is splitmix64 of a value below 10 M, partition merge == whole."""
    spec = T.events_spec(10_000_000)
    n = 100_000_000
    c, names = T.cols_of(spec)
    plan = P.QueryPlan(names, [c["ekey"], P.call("count", P.lit(1)), P.call("sum", c["v"])], where=c["v"] >= 0,
                       group=[c["ekey"]], expected_groups=10_000_000)
    tbl = gpu_ctx.synthesize(n, spec)
    q = gpu_ctx.query(plan)
    q.execute([tbl])
    cols = q.fetch_packed()
    stats = q.stats()
    q.close()
    ng = len(cols[0]) // 9
    keys = np.ascontiguousarray(np.frombuffer(cols[0], dtype=np.uint8).reshape(ng, 9)[:, :8]).view("<u8").reshape(ng)
    cnt = np.ascontiguousarray(np.frombuffer(cols[1], dtype=np.uint8).reshape(ng, 9)[:, :8]).view("<u8").reshape(ng)
    sm = np.ascontiguousarray(np.frombuffer(cols[2], dtype=np.uint8).reshape(ng, 9)[:, :8]).view("<u8").reshape(ng)
    assert stats["strategy"] in (2, 4)   # the hash tier, filled directly or by partitioned aggregation
    assert int(cnt.sum()) == n and len(np.unique(keys)) == ng
    assert 9_990_000 < ng <= 10_000_000          # 100 M draws cover all but ~450 of the 10 M keys
    assert np.isin(keys[:100000], T.splitmix64(np.arange(10_000_000, dtype=np.uint64))).all()
    # total of sums == sum of the v column, which the Q6-style global aggregate computes independently
    tot, _ = run_gpu(gpu_ctx, [tbl], P.QueryPlan(names, [P.call("sum", c["v"])], where=c["v"] >= 0))
    assert int(sm.sum()) == tot[0][0]
    tbl.close()
    m = 2_000_000
    small_spec = T.events_spec(100_000)
    small = gpu_ctx.synthesize(m, small_spec)
    _sql, p2 = T.q_highcard(small_spec)
    inputs = [O.Vec(P.UINT64, T.synth_values(s, m)[0], np.zeros(m, dtype=np.uint8)) for s in small_spec]
    compare(run_gpu(gpu_ctx, [small], p2)[0], O.run_query_on(inputs, m, p2).rows(), False)
    small.close()


def test_timeseries_partition_properties(gpu_ctx):
    """C5 shape: 1-minute buckets x 1 K sensors; partitions of different days have disjoint groups (merge = concatenation)."""
    n = 20_000_000
    parts = [gpu_ctx.synthesize(n, T.readings_spec(d)) for d in range(2)]
    _sql, plan = T.q_timeseries(T.readings_spec(0), expected_groups=1_440_000)
    per = [run_gpu(gpu_ctx, [p], plan)[0] for p in parts]
    both, stats = run_gpu(gpu_ctx, parts, plan)
    assert stats["strategy"] == 3            # 2 x 1440 x 1000 key box: the direct-addressed group array, not the hash table
    assert len(both) == len(per[0]) + len(per[1])
    assert sorted(both) == sorted(per[0] + per[1])
    assert sum(r[2] for r in both) == 2 * n
    day0 = 1_438_041_600_000_000 // 60_000_000
    assert all(day0 <= r[0] < day0 + 1440 for r in per[0]) and all(day0 + 1440 <= r[0] < day0 + 2880 for r in per[1])
    # the hash tier must give the same groups
    os.environ["EVQGPU_NO_DENSE_GLOBAL"] = "1"
    try:
        hashed, hstats = run_gpu(gpu_ctx, parts, plan)
    finally:
        del os.environ["EVQGPU_NO_DENSE_GLOBAL"]
    assert hstats["strategy"] in (2, 4) and sorted(hashed) == sorted(both)
    for p in parts:
        p.close()
