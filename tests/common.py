"""Shared test material: deterministic synthetic tables and the parity query set.

Values are a pure function of (seed, row) - r = splitmix64(seed + row); v = lo + r % span - the same
definition the device-side generator uses (eventql_b200/csrc/synth.cu), so numpy can regenerate any table.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from eventql_b200 import plan as P  # noqa: E402

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
SEED0 = 0xE7E2700


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_values(spec: dict, num_rows: int, row_offset: int = 0):
    """(values uint64 raw bits, nulls bool) of one synthetic column."""
    rows = np.arange(num_rows, dtype=np.uint64) + np.uint64(row_offset)
    with np.errstate(over="ignore"):
        r = splitmix64(np.uint64(spec.get("seed", 0)) + rows)
        v = np.uint64(spec.get("lo", 0)) + r % np.uint64(spec.get("span", 1))
    tr = spec.get("transform", 0)
    if tr == 1:
        v = splitmix64(v)
    elif tr == 2:
        v = (v.astype(np.float64) / 100.0).view(np.uint64)
    k = spec.get("null_every", 0)
    nulls = ((rows % np.uint64(k)) == np.uint64(k - 1)) if k else np.zeros(num_rows, dtype=bool)
    v = np.where(nulls, np.uint64(0), v)
    if spec.get("logical_type", P.COL_UNSIGNED_INT) == P.COL_BOOLEAN:
        v = (v > 0).astype(np.uint64)
    return v, nulls


# the lineitem-style schema of SURVEY §8(d) (C2 / C3)
def lineitem_spec(encoding=P.ENC_UINT64_LEB128, null_every=0):
    cols = [("shipdate", 8036, 2526), ("discount", 0, 11), ("quantity", 1, 50), ("price", 90000, 10000000),
            ("tax", 0, 9), ("flag", 0, 2), ("status", 0, 2)]
    return [dict(name=n, logical_type=P.COL_UNSIGNED_INT, encoding=encoding, seed=SEED0 + i, lo=lo, span=span,
                 null_every=null_every if n in ("price", "tax", "flag") else 0)
            for i, (n, lo, span) in enumerate(cols)]


def events_spec(num_keys=10_000_000, key_encoding=P.ENC_UINT64_PLAIN):   # C4
    return [dict(name="ekey", logical_type=P.COL_UNSIGNED_INT, encoding=key_encoding, seed=SEED0 + 100, lo=0, span=num_keys, transform=1),
            dict(name="v", logical_type=P.COL_UNSIGNED_INT, encoding=P.ENC_UINT64_LEB128, seed=SEED0 + 101, lo=0, span=1000000)]


def readings_spec(day: int):   # C5: one time partition per day
    return [dict(name="time", logical_type=P.COL_DATETIME, encoding=P.ENC_UINT64_LEB128, seed=SEED0 + 200,
                 lo=1_438_041_600_000_000 + day * 86_400_000_000, span=86_400_000_000),
            dict(name="sensor_id", logical_type=P.COL_UNSIGNED_INT, encoding=P.ENC_UINT64_LEB128, seed=SEED0 + 201, lo=0, span=1000),
            dict(name="value", logical_type=P.COL_UNSIGNED_INT, encoding=P.ENC_UINT64_LEB128, seed=SEED0 + 202, lo=0, span=65536)]


def mixed_spec(null_every=7):
    """Every encoding the numeric path reads, with optional columns."""
    E = P
    return [
        dict(name="a", logical_type=E.COL_UNSIGNED_INT, encoding=E.ENC_UINT64_LEB128, seed=1, lo=0, span=1 << 40, null_every=null_every),
        dict(name="b", logical_type=E.COL_UNSIGNED_INT, encoding=E.ENC_UINT64_PLAIN, seed=2, lo=0, span=100),
        dict(name="c", logical_type=E.COL_UNSIGNED_INT, encoding=E.ENC_UINT32_BITPACKED, seed=3, lo=0, span=1 << 32),
        dict(name="d", logical_type=E.COL_UNSIGNED_INT, encoding=E.ENC_UINT32_PLAIN, seed=4, lo=5, span=1000, null_every=null_every + 4 if null_every else 0),
        dict(name="f", logical_type=E.COL_FLOAT, encoding=E.ENC_FLOAT_IEEE754, seed=5, lo=0, span=100000, transform=2, null_every=null_every),
        dict(name="bo", logical_type=E.COL_BOOLEAN, encoding=E.ENC_BOOLEAN_BITPACKED, seed=6, lo=0, span=2),
        dict(name="t", logical_type=E.COL_DATETIME, encoding=E.ENC_UINT64_LEB128, seed=7, lo=1_438_041_600_000_000, span=14_400_000_000),
        dict(name="k", logical_type=E.COL_UNSIGNED_INT, encoding=E.ENC_UINT64_LEB128, seed=8, lo=0, span=3, null_every=null_every),
        dict(name="big", logical_type=E.COL_UNSIGNED_INT, encoding=E.ENC_UINT64_LEB128, seed=9, lo=0, span=(1 << 64) - 1, transform=1),
    ]


def write_table(path: str, spec: list, num_rows: int, row_offset: int = 0, interleave=True):
    """Write the synthetic table with the ORACLE's cstable writer (test infrastructure)."""
    from oracle import evq_oracle as O
    cols = []
    for s in spec:
        v, nulls = synth_values(s, num_rows, row_offset)
        cols.append(O.WriteColumn(s["name"], s.get("logical_type", P.COL_UNSIGNED_INT), s["encoding"], v,
                                  nulls if s.get("null_every", 0) else None))
    return O.write_cstable(path, num_rows, cols, interleave=interleave)


def sql_type_of(spec: dict) -> int:
    lt = spec.get("logical_type", P.COL_UNSIGNED_INT)
    return {P.COL_BOOLEAN: P.BOOL, P.COL_FLOAT: P.FLOAT64}.get(lt, P.UINT64)


def cols_of(spec: list):
    return {s["name"]: P.Col(i, sql_type_of(s)) for i, s in enumerate(spec)}, [s["name"] for s in spec]


# ---- the parity query set: (name, sql accepted by the reference planner, plan builder) ----
def q6(spec):
    c, names = cols_of(spec)
    where = ((c["shipdate"] >= 8766) & (c["shipdate"] < 9131) & (c["discount"] >= 5) & (c["discount"] <= 7) &
             (c["quantity"] < 24) & (c["price"] > 0))
    sql = ("select count(1), sum(price * discount) from lineitem where shipdate >= 8766 and shipdate < 9131 and "
           "discount >= 5 and discount <= 7 and quantity < 24 and price > 0;")
    return sql, P.QueryPlan(names, [P.call("count", P.lit(1)), P.call("sum", c["price"] * c["discount"])], where=where)


def q1(spec, means=True):
    c, names = cols_of(spec)
    where = ((c["shipdate"] <= 10471) & (c["quantity"] > 0) & (c["price"] > 0) & (c["discount"] >= 0) & (c["tax"] >= 0))
    disc = c["price"] * (P.lit(100) - c["discount"])
    sel = [c["flag"], c["status"], P.call("count", P.lit(1)), P.call("sum", c["quantity"]), P.call("sum", c["price"]),
           P.call("sum", disc), P.call("sum", disc * (P.lit(100) + c["tax"])), P.call("sum", c["discount"])]
    sql = ("select flag, status, count(1), sum(quantity), sum(price), sum(price * (100 - discount)), "
           "sum(price * (100 - discount) * (100 + tax)), sum(discount)")
    if means:
        sel += [P.call("mean", c["quantity"]), P.call("mean", c["price"]), P.call("mean", c["discount"])]
        sql += ", mean(quantity), mean(price), mean(discount)"
    sql += (" from lineitem where shipdate <= 10471 and quantity > 0 and price > 0 and discount >= 0 and tax >= 0 "
            "group by flag, status;")
    return sql, P.QueryPlan(names, sel, where=where, group=[c["flag"], c["status"]])


def q_highcard(spec, expected_groups=0):
    c, names = cols_of(spec)
    sql = "select ekey, count(1), sum(v), mean(v) from events where v >= 0 group by ekey;"   # `key` is a reserved word in csql
    return sql, P.QueryPlan(names, [c["ekey"], P.call("count", P.lit(1)), P.call("sum", c["v"]), P.call("mean", c["v"])],
                            where=c["v"] >= 0, group=[c["ekey"]], expected_groups=expected_groups)


def q_timeseries(spec, expected_groups=0):
    c, names = cols_of(spec)
    bucket = c["time"] / 60000000
    sql = ("select time / 60000000, sensor_id, count(1), sum(value) from readings where value >= 0 "
           "group by time / 60000000, sensor_id;")
    return sql, P.QueryPlan(names, [bucket, c["sensor_id"], P.call("count", P.lit(1)), P.call("sum", c["value"])],
                            where=c["value"] >= 0, group=[bucket, c["sensor_id"]], expected_groups=expected_groups)


def rows_equal(a, b, rel=1e-9):
    """Order-insensitive comparison of result rows; floats within `rel` relative tolerance (BASELINE.json)."""
    def key(row):
        return tuple((0, 0) if v is None else (1, v) if not isinstance(v, float) else (2, 0.0) for v in row)
    if len(a) != len(b):
        return False, "row count %d != %d" % (len(a), len(b))
    sa, sb = sorted(a, key=key), sorted(b, key=key)
    # float columns do not take part in the ordering: pair rows by their exact columns, then compare floats
    for ra, rb in zip(sa, sb):
        if len(ra) != len(rb):
            return False, "column count differs"
        for x, y in zip(ra, rb):
            if isinstance(x, float) or isinstance(y, float):
                if x is None or y is None:
                    if x is not y:
                        return False, "%r != %r" % (ra, rb)
                    continue
                if x != y and not (np.isnan(x) and np.isnan(y)):
                    if abs(x - y) > rel * max(abs(x), abs(y)):
                        return False, "%r != %r" % (ra, rb)
            elif x != y:
                return False, "%r != %r" % (ra, rb)
    return True, ""


# ---- semantics cases over mixed_spec(): NULL handling, floats, bools, if(), division by zero ----
def semantic_queries(spec):
    """[(name, sql, plan)] - every SQL is accepted by the reference planner."""
    c, names = cols_of(spec)
    Q = []

    def add(name, sql, select, where=None, group=(), flags=P.QUERY_GROUPBY):
        Q.append((name, sql, P.QueryPlan(names, list(select), where=where, group=list(group), flags=flags)))

    cnt = P.call("count", P.lit(1))
    add("null_group_bare", "select k, count(1), sum(k) from t where b >= 0 group by k;",
        [c["k"], cnt, P.call("sum", c["k"])], where=c["b"] >= 0, group=[c["k"]])
    add("null_group_expr", "select k + 1, count(1) from t where b >= 0 group by k + 1;",
        [c["k"] + 1, cnt], where=c["b"] >= 0, group=[c["k"] + 1])
    add("null_where_lt", "select count(1), sum(a) from t where k < 1 and a >= 0;",
        [cnt, P.call("sum", c["a"])], where=(c["k"] < 1) & (c["a"] >= 0))
    add("null_minmaxmean", "select min(a), max(a), mean(a), count(1), min(d), max(d), mean(d) from t where a >= 0 and d >= 0;",
        [P.call("min", c["a"]), P.call("max", c["a"]), P.call("mean", c["a"]), cnt, P.call("min", c["d"]),
         P.call("max", c["d"]), P.call("mean", c["d"])], where=(c["a"] >= 0) & (c["d"] >= 0))
    add("float_aggs", "select bo, sum(f), min(f), max(f), mean(f), count(1) from t where f >= 0.0 group by bo;",
        [c["bo"], P.call("sum", c["f"]), P.call("min", c["f"]), P.call("max", c["f"]), P.call("mean", c["f"]), cnt],
        where=c["f"] >= 0.0, group=[c["bo"]])
    add("float_pred", "select count(1), sum(c) from t where f * 2.0 > 700.5 and c > 0;",
        [cnt, P.call("sum", c["c"])], where=((c["f"] * 2.0) > 700.5) & (c["c"] > 0))
    add("if_expr", "select if(b < 50, 1, 2), count(1), sum(if(b < 50, c, d)) from t where b >= 0 and c >= 0 and d >= 0 group by if(b < 50, 1, 2);",
        [P.If(c["b"] < 50, P.lit(1), P.lit(2)), cnt, P.call("sum", P.If(c["b"] < 50, c["c"], c["d"]))],
        where=(c["b"] >= 0) & (c["c"] >= 0) & (c["d"] >= 0), group=[P.If(c["b"] < 50, P.lit(1), P.lit(2))])
    add("wrap_sum", "select count(1), sum(big), sum(big * big), min(big), max(big) from t where big >= 0;",
        [cnt, P.call("sum", c["big"]), P.call("sum", c["big"] * c["big"]), P.call("min", c["big"]), P.call("max", c["big"])],
        where=c["big"] >= 0)
    add("mod_div", "select b % 7, count(1), sum(c / (b + 1)) from t where b >= 0 and c >= 0 group by b % 7;",
        [c["b"] % 7, cnt, P.call("sum", c["c"] / (c["b"] + 1))], where=(c["b"] >= 0) & (c["c"] >= 0), group=[c["b"] % 7])
    add("time_bucket", "select t / 3600000000, count(1), min(t), max(t) from t where t > 0 group by t / 3600000000;",
        [c["t"] / 3600000000, cnt, P.call("min", c["t"]), P.call("max", c["t"])], where=c["t"] > 0,
        group=[c["t"] / 3600000000])
    add("or_neg", "select count(1) from t where not (b < 10 or bo) and d >= 0;",
        [cnt], where=(~((c["b"] < 10) | c["bo"])) & (c["d"] >= 0))
    add("empty_result", "select count(1), sum(b) from t where b > 1000;", [cnt, P.call("sum", c["b"])], where=c["b"] > 1000)
    add("empty_groups", "select b, count(1) from t where b > 1000 group by b;", [c["b"], cnt], where=c["b"] > 1000, group=[c["b"]])
    add("div_zero", "select count(1), sum(c / b) from t where b >= 0 and c >= 0;",
        [cnt, P.call("sum", c["c"] / c["b"])], where=(c["b"] >= 0) & (c["c"] >= 0))
    add("signed", "select count(1), sum(to_int64(b) - 50), min(to_int64(b) - 50), max(to_int64(b) - 50) from t where b >= 0;",
        [cnt, P.call("sum", P.call("sub", P.call("to_int64", c["b"]), P.lit(50))),
         P.call("min", P.call("sub", P.call("to_int64", c["b"]), P.lit(50))),
         P.call("max", P.call("sub", P.call("to_int64", c["b"]), P.lit(50)))], where=c["b"] >= 0)
    # count_distinct_uint64 (aggregate.cc:80-137): std::set of the values per group, a NULL counts as its value 0
    cd = lambda e: P.call("count_distinct", e)
    add("distinct_global", "select count_distinct(b), count(1) from t where b >= 0;", [cd(c["b"]), cnt], where=c["b"] >= 0)
    add("distinct_null_groups", "select k, count_distinct(b), count_distinct(a), count(1), sum(b) from t where b >= 0 and k >= 0 and a >= 0 group by k;",
        [c["k"], cd(c["b"]), cd(c["a"]), cnt, P.call("sum", c["b"])], where=(c["b"] >= 0) & (c["k"] >= 0) & (c["a"] >= 0), group=[c["k"]])
    add("distinct_many_groups", "select d, count_distinct(c), count(1) from t where d >= 0 and c >= 0 group by d;",
        [c["d"], cd(c["c"]), cnt], where=(c["d"] >= 0) & (c["c"] >= 0), group=[c["d"]])
    add("distinct_required_hash", "select b, count_distinct(c % 50), count_distinct(b), sum(c) from t where b >= 0 and c >= 0 group by b;",
        [c["b"], cd(c["c"] % 50), cd(c["b"]), P.call("sum", c["c"])], where=(c["b"] >= 0) & (c["c"] >= 0), group=[c["b"]])
    add("distinct_required_dense", "select b % 3, count_distinct(c % 50), sum(b), count(1) from t where b >= 0 and c >= 0 group by b % 3;",
        [c["b"] % 3, cd(c["c"] % 50), P.call("sum", c["b"]), cnt], where=(c["b"] >= 0) & (c["c"] >= 0), group=[c["b"] % 3])
    # scan-only plans (FastCSTableScan alone): filtered projection keeps table order
    add("scan_project", "select a, d, f, bo, c + d from t where b < 5;",
        [c["a"], c["d"], c["f"], c["bo"], c["c"] + c["d"]], where=c["b"] < 5, flags=0)
    add("scan_all", "select k, big from t;", [c["k"], c["big"]], flags=0)
    return Q


def testtbl_queries():
    """C1: the reference's own fixture test/sql_testdata/testtbl.cst (v0.1.0, 213 rows). Input column: time."""
    t = P.Col(0, P.UINT64)
    names = ["time"]
    cnt = P.call("count", P.lit(1))
    aggs = [cnt, P.call("sum", t), P.call("min", t), P.call("max", t)]
    return [
        ("c1_group_time", "select time, count(1), sum(time), min(time), max(time) from testtable group by time;",
         P.QueryPlan(names, [t] + aggs, group=[t])),
        ("c1_group_hour", "select time / 3600000000, count(1), sum(time), min(time), max(time) from testtable group by time / 3600000000;",
         P.QueryPlan(names, [t / 3600000000] + aggs, group=[t / 3600000000])),
        ("c1_global", "select count(1), sum(time), min(time), max(time), mean(time) from testtable where time > 0;",
         P.QueryPlan(names, aggs + [P.call("mean", t)], where=t > 0)),
        ("c1_filtered", "select count(1), sum(time) from testtable where time > 1438055000000000;",
         P.QueryPlan(names, [cnt, P.call("sum", t)], where=t > 1438055000000000)),
        ("c1_scan", "select testtable.time from testtable;", P.QueryPlan(names, [t], flags=0)),
    ]


def parse_ref_rows(rows, types):
    """Golden rows (lists of strings as printed by oracle/ref_tools/evqlref.cc) -> python tuples."""
    out = []
    for r in rows:
        vals = []
        for s, t in zip(r, types):
            if s == "NULL":
                vals.append(None)
            elif t == "float64":
                vals.append(float(s))
            elif t == "bool":
                vals.append(s == "true")
            else:
                vals.append(int(s))
        out.append(tuple(vals))
    return out


# the golden case table: name -> (spec builder, rows, query builder).  tests/golden/make_golden.py runs the
# REFERENCE on these; tests compare the oracle (CPU) and the CUDA path against the stored reference rows.
GOLDEN_TABLES = {
    "lineitem_leb": (lambda: lineitem_spec(), 120_000),
    "lineitem_plain": (lambda: lineitem_spec(P.ENC_UINT64_PLAIN), 40_000),
    "lineitem_null": (lambda: lineitem_spec(null_every=7), 60_000),
    "events": (lambda: events_spec(5000), 80_000),
    "readings": (lambda: readings_spec(0), 60_000),
    "mixed": (lambda: mixed_spec(), 70_000),
}


def golden_cases():
    """[(case name, table name, sql table alias, sql, plan)]"""
    out = []
    for tname in ("lineitem_leb", "lineitem_plain", "lineitem_null"):
        spec = GOLDEN_TABLES[tname][0]()
        sql, plan = q6(spec)
        out.append(("q6_" + tname, tname, "lineitem", sql, plan))
        sql, plan = q1(spec)
        out.append(("q1_" + tname, tname, "lineitem", sql, plan))
    spec = GOLDEN_TABLES["events"][0]()
    sql, plan = q_highcard(spec)
    out.append(("highcard_events", "events", "events", sql, plan))
    spec = GOLDEN_TABLES["readings"][0]()
    sql, plan = q_timeseries(spec)
    out.append(("timeseries_readings", "readings", "readings", sql, plan))
    spec = GOLDEN_TABLES["mixed"][0]()
    for name, sql, plan in semantic_queries(spec):
        out.append(("sem_" + name, "mixed", "t", sql, plan))
    return out


def orderby_cases():
    """ORDER BY / LIMIT over the mixed golden table: [(case, sql, plan without ORDER BY / LIMIT, sort specs
    [(result column, descending)], limit or None, offset, number of select columns)].  The sort keys are total orders
    (the reference's std::sort is not stable) and every referenced column also appears in WHERE."""
    spec = GOLDEN_TABLES["mixed"][0]()
    c, names = cols_of(spec)
    cnt = P.call("count", P.lit(1))
    out = []
    gb = P.QueryPlan(names, [c["b"], cnt, P.call("sum", c["c"])], where=(c["b"] >= 0) & (c["c"] >= 0), group=[c["b"]])
    base = "select b, count(1), sum(c) from t where b >= 0 and c >= 0 group by b"
    out.append(("ob_uint_asc_limit", base + " order by b limit 5;", gb, [(0, False)], 5, 0, 3))
    out.append(("ob_uint_desc_limit_offset", base + " order by b desc limit 4 offset 2;", gb, [(0, True)], 4, 2, 3))
    out.append(("ob_uint_all", base + " order by b;", gb, [(0, False)], None, 0, 3))
    k0, k1 = c["b"] % 7, c["d"] % 3
    two = P.QueryPlan(names, [k0, k1, cnt, P.call("sum", c["f"])], where=(c["b"] >= 0) & (c["f"] >= 0.0) & (c["d"] >= 0), group=[k0, k1])
    out.append(("ob_two_keys", "select b % 7, d % 3, count(1), sum(f) from t where b >= 0 and f >= 0.0 and d >= 0 group by b % 7, d % 3 "
                "order by b % 7 desc, d % 3 asc;", two, [(0, True), (1, False)], None, 0, 4))
    fl = P.QueryPlan(names, [c["b"], P.call("sum", c["f"])], where=(c["f"] >= 0.0) & (c["b"] >= 0), group=[c["b"]])
    out.append(("ob_float_desc", "select b, sum(f) as s from t where f >= 0.0 and b >= 0 group by b order by s desc limit 3;", fl,
                [(1, True)], 3, 0, 2))
    sk = P.call("to_int64", c["b"]) - 50
    si = P.QueryPlan(names, [sk, cnt], where=c["b"] >= 0, group=[sk])
    out.append(("ob_int64", "select to_int64(b) - 50, count(1) from t where b >= 0 group by to_int64(b) - 50 order by to_int64(b) - 50 "
                "limit 4;", si, [(0, False)], 4, 0, 2))
    sc = P.QueryPlan(names, [c["a"], c["d"]], where=(c["b"] < 2) & (c["a"] >= 0) & (c["d"] >= 0), flags=0)
    out.append(("ob_scan_desc", "select a, d from t where b < 2 and a >= 0 and d >= 0 order by a desc limit 5;", sc, [(0, True)], 5, 0, 2))
    ts = P.QueryPlan(names, [c["t"], c["b"]], where=(c["b"] < 1) & (c["t"] > 0), flags=0)
    out.append(("ob_timestamp", "select t, b from t where b < 1 and t > 0 order by t limit 3;", ts, [(0, False)], 3, 0, 2))
    lm = P.QueryPlan(names, [c["a"], c["b"]], where=(c["b"] < 5) & (c["a"] >= 0), flags=0)
    out.append(("limit_scan_offset", "select a, b from t where b < 5 and a >= 0 limit 7 offset 3;", lm, [], 7, 3, 2))
    return out


def partial_cases():
    """Partial aggregation (PartialGroupByExpression rows) over the mixed golden table: [(case, sql, plan)]; the plans carry
    QUERY_WIRE.  Every referenced column also appears in WHERE."""
    spec = GOLDEN_TABLES["mixed"][0]()
    c, names = cols_of(spec)
    cnt = P.call("count", P.lit(1))
    W = P.QUERY_GROUPBY | P.QUERY_WIRE
    out = []
    out.append(("pa_null_key", "select k, count(1), sum(b) from t where b >= 0 and k >= 0 group by k;",
                P.QueryPlan(names, [c["k"], cnt, P.call("sum", c["b"])], where=(c["b"] >= 0) & (c["k"] >= 0), group=[c["k"]], flags=W)))
    k0, k1 = c["b"] % 3, c["d"] % 2
    out.append(("pa_two_keys_all_aggs",
                "select b % 3, d % 2, count(1), sum(c), min(c), max(c), mean(c), sum(f), min(f), mean(f) from t "
                "where b >= 0 and c >= 0 and d >= 0 and f >= 0.0 group by b % 3, d % 2;",
                P.QueryPlan(names, [k0, k1, cnt, P.call("sum", c["c"]), P.call("min", c["c"]), P.call("max", c["c"]), P.call("mean", c["c"]),
                                    P.call("sum", c["f"]), P.call("min", c["f"]), P.call("mean", c["f"])],
                            where=(c["b"] >= 0) & (c["c"] >= 0) & (c["d"] >= 0) & (c["f"] >= 0.0), group=[k0, k1], flags=W)))
    out.append(("pa_global", "select count(1), sum(b) from t where b >= 0;",
                P.QueryPlan(names, [cnt, P.call("sum", c["b"])], where=c["b"] >= 0, flags=W)))
    out.append(("pa_bool_key", "select bo, count(1), max(a) from t where b >= 0 and a >= 0 group by bo;",
                P.QueryPlan(names, [c["bo"], cnt, P.call("max", c["a"])], where=(c["b"] >= 0) & (c["a"] >= 0), group=[c["bo"]], flags=W)))
    sk = P.call("to_int64", c["b"]) - 50
    out.append(("pa_int64_key", "select to_int64(b) - 50, count(1), sum(to_int64(b) - 50) from t where b < 9 group by to_int64(b) - 50;",
                P.QueryPlan(names, [sk, cnt, P.call("sum", sk)], where=c["b"] < 9, group=[sk], flags=W)))
    # count_distinct: the saved state is the value set (aggregate.cc:110-116), dense tier and hash tier
    cd = lambda e: P.call("count_distinct", e)
    out.append(("pa_count_distinct_dense", "select b % 3, count_distinct(c % 50), count_distinct(a % 97), count(1) from t where b >= 0 and c >= 0 and a >= 0 group by b % 3;",
                P.QueryPlan(names, [k0, cd(c["c"] % 50), cd(c["a"] % 97), cnt], where=(c["b"] >= 0) & (c["c"] >= 0) & (c["a"] >= 0), group=[k0], flags=W)))
    out.append(("pa_count_distinct_many_groups", "select d, count_distinct(c % 7), sum(c) from t where d >= 0 and c >= 0 group by d;",
                P.QueryPlan(names, [c["d"], cd(c["c"] % 7), P.call("sum", c["c"])], where=(c["d"] >= 0) & (c["c"] >= 0), group=[c["d"]], flags=W)))
    out.append(("pa_many_groups_key_not_selected", "select count(1), sum(c), mean(b) from t where d >= 0 and c >= 0 and b >= 0 group by d;",
                P.QueryPlan(names, [cnt, P.call("sum", c["c"]), P.call("mean", c["b"])], where=(c["d"] >= 0) & (c["c"] >= 0) & (c["b"] >= 0),
                            group=[c["d"]], flags=W)))
    return out


def string_partial_cases():
    """Partial aggregation (PartialGroupByExpression rows) with STRING group keys over the string fixture table: the key hash
    covers [u32 length][bytes][tag] of the string (groupby.cc:112-135), a string select item travels as SValue::encode.
    [(case, sql, plan)]"""
    names = ["k", "s_req", "s_opt"]
    k, s_req, s_opt = P.Col(0, P.UINT64), P.Col(1, P.STRING), P.Col(2, P.STRING)
    cnt = P.call("count", P.lit(1))
    W = P.QUERY_GROUPBY | P.QUERY_WIRE
    return [
        ("sp_null_key_selected", "select s_opt, count(1), sum(k) from t where k >= 0 group by s_opt;",
         P.QueryPlan(names, [s_opt, cnt, P.call("sum", k)], where=k >= 0, group=[s_opt], flags=W)),
        ("sp_string_and_numeric_key", "select count(1), sum(k), max(k) from t where k >= 0 group by s_req, k % 3;",
         P.QueryPlan(names, [cnt, P.call("sum", k), P.call("max", k)], where=k >= 0, group=[s_req, k % 3], flags=W)),
        ("sp_two_string_keys", "select s_req, k % 2, s_opt, count(1) from t where k >= 0 group by s_req, k % 2, s_opt;",
         P.QueryPlan(names, [s_req, k % 2, s_opt, cnt], where=k >= 0, group=[s_req, k % 2, s_opt], flags=W)),
    ]


def digest_partial_rows(rows):
    """[(key bytes, data bytes)] -> sorted [[key hex, data hex or sha1:<hex>:<length> for long data]] (the fixture has ~100 KB strings)"""
    import hashlib
    out = []
    for kb, db in rows:
        out.append([kb.hex(), db.hex() if len(db) <= 200 else "sha1:%s:%d" % (hashlib.sha1(db).hexdigest(), len(db))])
    return sorted(out)


def parse_partial_data(plan, data: bytes):
    """The saved states of one PartialGroupByExpression row as python values (floats stay floats: they are compared with the
    1e-9 tolerance, the summation order differs), following the select items of the plan."""
    import struct
    pos = 0

    def varuint():
        nonlocal pos
        v, sh = 0, 0
        while True:
            b = data[pos]
            pos += 1
            v |= (b & 0x7F) << sh
            sh += 7
            if not b & 0x80:
                return v

    out = []
    for s in plan.select:
        agg = P.find_aggregate(s)
        if agg is None:
            t = data[pos]
            pos += 1
            n = varuint()
            raw = data[pos: pos + n]
            pos += n
            out.append((t, raw.hex()))
        elif agg.name in ("count",) or (agg.name == "sum" and agg.type != P.FLOAT64):
            out.append(varuint())
        elif agg.name == "count_distinct":       # the set: its size, then its members in std::set order
            n = varuint()
            out.append(n)
            out.append(tuple(varuint() for _ in range(n)))
        elif agg.name == "sum":
            out.append(struct.unpack_from("<d", data, pos)[0])
            pos += 8
        elif agg.name in ("min", "max"):
            v, seen = struct.unpack_from("<QQ", data, pos)
            pos += 16
            out.append(struct.unpack("<d", struct.pack("<Q", v))[0] if agg.type == P.FLOAT64 else v)
            out.append(seen)
        elif agg.name == "mean":
            sm, n = struct.unpack_from("<dQ", data, pos)
            pos += 16
            out.append(sm)
            out.append(n)
        else:
            raise ValueError(agg.name)
    assert pos == len(data), (pos, len(data))
    return tuple(out)


def partial_rows_equal(plan, got, want):
    """[(key, data)] against [(key, data)] as sets of groups: keys exact, states exact except floats (1e-9 relative)."""
    a = sorted((k.hex(),) + parse_partial_data(plan, d) for k, d in got)
    b = sorted((k.hex(),) + parse_partial_data(plan, d) for k, d in want)
    if [r[0] for r in a] != [r[0] for r in b]:
        return False, "group keys differ (%d vs %d groups)" % (len(a), len(b))
    for ra, rb in zip(a, b):
        ok, why = rows_equal([ra[1:]], [rb[1:]])
        if not ok:
            return False, "group %s: %s" % (ra[0], why)
    return True, ""


def rows_digest(rows, types, ordered):
    """sha256 over the exact (non-float) columns of the rows; GROUP BY results are sorted first."""
    import hashlib
    keep = [i for i, t in enumerate(types) if t != "float64"]
    lines = [";".join("NULL" if r[i] is None else str(int(r[i])) if not isinstance(r[i], bool) else str(r[i]).lower() for i in keep)
             for r in rows]
    if not ordered:
        lines.sort()
    return hashlib.sha256("\n".join(lines).encode()).hexdigest()


_GOLDEN = None


def golden():
    """tests/golden/ref_results.json: rows returned by the REFERENCE engine (tests/golden/make_golden.py)."""
    global _GOLDEN
    if _GOLDEN is None:
        import json
        with open(os.path.join(ROOT, "tests", "golden", "ref_results.json")) as fh:
            _GOLDEN = json.load(fh)["cases"]
    return _GOLDEN


def check_against_golden(case: str, got_rows, ordered: bool):
    """Compare result rows (python tuples, None = NULL) with the reference's stored rows for `case`."""
    g = golden()[case]
    assert "error" not in g, "case %s expects an error: %s" % (case, g.get("error"))
    assert len(got_rows) == g["num_rows"], "%s: %d rows, reference returned %d" % (case, len(got_rows), g["num_rows"])
    if "sha256" in g:
        assert rows_digest(got_rows, g["types"], ordered) == g["sha256"], "%s: digest of exact columns differs" % case
        head = parse_ref_rows(g["rows"], g["types"])
        mine = got_rows if ordered else sorted(got_rows, key=lambda r: [";".join("NULL" if v is None else str(v) for v in r)])
        if ordered:
            for a, b in zip(mine[:len(head)], head):
                ok, why = rows_equal([a], [b])
                assert ok, "%s: %s" % (case, why)
        return
    want = parse_ref_rows(g["rows"], g["types"])
    if ordered:
        for i, (a, b) in enumerate(zip(got_rows, want)):
            ok, why = rows_equal([a], [b])
            assert ok, "%s row %d: %s" % (case, i, why)
    else:
        ok, why = rows_equal(got_rows, want)
        assert ok, "%s: %s" % (case, why)


_TABLE_CACHE = {}


def golden_table_path(tname: str) -> str:
    """The golden table `tname`, written (once per session) with the oracle's writer into a temp dir."""
    import tempfile
    if tname == "testtbl.cst":
        return os.path.join(ROOT, "tests", "golden", "testtbl.cst")
    if tname not in _TABLE_CACHE:
        d = _TABLE_CACHE.setdefault("__dir__", tempfile.mkdtemp(prefix="evqtables"))
        mk, nrows = GOLDEN_TABLES[tname]
        p = os.path.join(d, tname + ".cst")
        write_table(p, mk(), nrows)
        _TABLE_CACHE[tname] = p
    return _TABLE_CACHE[tname]


def all_golden_cases():
    """[(case, table name, plan, ordered)] incl. the C1 fixture cases"""
    out = [(name, tname, plan, not plan.is_groupby) for name, tname, _alias, _sql, plan in golden_cases()]
    out += [(name, "testtbl.cst", plan, not plan.is_groupby) for name, _sql, plan in testtbl_queries()]
    return out


# ---- string columns (SURVEY §8 a4 / a9 readString, fetchColumnString) and LSM segments (§8 f2) ----
_VOCAB = [b"google.de", b"googleadservices.com", b"facebook", b"newsletter", b"DE-Special-post", b"", b"x", b"\x00\xff;\n",
          b"GA2-DE-Search-Brand", b"a" * 127, b"b" * 128, b"c" * 129, "grüße".encode()]


def synth_strings(seed: int, num_rows: int, null_every: int = 0):
    """(list of bytes, nulls bool).  Lengths cover 0, 1, the 1 -> 2 byte varuint boundary (127 / 128), a few values of
    ~100 KB (3-byte length prefixes; with 512 KiB pages several values straddle a page boundary) and embedded NUL /
    separator bytes."""
    rows = np.arange(num_rows, dtype=np.uint64)
    r = splitmix64(np.uint64(seed) + rows)
    out = []
    for i in range(num_rows):
        x = int(r[i])
        w = _VOCAB[x % len(_VOCAB)]
        if i % 397 == 396:
            v = (w + b"|" + str(i).encode()) * (100_000 // (len(w) + 6))
        elif x % 7 == 0:
            v = w + b"-" + str(x % 1000).encode()
        else:
            v = w
        out.append(v)
    nulls = ((rows % np.uint64(null_every)) == np.uint64(null_every - 1)) if null_every else np.zeros(num_rows, dtype=bool)
    return out, nulls


STRINGS_ROWS = 2500


def strings_table_columns(num_rows: int = STRINGS_ROWS, seed: int = 0):
    """name -> (kind, values, nulls) of the string fixture table (tests/golden/ref_strings_v{1,2}.cst.gz are seed 0;
    other seeds give other value sets, as the partitions of different ranks have)."""
    import hashlib
    k, _ = synth_values(dict(seed=41 + seed, lo=0, span=1000), num_rows)
    s_req, _ = synth_strings(42 + 1000 * seed, num_rows)
    s_opt, nulls = synth_strings(43 + 1000 * seed, num_rows, null_every=3)
    ids = [hashlib.sha1(b"row-%d" % (int(x) % 600)).digest() for x in k]
    return [("k", "uint", k, None), ("s_req", "string", s_req, None), ("s_opt", "string", s_opt, nulls), ("id", "string", ids, None)]


def write_strings_table(path: str, num_rows: int = STRINGS_ROWS, seed: int = 0):
    """The string fixture table written by the ORACLE's writer (test infrastructure)."""
    from oracle import evq_oracle as O
    cols = []
    for name, kind, vals, nulls in strings_table_columns(num_rows, seed):
        if kind == "string":
            cols.append(O.WriteColumn(name, P.COL_STRING, P.ENC_STRING_PLAIN, None, nulls, strings=vals))
        else:
            cols.append(O.WriteColumn(name, P.COL_UNSIGNED_INT, P.ENC_UINT64_LEB128, vals, nulls))
    return O.write_cstable(path, num_rows, cols)


def lsm_segment_columns(seg: int, num_rows: int, key_space: int = 400, seed: int = 77):
    """Columns of one synthetic partition segment: the reference's bookkeeping columns (db/partition_arena.cc:41-45:
    __lsm_is_update, __lsm_skip, __lsm_id = 20 raw bytes, __lsm_version, __lsm_sequence; all required) plus a payload
    column.  Ids repeat inside a segment and across segments."""
    import hashlib
    rows = np.arange(num_rows, dtype=np.uint64)
    r = splitmix64(np.uint64(seed + 1000 * seg) + rows)
    key = r % np.uint64(key_space)
    ids = [hashlib.sha1(b"id-%d" % int(x)).digest() for x in key]
    is_update = ((r >> np.uint64(20)) % np.uint64(3) == 0).astype(np.uint64)
    skip = ((r >> np.uint64(30)) % np.uint64(10) == 0).astype(np.uint64)
    version = (r >> np.uint64(40)) % np.uint64(1000)
    seq = rows + np.uint64(1 + 100000 * seg)
    v = (r >> np.uint64(8)) % np.uint64(5000)
    return dict(ids=ids, is_update=is_update, skip=skip, version=version, seq=seq, v=v, key=key)


def write_lsm_segment(path: str, seg: int, num_rows: int, **kw):
    from oracle import evq_oracle as O
    c = lsm_segment_columns(seg, num_rows, **kw)
    cols = [O.WriteColumn("__lsm_is_update", P.COL_BOOLEAN, P.ENC_BOOLEAN_BITPACKED, c["is_update"]),
            O.WriteColumn("__lsm_skip", P.COL_BOOLEAN, P.ENC_BOOLEAN_BITPACKED, c["skip"]),
            O.WriteColumn("__lsm_id", P.COL_STRING, P.ENC_STRING_PLAIN, None, strings=c["ids"]),
            O.WriteColumn("__lsm_version", P.COL_UNSIGNED_INT, P.ENC_UINT64_LEB128, c["version"]),
            O.WriteColumn("__lsm_sequence", P.COL_UNSIGNED_INT, P.ENC_UINT64_LEB128, c["seq"]),
            O.WriteColumn("key", P.COL_UNSIGNED_INT, P.ENC_UINT64_LEB128, c["key"]),
            O.WriteColumn("v", P.COL_UNSIGNED_INT, P.ENC_UINT64_LEB128, c["v"])]
    O.write_cstable(path, num_rows, cols)
    return c


def string_query_cases():
    """[(name, sql for the reference planner, plan, ordered)] over the string fixture table (strings_table_columns):
    eq / neq between string columns and literals (NULL compares as "", boolean.cc:235-257), string GROUP BY keys (NULL and ""
    are different groups: the key bytes include the tag, groupby.cc:112-135), a projected string column."""
    names = ["k", "s_req", "s_opt"]
    k, s_req, s_opt = P.Col(0, P.UINT64), P.Col(1, P.STRING), P.Col(2, P.STRING)
    cnt = P.call("count", P.lit(1))
    Q = []

    def add(name, sql, select, where=None, group=(), flags=P.QUERY_GROUPBY, ordered=False):
        Q.append((name, sql, P.QueryPlan(names, list(select), where=where, group=list(group), flags=flags), ordered))

    add("str_eq_lit", "select count(1), sum(k) from t where s_opt = 'google.de' and k >= 0;",
        [cnt, P.call("sum", k)], where=s_opt.eq(P.lit("google.de")) & (k >= 0))
    add("str_lit_eq_col", "select count(1), sum(k) from t where 'facebook' = s_req and k >= 0;",
        [cnt, P.call("sum", k)], where=P.lit("facebook").eq(s_req) & (k >= 0))
    add("str_eq_empty_matches_null", "select count(1), sum(k) from t where s_opt = '' and k >= 0;",
        [cnt, P.call("sum", k)], where=s_opt.eq(P.lit("")) & (k >= 0))
    add("str_neq_cols", "select count(1), sum(k) from t where s_opt != s_req and k >= 0;",
        [cnt, P.call("sum", k)], where=s_opt.neq(s_req) & (k >= 0))
    add("str_eq_cols", "select count(1), min(k), max(k) from t where s_opt = s_req and k >= 0;",
        [cnt, P.call("min", k), P.call("max", k)], where=s_opt.eq(s_req) & (k >= 0))
    add("str_eq_unknown_literal", "select count(1), sum(k) from t where s_req = 'no such value' and k >= 0;",
        [cnt, P.call("sum", k)], where=s_req.eq(P.lit("no such value")) & (k >= 0))
    add("str_neq_unknown_literal", "select count(1), sum(k) from t where s_req != 'another unknown value' and k >= 0;",
        [cnt, P.call("sum", k)], where=s_req.neq(P.lit("another unknown value")) & (k >= 0))
    add("str_group_nullable", "select s_opt, count(1), sum(k) from t where k < 500 group by s_opt;",
        [s_opt, cnt, P.call("sum", k)], where=k < 500, group=[s_opt])
    add("str_group_two_keys", "select s_req, s_opt, count(1), max(k) from t where k >= 0 and s_req != 'x' group by s_req, s_opt;",
        [s_req, s_opt, cnt, P.call("max", k)], where=(k >= 0) & s_req.neq(P.lit("x")), group=[s_req, s_opt])
    add("str_group_mixed_key", "select s_req, k / 500, count(1) from t where k >= 0 group by s_req, k / 500;",
        [s_req, k / 500, cnt], where=k >= 0, group=[s_req, k / 500])
    # ordering comparisons (strncmp over the shorter length, then the lengths; a NULL compares as "") and startswith / endswith
    # between a string column and a literal: evaluated once per dictionary entry on the device path
    add("str_lt_lit", "select count(1), sum(k) from t where s_req < 'google' and k >= 0;",
        [cnt, P.call("sum", k)], where=(s_req < P.lit("google")) & (k >= 0))
    add("str_gte_lit_nullable", "select count(1), sum(k) from t where s_opt >= 'f' and k >= 0;",
        [cnt, P.call("sum", k)], where=(s_opt >= P.lit("f")) & (k >= 0))
    add("str_lte_null_is_empty", "select count(1), sum(k) from t where s_opt <= 'a' and k >= 0;",
        [cnt, P.call("sum", k)], where=(s_opt <= P.lit("a")) & (k >= 0))
    add("str_lit_gt_col", "select count(1), min(k) from t where 'newsletter' > s_req and k >= 0;",
        [cnt, P.call("min", k)], where=(P.lit("newsletter") > s_req) & (k >= 0))
    add("str_startswith", "select count(1), sum(k) from t where startswith(s_req, 'google') and k >= 0;",
        [cnt, P.call("sum", k)], where=P.call("startswith", s_req, P.lit("google")) & (k >= 0))
    add("str_endswith_nullable_group", "select k / 500, count(1) from t where endswith(s_opt, '.com') and k >= 0 group by k / 500;",
        [k / 500, cnt], where=P.call("endswith", s_opt, P.lit(".com")) & (k >= 0), group=[k / 500])
    add("str_startswith_empty_matches_all", "select count(1) from t where startswith(s_opt, '') and k >= 0;",
        [cnt], where=P.call("startswith", s_opt, P.lit("")) & (k >= 0))
    add("str_scan_projection", "select s_opt, k, s_req from t where s_req = 'facebook' or s_opt = 'x';",
        [s_opt, k, s_req], where=s_req.eq(P.lit("facebook")) | s_opt.eq(P.lit("x")), flags=0, ordered=True)
    return Q


def parse_string_query_rows(rows, types):
    """Golden rows of ref_strings.json 'queries' -> tuples: numbers as int, strings as their digest text, NULL as None."""
    out = []
    for r in rows:
        vals = []
        for s, t in zip(r, types):
            if s == "NULL":
                vals.append(None)
            elif t == "string":
                vals.append(s)
            elif t == "float64":
                vals.append(float(s))
            elif t == "bool":
                vals.append(s == "true")
            else:
                vals.append(int(s))
        out.append(tuple(vals))
    return out


# ---- LSM partitions scanned by the reference's own PartitionCursor (tests/golden/ref_lsm.json) ----
LSM_SIZES = [900, 1500, 700]
LSM_KEY_SPACE = 300
# name -> per on-disk table, newest first: (has_skiplist, has_updates); the last one is the partition's oldest table
LSM_CASES = {
    "skiplist_then_updates": [(True, True), (False, True), (False, False)],
    "first_unfiltered": [(False, False), (False, True), (False, True)],
    "nothing_filtered": [(False, False), (False, False), (False, True)],
    "all_skiplists": [(True, False), (True, False), (True, False)],
    "updates_everywhere": [(False, True), (False, True), (False, True)],
}


def lsm_case_segments(case: str):
    """[(use_skip_column, has_updates, oldest)] of LSM_CASES[case], the arguments of O.LsmSegment / Context.lsm_build_filters."""
    meta = LSM_CASES[case]
    return [(m[0], m[1], i == len(meta) - 1) for i, m in enumerate(meta)]
