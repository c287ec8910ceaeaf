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
    return [dict(name="key", logical_type=P.COL_UNSIGNED_INT, encoding=key_encoding, seed=SEED0 + 100, lo=0, span=num_keys, transform=1),
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
    sql = "select key, count(1), sum(v), mean(v) from events where v >= 0 group by key;"
    return sql, P.QueryPlan(names, [c["key"], P.call("count", P.lit(1)), P.call("sum", c["v"]), P.call("mean", c["v"])],
                            where=c["v"] >= 0, group=[c["key"]], expected_groups=expected_groups)


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
