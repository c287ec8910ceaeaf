#!/usr/bin/env python3
"""Throughput of a TPC-H-Q1-style GROUP BY whose keys are STRING columns (returnflag 'A'/'N'/'R', linestatus 'O'/'F') with
a string predicate - the dictionary-code path of csrc/strings.cu - on one B200.  A secondary measurement for profiles/, not
the bench.py headline.  The table is built from numpy streams through evqgpu_table_add_stream (numeric columns UINT64_PLAIN,
string columns STRING_PLAIN); the scan kernel time comes from CUDA events around the launches (evqgpu_ctx_set_profiling).

Usage: python scripts/strq_bench.py [rows] [reps]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventql_b200 import capi, plan as P  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    ctx = capi.Context(0)
    rng = np.random.default_rng(7)
    t = ctx.create_table(rows)
    num = {"shipdate": rng.integers(8036, 10562, rows), "quantity": rng.integers(1, 51, rows),
           "price": rng.integers(90000, 10090000, rows), "discount": rng.integers(0, 11, rows)}
    for name, v in num.items():
        t.add_column(name, P.COL_UNSIGNED_INT, P.ENC_UINT64_PLAIN)
        t.add_stream(name, P.STREAM_DATA, v.astype("<u8").view(np.uint8))
    flags = {"returnflag": np.frombuffer(b"ANR", dtype=np.uint8)[rng.integers(0, 3, rows)],
             "linestatus": np.frombuffer(b"OF", dtype=np.uint8)[rng.integers(0, 2, rows)]}
    t0 = time.perf_counter()
    for name, ch in flags.items():
        stream = np.empty((rows, 2), dtype=np.uint8)
        stream[:, 0] = 1
        stream[:, 1] = ch
        t.add_column(name, P.COL_STRING, P.ENC_STRING_PLAIN)
        t.add_stream(name, P.STREAM_DATA, stream.reshape(-1))
    ctx.synchronize()
    t_load = time.perf_counter() - t0
    names = ["shipdate", "quantity", "price", "discount", "returnflag", "linestatus"]
    c = {n: P.Col(i, P.STRING if n in flags else P.UINT64) for i, n in enumerate(names)}
    plan = P.QueryPlan(names, [c["returnflag"], c["linestatus"], P.call("count", P.lit(1)), P.call("sum", c["quantity"]),
                               P.call("sum", c["price"]), P.call("sum", c["price"] * (P.lit(100) - c["discount"]))],
                       where=(c["shipdate"] <= 10471) & c["returnflag"].neq(P.lit("X")), group=[c["returnflag"], c["linestatus"]])
    q = ctx.query(plan)
    t0 = time.perf_counter()
    q.execute([t])                      # first execution: dictionary codes of the two columns + NVRTC
    t_first = time.perf_counter() - t0
    got = sorted(q.rows())
    # check against numpy
    keep = num["shipdate"] <= 10471
    want = []
    for a in b"ANR":
        for b in b"OF":
            m = keep & (flags["returnflag"] == a) & (flags["linestatus"] == b)
            want.append((bytes([a]), bytes([b]), int(m.sum()), int(num["quantity"][m].sum()), int(num["price"][m].sum()),
                         int((num["price"][m] * (100 - num["discount"][m])).sum())))
    assert got == sorted(want), (got[:2], want[:2])
    ctx.set_profiling(True)
    ms = []
    for r in range(reps):
        q.execute([t])
        st = q.stats()
        ms.append(st["scan_ms"])
    st = q.stats()
    med = float(np.median(ms))
    out = {"workload": "Q1-style GROUP BY on two STRING keys + string predicate (dictionary codes), UINT64_PLAIN measures",
           "rows": rows, "groups": len(got), "strategy": st["strategy"], "reps": reps,
           "scan_ms_median": round(med, 4), "scan_ms_min": round(min(ms), 4),
           "rows_per_s": round(rows / med * 1e3), "algorithmic_bytes": st["algorithmic_bytes"],
           "bytes_per_row": round(st["algorithmic_bytes"] / rows, 2), "GBps": round(st["algorithmic_bytes"] / med / 1e6, 1),
           "string_columns_load_s": round(t_load, 3), "first_execute_s": round(t_first, 3),
           "result_matches_numpy": True,
           "timing": "CUDA events around the scan kernel launches (evqgpu_query_stats.scan_ms)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
