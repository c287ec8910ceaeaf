"""Where the end-to-end step (bench.py's e2e leg) spends its time: host timers around every C-ABI call of one step.

  python scripts/e2e_phases.py [rows_per_partition] [partitions] [steps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventql_b200 import capi, plan as P
from tests import common as T

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000_000
nparts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3

ctx = capi.Context(0)
spec = T.lineitem_spec()
sql, plan = T.q1(spec)
used = set()
for s in spec:
    used.add(s["name"])
host = []
h2d = 0
for p in range(nparts):
    t = ctx.synthesize(rows, spec, row_offset=p * rows)
    cols = []
    for info in t.columns():
        data, mx = t.read_stream(info["name"], P.STREAM_DATA)
        pin = ctx.host_alloc(max(1, data.nbytes))
        pin[: data.nbytes] = data
        cols.append((info, pin[: data.nbytes], mx))
        h2d += data.nbytes
    host.append((t.num_rows, cols))
    t.close()
q = ctx.query(plan)


def now():
    return time.perf_counter()


for it in range(steps + 1):
    ph = {}
    ctx.synchronize()
    t_begin = now()
    tbls = []
    for nrows_t, cols in host:
        t0 = now()
        t = ctx.create_table(nrows_t)
        ph["create"] = ph.get("create", 0) + now() - t0
        for info, pin, mx in cols:
            t0 = now()
            t.add_column(info["name"], info["logical_type"], info["encoding"], info["dlevel_max"])
            t1 = now()
            t.add_stream(info["name"], P.STREAM_DATA, pin, mx)
            t2 = now()
            ph["add_column"] = ph.get("add_column", 0) + t1 - t0
            ph["add_stream:" + info["name"]] = ph.get("add_stream:" + info["name"], 0) + t2 - t1
        tbls.append(t)
    t0 = now()
    q.execute(tbls)
    t1 = now()
    out = q.fetch_packed()
    t2 = now()
    for t in tbls:
        t.close()
    t3 = now()
    ph["execute"] = t1 - t0
    ph["fetch"] = t2 - t1
    ph["close"] = t3 - t2
    total = t3 - t_begin
    print("step %d: %.1f ms total, %.2f G rows/s, H2D %.2f GB -> %.1f GB/s if it were all copy" %
          (it, total * 1e3, rows * nparts / total / 1e9, h2d / 1e9, h2d / total / 1e9))
    for k, v in ph.items():
        print("    %-22s %8.2f ms" % (k, v * 1e3))
    st = q.stats()
    print("    stats: jit_ms=%.1f scan_launches=%d" % (st.get("jit_ms", 0), st.get("scan_launches", 0)))
