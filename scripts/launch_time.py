"""Average evq_scan launch time of a workload (CUDA events around every launch), for knob sweeps:
  EVQGPU_NSTAGES=3 python scripts/launch_time.py [c3_q1|c2_q6|...] [partitions] [reps] [rows per partition]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from eventql_b200 import capi

wl = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3_q1")
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
if len(sys.argv) > 4:
    wl["rows"] = int(sys.argv[4])
ctx = capi.Context(0)
tables = [ctx.synthesize(wl["rows"], wl["spec"](p), row_offset=p * wl["rows"]) for p in range(parts)]
sql, plan = wl["query"](wl["spec"](0))
q = ctx.query(plan)
for _ in range(3):
    q.enqueue(tables)
try:
    q.finish()
except capi.EvqError as e:
    print("warm-up:", e)
ctx.set_profiling(True)
for _ in range(reps):
    q.enqueue(tables)
try:
    q.finish()
except capi.EvqError as e:
    print("run:", e)
st = q.stats()
ms = st["scan_ms"] / max(1, st["scan_launches"])
gbs = st["algorithmic_bytes"] / parts / (ms / 1e3) / 1e9
print("%s: launch %.4f ms  %.0f GB/s  frac %.3f  (%d launches)  env=%s" % (
    sys.argv[1] if len(sys.argv) > 1 else "c3_q1", ms, gbs, gbs / 6551.0, st["scan_launches"],
    {k: v for k, v in os.environ.items() if k.startswith("EVQGPU_")}))
