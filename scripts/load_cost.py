"""Time the load-time passes (index + statistics) per column: synthesize a 1-partition lineitem table under CUDA events."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventql_b200 import capi
from tests import common as T
ctx = capi.Context(0)
spec = T.lineitem_spec()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000_000
for s in spec:
    ctx.synchronize(); t0 = time.perf_counter()
    t = ctx.synthesize(n, [s])
    ctx.synchronize(); dt = time.perf_counter() - t0
    info = t.columns()[0]
    print("%-10s %.1f ms  bytes=%d leb_len=%d vmax_bits=%d" % (s["name"], dt * 1e3, info["data_bytes"], info["leb_max_len"], info["value_bits"]))
    t.close()
