"""First contact with a real GPU: run a few parity cases and print diagnostics."""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eventql_b200 import capi, plan as P
from oracle import evq_oracle as O
from tests import common as T

ctx = capi.Context(0)
os.makedirs("/tmp/evq", exist_ok=True)

def run_case(name, spec, nrows, qf, **kw):
    path = "/tmp/evq/%s.cst" % name
    T.write_table(path, spec, nrows)
    sql, plan = qf(spec, **kw) if kw else qf(spec)
    f = O.read_cstable(path)
    exp = O.run_query([f], plan).rows()
    try:
        t = ctx.open_table_file(path)
        q = ctx.query(plan)
        t0 = time.time(); q.execute([t]); dt = time.time() - t0
        got = q.rows()
        ok, why = T.rows_equal(got, exp)
        print("%-28s rows=%d out=%d ok=%s %s first_exec=%.3fs stats=%s" % (name, nrows, len(got), ok, why[:200], dt, q.stats()))
        if not ok:
            print("  got :", sorted(got)[:4]); print("  want:", sorted(exp)[:4])
        return ok
    except Exception as e:
        print("%-28s FAILED: %s" % (name, e)); traceback.print_exc()
        return False

res = []
res.append(run_case("q6_plain_small", T.lineitem_spec(P.ENC_UINT64_PLAIN), 5000, T.q6))
res.append(run_case("q6_leb_small", T.lineitem_spec(), 5000, T.q6))
res.append(run_case("q1_leb_small", T.lineitem_spec(), 5000, T.q1))
res.append(run_case("q1_leb_1m", T.lineitem_spec(), 1_000_000, T.q1))
res.append(run_case("q1_null_300k", T.lineitem_spec(null_every=7), 300_000, T.q1))
res.append(run_case("hc_100k", T.events_spec(5000), 100_000, T.q_highcard))
res.append(run_case("ts_200k", T.readings_spec(0), 200_000, T.q_timeseries))
# decode parity of every encoding
spec = T.mixed_spec()
path = "/tmp/evq/mixed.cst"; T.write_table(path, spec, 70_000)
f = O.read_cstable(path); t = ctx.open_table_file(path)
for s in spec:
    try:
        d = O.decode_column(f, s["name"])
        st = d.sql_type
        vals = d.values.view(np.float64) if st == P.FLOAT64 else (d.values.astype(bool) if st == P.BOOL else d.values)
        want = O.pack_svector(O.Vec(st, vals, np.where(d.present, 0, 1).astype(np.uint8)))
        got = t.decode_column(s["name"])
        ok = got == want
        if not ok:
            w = 2 if st == P.BOOL else 9
            a = np.frombuffer(got, dtype=np.uint8).reshape(-1, w); b = np.frombuffer(want, dtype=np.uint8).reshape(-1, w)
            bad = np.flatnonzero((a != b).any(axis=1))
            print("  first bad rows", bad[:5], a[bad[:2]], b[bad[:2]])
        print("decode %-4s ok=%s" % (s["name"], ok)); res.append(ok)
    except Exception as e:
        print("decode %-4s FAILED: %s" % (s["name"], e)); res.append(False)
print("SUMMARY", sum(res), "/", len(res))
