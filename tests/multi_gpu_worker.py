"""Run under torchrun (one rank per GPU): partial aggregates on every rank, merged over NVLink with NCCL
(evqgpu_query_merge), compared on rank 0 with the oracle on the whole table.  Launched by tests/test_multi_gpu.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from eventql_b200 import capi, plan as P, sharding  # noqa: E402
from oracle import evq_oracle as O  # noqa: E402
from tests import common as T  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = capi.Context(local)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(capi.Context.comm_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    ctx.comm_init(idt.cpu().numpy().tobytes(), rank, world)

    nparts, rows = 2 * world + 1, 150_000
    cases = []
    spec = T.lineitem_spec()
    cases.append(("q1_dense_required", spec, T.q1(spec)[1]))
    cases.append(("q6_single_group", spec, T.q6(spec)[1]))
    spec = T.lineitem_spec(null_every=7)
    cases.append(("q1_dense_optional", spec, T.q1(spec)[1]))
    spec = T.events_spec(40_000)
    c, names = T.cols_of(spec)
    cases.append(("highcard_hash", spec, P.QueryPlan(names, [c["ekey"], P.call("count", P.lit(1)), P.call("sum", c["v"]), P.call("mean", c["v"]),
                                                            P.call("min", c["v"]), P.call("max", c["v"])], where=c["v"] >= 0, group=[c["ekey"]])))
    # the same through the partitioned form of the hash tier (two partitioning levels, table slices in shared memory, compact
    # table), forced on a small table: its table must merge like any other
    cases.append(("highcard_partitioned", spec, P.QueryPlan(names, [c["ekey"], P.call("count", P.lit(1)), P.call("sum", c["v"]), P.call("min", c["v"]),
                                                                   P.call("max", c["v"])], where=c["v"] >= 0, group=[c["ekey"]],
                                                            expected_groups=1 << 20)))
    spec = T.mixed_spec(null_every=5)
    c, names = T.cols_of(spec)
    cases.append(("minmax_float_dense", spec, P.QueryPlan(names, [c["bo"], P.call("count", P.lit(1)), P.call("sum", c["f"]), P.call("min", c["f"]),
                                                                 P.call("max", c["a"]), P.call("mean", c["big"]), P.call("sum", c["big"])],
                                                          where=c["b"] >= 0, group=[c["bo"]])))
    # first-row items (non-aggregate, not a function of the key): the pair with the smallest rank-major row ordinal wins the merge
    spec = T.mixed_spec()
    c, names = T.cols_of(spec)
    cases.append(("first_row_dense", spec, P.QueryPlan(names, [c["b"] % 3, P.call("count", P.lit(1)), c["a"], c["d"]], where=c["b"] >= 0,
                                                       group=[c["b"] % 3])))
    cases.append(("first_row_hash", spec, P.QueryPlan(names, [c["c"] % 5003, P.call("count", P.lit(1)), c["a"], c["k"]], where=c["b"] >= 0,
                                                      group=[c["c"] % 5003], expected_groups=1 << 25)))
    # count_distinct across ranks (dense tier): the ranks' value sets are united (count_distinct_uint64_merge, aggregate.cc:102-108)
    cd = lambda e: P.call("count_distinct", e)
    cases.append(("distinct_dense", spec, P.QueryPlan(names, [c["b"] % 3, cd(c["c"] % 50), cd(c["a"]), P.call("sum", c["b"]), P.call("count", P.lit(1))],
                                                      where=(c["b"] >= 0) & (c["c"] >= 0), group=[c["b"] % 3])))
    cases.append(("distinct_hash", spec, P.QueryPlan(names, [c["c"] % 5003, cd(c["b"]), cd(c["a"] % 11), P.call("count", P.lit(1))], where=c["b"] >= 0,
                                                     group=[c["c"] % 5003], expected_groups=1 << 25)))
    cases.append(("distinct_global", spec, P.QueryPlan(names, [cd(c["c"]), P.call("count", P.lit(1))], where=c["b"] >= 0)))
    spec = T.readings_spec(0)
    cases.append(("timeseries_direct_addressed", spec, T.q_timeseries(spec, expected_groups=1_440_000)[1]))
    failures = []

    # ---- collectives are never entered on a per-rank decision (evqgpu_query_prepare)
    spec = T.mixed_spec()
    c, names = T.cols_of(spec)
    # (1) enqueue without a prepare for these tables fails loudly instead of entering a collective alone
    tbl = ctx.synthesize(20_000, [s for s in spec if not (s["encoding"] == P.ENC_UINT32_BITPACKED and s.get("null_every"))], row_offset=rank * 20_000)
    q = ctx.query(P.QueryPlan(names, [c["b"] % 3, P.call("count", P.lit(1))], where=c["b"] >= 0, group=[c["b"] % 3],
                              flags=P.QUERY_GROUPBY | P.QUERY_PARTIAL))
    try:
        q.enqueue([tbl])
        failures.append("enqueue without prepare did not fail")
    except capi.EvqError as e:
        if "evqgpu_query_prepare" not in e.message:
            failures.append("enqueue without prepare: " + e.message)
    q.prepare([tbl])
    q.enqueue([tbl])
    q.merge()
    if sum(r[1] for r in q.rows()) != 20_000 * world:
        failures.append("prepare + enqueue + merge: wrong row count")
    q.close()
    tbl.close()
    # (2) a key expression that divides by zero on ONE rank only: that rank still takes part in the agreement and the
    # execution fails on EVERY rank (nobody is left waiting in a collective)
    bad = [dict(s) for s in spec if not (s["encoding"] == P.ENC_UINT32_BITPACKED and s.get("null_every"))]
    for s in bad:
        if s["name"] == "b":
            s["lo"] = 0 if rank == 1 % world else 1
    tbl = ctx.synthesize(20_000, bad, row_offset=rank * 20_000)
    q = ctx.query(P.QueryPlan(names, [P.lit(1000) / c["b"], P.call("count", P.lit(1))], where=c["d"] >= 0, group=[P.lit(1000) / c["b"]],
                              flags=P.QUERY_GROUPBY | P.QUERY_PARTIAL))
    try:
        q.execute([tbl])
        failures.append("rank %d: a failure on rank %d went unnoticed" % (rank, 1 % world))
    except capi.EvqError as e:
        if "rank %d failed" % (1 % world) not in e.message:
            failures.append("rank %d: unexpected message %r" % (rank, e.message))
    q.close()
    tbl.close()
    nf = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(nf)
    if rank == 0:
        print("collective agreement checks: %s" % ("ok" if int(nf.item()) == 0 else failures), flush=True)

    # ---- string GROUP BY keys and string predicates across ranks: the ranks' dictionaries are synchronised in prepare, so
    # equal strings have equal codes everywhere.  Every rank scans a DIFFERENT string table (different value sets, so
    # the local dictionaries differ before the synchronisation); the merged rows == the oracle over all tables.
    import tempfile
    sdir = tempfile.mkdtemp(prefix="evqstr%d_" % rank)
    sizes = [900 + 137 * r for r in range(world)]
    spath = os.path.join(sdir, "s.cst")
    T.write_strings_table(spath, sizes[rank], seed=1 + rank)
    stbl = ctx.open_table_file(spath)
    snames = ["k", "s_req", "s_opt"]
    sk, s_req, s_opt = P.Col(0, P.UINT64), P.Col(1, P.STRING), P.Col(2, P.STRING)
    cnt = P.call("count", P.lit(1))
    splans = [("string_keys_dense_or_hash", P.QueryPlan(snames, [s_opt, cnt, P.call("sum", sk)], where=(sk >= 0) & s_req.neq(P.lit("x")), group=[s_opt])),
              ("string_keys_hash", P.QueryPlan(snames, [s_req, s_opt, cnt, P.call("max", sk)], where=sk >= 0, group=[s_req, s_opt],
                                               expected_groups=1 << 25)),
              ("string_predicates", P.QueryPlan(snames, [cnt, P.call("sum", sk)], where=(s_req < P.lit("google")) & s_opt.neq(P.lit("facebook"))))]
    for sname, splan in splans:
        splan.flags |= P.QUERY_PARTIAL
        q = ctx.query(splan)
        q.execute([stbl])
        q.merge()
        part = q.rows()
        sstrat = q.stats()["strategy"]
        q.close()
        gathered = [None] * world
        dist.all_gather_object(gathered, part)
        allp = [None] * world
        dist.all_gather_object(allp, open(spath, "rb").read())
        if rank == 0:
            files = [O.parse_cstable(b) for b in allp]
            want = O.run_query(files, splan).rows()
            got_sets = gathered if sstrat in (1, 3) else [sum(gathered, [])]
            for got in got_sets:
                ok, why = T.rows_equal(got, want)
                if not ok:
                    failures.append("%s: %s" % (sname, why))
            print("case %-22s tier=%d groups=%d %s" % (sname, sstrat, len(want), "ok" if not failures else failures[-1]), flush=True)
    stbl.close()

    for name, spec, plan in cases:
        ts = name == "timeseries"
        mine = sharding.assign_partitions(nparts, rank, world)
        dev_spec = [s for s in spec if not (s["encoding"] == P.ENC_UINT32_BITPACKED and s.get("null_every"))]
        tables = [ctx.synthesize(rows, dev_spec, row_offset=p * rows) for p in mine]
        plan.flags |= P.QUERY_PARTIAL
        if name == "highcard_partitioned":
            os.environ["EVQGPU_PART_MIN_MB"], os.environ["EVQGPU_PART_SLICE_MB"] = "0", "1"
        q = ctx.query(plan)
        for _ in range(2):          # twice: the second run reuses the cached kernels and the agreed dense slot map
            q.execute(tables)
            q.merge()
            part = q.rows()
        stats = q.stats()
        if name == "highcard_partitioned":
            del os.environ["EVQGPU_PART_MIN_MB"], os.environ["EVQGPU_PART_SLICE_MB"]
            if stats["strategy"] != 4:
                failures.append("highcard_partitioned ran as strategy %d" % stats["strategy"])
        gathered = [None] * world
        dist.all_gather_object(gathered, part)
        if rank == 0:
            cols = []
            n = rows * nparts
            for s in spec:
                vs, ns = zip(*[T.synth_values(s, rows, row_offset=p * rows) for p in range(nparts)])
                v, st = np.concatenate(vs), T.sql_type_of(s)
                v = v.view(np.float64) if st == P.FLOAT64 else (v.astype(bool) if st == P.BOOL else v)
                cols.append(O.Vec(st, v, np.concatenate(ns).astype(np.uint8)))
            want = O.run_query_on(cols, n, plan).rows()
            if stats["strategy"] in (1, 3):
                got_sets = gathered            # dense tiers: every rank ends with the full result (hash tiers 2 / 4: distributed)
            else:
                got_sets = [sum(gathered, [])]  # hash tier: results stay distributed, each group on exactly one rank
            for got in got_sets:
                ok, why = T.rows_equal(got, want)
                if not ok:
                    failures.append("%s: %s" % (name, why))
            print("case %-22s tier=%d groups=%d %s" % (name, stats["strategy"], len(want), "ok" if not failures else failures[-1]), flush=True)
        q.close()
        for t in tables:
            t.close()
    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
