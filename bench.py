#!/usr/bin/env python3
"""bench.py - rows/s and GB/s of scan + filter + GROUP BY (BASELINE.json) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3_q1|c2_q6|c4_highcard|c5_timeseries]
  python bench.py --impl reference ...      the reference's own CPU engine (oracle/_ref/evqlref) on the host cores

One "step" = one pass of the query over every partition resident on this rank + the cross-GPU merge of the partial
aggregates.  Default workload: C3 of SURVEY.md 8(d) - the TPC-H-Q1-style query over 125 M-row lineitem partitions.
  N = 1   `value`: 8 partitions = 1 B rows on the one GPU (BASELINE.json configs[2]).
  N > 1   `value`: the SAME 1 B rows split over the N GPUs (8 / N partitions each) - strong scaling, BASELINE's own split;
          the weak number (8 partitions = 1 B rows PER GPU) is reported next to it as `c3_weak`.  Before anything is timed
          the ranks run a rank-spanning parity check of both merge strategies against the reference engine
          (`parity_multi_rank`).
The default line also carries `configs`: the nullable twin of C3 and BASELINE's C2 / C4 / C5, a few steps each, with their
own roofline records; `cpu_baseline` (the reference engine on one host core, at every N) and `e2e` (host buffers).

Prints ONE JSON line (rank 0).  Timing: CUDA events on the library's stream, max over ranks; inputs (>= 1.3 GB per GPU
and step) are far larger than the 126 MB L2, so no explicit flush is needed between steps.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EVQLREF = os.path.join(ROOT, "oracle", "_ref", "evqlref")
_JSON_OUT = sys.stdout


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="evq", choices=["evq", "reference"])
    ap.add_argument("--workload", default="c3_q1", choices=["c3_q1", "c3_q1_plain", "c3_q1_null", "c2_q6", "c4_highcard", "c5_timeseries"])
    ap.add_argument("--rows-per-partition", type=int, default=0, help="0 = the workload's default")
    ap.add_argument("--partitions-per-gpu", type=int, default=0, help="0 = the workload's default")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-partitions", type=int, default=0, help="0 = 4 partitions per GPU and step (2 beyond 2 GPUs: pinned host memory)")
    ap.add_argument("--config-steps", type=int, default=5, help="timed steps of every sub-record under `configs`")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--cpu-sample-rows", type=int, default=8_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ref-rows-per-core", type=int, default=8_000_000)
    return ap.parse_args()


# ---- workloads ------------------------------------------------------------------------------------------

def workload(name):
    """-> dict(spec(partition index) -> column specs, query(spec) -> (sql, plan), table alias, default partition layout)"""
    from tests import common as T   # table / query definitions shared with the parity tests (no oracle import)
    if name == "c3_q1":
        return dict(spec=lambda p: T.lineitem_spec(), query=lambda s: T.q1(s), alias="lineitem", rows=125_000_000, parts=8,
                    desc="C3: TPC-H-Q1-style scan+filter+GROUP BY (4 groups, 8 aggregates + 3 means), lineitem UINT64_LEB128")
    if name == "c3_q1_plain":
        from eventql_b200 import plan as P
        return dict(spec=lambda p: T.lineitem_spec(P.ENC_UINT64_PLAIN), query=lambda s: T.q1(s), alias="lineitem", rows=125_000_000, parts=2,
                    desc="C3 (PLAIN twin): the same Q1 over lineitem stored UINT64_PLAIN (56 B/row)")
    if name == "c3_q1_null":
        return dict(spec=lambda p: T.lineitem_spec(null_every=7), query=lambda s: T.q1(s), alias="lineitem", rows=125_000_000, parts=1,
                    desc="C3 (nullable twin, SURVEY 8d): Q1 over lineitem with optional price / tax / flag (NULL every 7th row): the layout "
                         "columns not declared NOT NULL have (TableSchema.cc:252-322)")
    if name == "c2_q6":
        return dict(spec=lambda p: T.lineitem_spec(), query=lambda s: T.q6(s), alias="lineitem", rows=100_000_000, parts=1,
                    desc="C2: TPC-H-Q6-style selective filter + global SUM, 100 M-row lineitem UINT64_LEB128")
    if name == "c4_highcard":
        return dict(spec=lambda p: T.events_spec(10_000_000), query=lambda s: T.q_highcard(s, expected_groups=10_000_000),
                    alias="events", rows=125_000_000, parts=8,
                    desc="C4: high-cardinality GROUP BY (10 M distinct u64 keys) count/sum/mean")
    if name == "c5_timeseries":
        return dict(spec=lambda p: T.readings_spec(p), query=lambda s: T.q_timeseries(s, expected_groups=1_440_000),
                    alias="readings", rows=125_000_000, parts=1,
                    desc="C5: 1-minute bucket x 1 K sensors GROUP BY, one time partition per GPU")
    raise SystemExit("unknown workload " + name)


# ---- clocks -----------------------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for f in self.samples:
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- the reference's CPU engine on host cores --------------------------------------------------------------------------------

def run_evqlref(alias, path, sql, reps=1):
    """-> (best ms over reps, header, rows) running the unmodified reference engine on one cstable file"""
    r = subprocess.run([EVQLREF, "sql", "-t", "%s=%s" % (alias, path), "-n", str(reps), "-q", sql],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0 or "ERROR!" in r.stdout:
        raise RuntimeError("evqlref failed: %s %s" % (r.stdout[-500:], r.stderr[-500:]))
    ms = [float(l.split("ms=")[1].split()[0]) for l in r.stderr.splitlines() if l.startswith("TIMING")]
    lines = [l for l in r.stdout.split("\n") if l]
    return min(ms), lines[0], [l.split(";") for l in lines[1:]]


def _write_partition(args):
    path, wl_name, part, rows = args
    from tests import common as T
    wl = workload(wl_name)
    T.write_table(path, wl["spec"](part), rows, row_offset=part * rows)
    return path


def reference_arm(args):
    """bench.py --impl reference: the reference's own CPU implementation of the path (FastCSTableScan +
    GroupByExpression through its planner: oracle/_ref/evqlref, built from the unmodified sources) on all host cores:
    one process per partition file, as the reference has no intra-query threading."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.workload)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    rows = args.ref_rows_per_core
    spec0 = wl["spec"](0)
    sql, _plan = wl["query"](spec0)
    kind = "reference"
    if not os.path.exists(EVQLREF):
        kind = "port"
    tmp = tempfile.mkdtemp(prefix="evqref")
    import multiprocessing as mp
    with mp.Pool(min(cores, 64)) as pool:
        files = pool.map(_write_partition, [(os.path.join(tmp, "p%d.cst" % i), args.workload, i, rows) for i in range(cores)])

    def step():
        t0 = time.perf_counter()
        if kind == "reference":
            procs = [subprocess.Popen([EVQLREF, "sql", "-t", "%s=%s" % (wl["alias"], f), "-q", sql], stdout=subprocess.DEVNULL,
                                      stderr=subprocess.DEVNULL) for f in files]
            rcs = [p.wait() for p in procs]
            if any(rcs):
                raise RuntimeError("evqlref exited with %r" % rcs)
        else:
            from oracle import evq_oracle as O
            for f in files:
                O.run_query([O.read_cstable(f)], _plan)
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    total = sum(times)
    nrows = rows * len(files) * (1 if kind == "reference" else 1)
    value = nrows * args.steps / total
    used = cores if kind == "reference" else 1
    sample = "%d partition files x %d rows (%s), one reference process per host core" % (len(files), rows, args.workload)
    out = {
        "impl": "reference", "metric": "scan+filter+GROUP BY throughput", "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": wl["desc"], "sql": sql, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _JSON_OUT.write(json.dumps(out) + "\n")
    _JSON_OUT.flush()
    for f in files:
        try:
            os.unlink(f)
        except OSError:
            pass


def cpu_baseline(ctx, wl, args, q_rows_check=True):
    """The reference's CPU engine, one query on ONE host core (it is single-threaded per query), on a bounded sample of
    the same workload written as a cstable file by the device-side writer; its rows are also compared with the CUDA
    result on the same file (a parity check at bench time)."""
    from tests import common as T
    spec = wl["spec"](0)
    sql, plan = wl["query"](spec)
    rows = args.cpu_sample_rows
    tmp = tempfile.mkdtemp(prefix="evqcpu")
    path = os.path.join(tmp, "sample.cst")
    tbl = ctx.synthesize(rows, spec)
    tbl.write_file(path)
    q = ctx.query(plan)
    q.execute([tbl])
    gpu_rows = q.rows()
    q.close()
    tbl.close()
    sample = "%d rows of %s, 1 query, best of 2 runs" % (rows, args.workload)
    parity = None
    if os.path.exists(EVQLREF):
        kind = "reference"
        ms, hdr, raw = run_evqlref(wl["alias"], path, sql, reps=2)
        types = [h.rsplit(":", 1)[1] for h in hdr[1:].split(";")]
        ref_rows = T.parse_ref_rows(raw, types)
        parity, why = T.rows_equal(gpu_rows, ref_rows)
        if not parity:
            raise RuntimeError("bench: CUDA result differs from the reference engine on the CPU sample: " + why)
    else:
        kind = "port"
        from oracle import evq_oracle as O
        f = O.read_cstable(path)
        t0 = time.perf_counter()
        res = O.run_query([f], plan)
        ms = 1000.0 * (time.perf_counter() - t0)
        parity, why = T.rows_equal(gpu_rows, res.rows())
        if not parity:
            raise RuntimeError("bench: CUDA result differs from the oracle on the CPU sample: " + why)
    os.unlink(path)
    return {"value": rows / (ms / 1000.0), "unit": "rows/s", "cores": 1, "kind": kind, "sample": sample,
            "parity_with_cuda_on_sample": bool(parity), "host_cpus": os.cpu_count()}


# ---- our arm ------------------------------------------------------------------------------------------------------------

def referenced_columns(plan):
    """Names of the input columns the plan's expressions actually read."""
    from eventql_b200 import plan as P
    seen = set()

    def walk(e):
        if e is None:
            return
        if isinstance(e, P.Col):
            seen.add(e.index)
        elif isinstance(e, P.Call):
            for a in e.args:
                walk(a)
        elif isinstance(e, P.If):
            walk(e.cond), walk(e.then), walk(e.otherwise)

    walk(plan.where)
    for e in list(plan.group) + list(plan.select):
        walk(e)
    return [n for i, n in enumerate(plan.input_columns) if i in seen]


STRATEGY = {0: "scan-only", 1: "dense (registers/shared memory)", 2: "global hash table", 3: "direct-addressed group array (L2 atomics)",
            4: "global hash table, partitioned aggregation (records by home slot over two levels, table slices aggregated in shared memory)"}


class Job:
    """One rank of the bench: context, communicator, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from eventql_b200 import capi
        self.torch, self.dist = torch, dist
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run --nproc-per-node %d)" % (args.gpus, self.world, args.gpus))
        torch.cuda.set_device(self.local)
        # run (and allocate the pinned host buffers of the e2e leg) on the CPUs next to this GPU: the H2D copies of the
        # encoded streams otherwise cross the socket interconnect
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(self.local))
        except Exception:
            pass
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.ctx = capi.Context(self.local)          # raises without a device: there is no CPU fallback
        if self.world > 1:
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if self.rank == 0:
                idt = torch.frombuffer(bytearray(capi.Context.comm_unique_id()), dtype=torch.uint8).cuda()
            dist.broadcast(idt, 0)
            self.ctx.comm_init(bytes(idt.cpu().numpy().tobytes()), self.rank, self.world)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            self.peak, self.peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            self.peak, self.peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    def barrier(self):
        self.ctx.synchronize()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.int64, device="cuda")
        self.dist.all_reduce(t)
        return int(t.item())

    def timed(self, q, tbls, k):
        """k steps (scan of every resident partition + merge), CUDA events on the library's stream, barrier + synchronize
        on both sides; -> (ms max over ranks, stats of the last step)"""
        torch = self.torch
        ext = torch.cuda.ExternalStream(self.ctx.stream, device=torch.device("cuda", self.local))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(ext)
        for _ in range(k):
            q.enqueue(tbls)
            if self.world > 1:
                q.merge()
        e1.record(ext)
        self.barrier()
        ms = e0.elapsed_time(e1)
        q.finish()
        return self.max_over_ranks(ms), q.stats()

    def close(self):
        self.ctx.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def run_workload(job, name, steps, warmup, rows=0, parts=0, keep_tables=False, sampler=None):
    """Synthesize the workload's partitions on this rank, run W warm-up + K timed steps, return the record (and the live
    tables / query when keep_tables)."""
    from eventql_b200 import plan as P
    ctx, rank, world = job.ctx, job.rank, job.world
    wl = workload(name)
    rows = rows or wl["rows"]
    parts = parts or wl["parts"]
    tables = []
    for p in range(parts):
        gp = rank * parts + p                      # global partition index: every rank holds different rows
        tables.append(ctx.synthesize(rows, wl["spec"](gp), row_offset=gp * rows))
    sql, plan = wl["query"](wl["spec"](0))
    if world > 1:
        plan.flags |= P.QUERY_PARTIAL
    q = ctx.query(plan)
    # the first execution: NVRTC of the specialised kernel (cold unless the on-disk cubin cache has it), the key-bounds
    # pre-pass, and - at N > 1 - the collective that agrees on slot assignment and state layout (evqgpu_query_prepare)
    t0 = time.perf_counter()
    q.execute(tables)
    if world > 1:
        q.merge()
    first = q.stats()
    first_ms = 1000.0 * (time.perf_counter() - t0)
    for _ in range(max(warmup, 1)):
        q.enqueue(tables)
        if world > 1:
            q.merge()
    q.finish()

    # the timed region: K steps, nothing between the kernels (no per-launch events)
    launches0 = ctx.kernel_launches
    ms, stats = job.timed(q, tables, steps)
    launches = ctx.kernel_launches - launches0
    # a second pass of K steps with a CUDA event pair around every scan launch: the kernel's mean launch duration (roofline)
    ctx.set_profiling(True)
    _pms, pstats = job.timed(q, tables, steps)
    ctx.set_profiling(False)
    stats["scan_ms"], stats["scan_launches"] = pstats["scan_ms"], pstats["scan_launches"]
    if sampler is not None and ms < 400:
        # keep the GPU busy a little longer when the region is shorter than the sampler period, so the clocks are seen under load
        job.timed(q, tables, int(min(200, max(1, 400 / max(ms / steps, 0.01)))))
    result_rows = q.rows()
    rows_rank = rows * parts
    total_rows = rows_rank * world
    algo = stats["algorithmic_bytes"]
    nscan = max(1, stats["scan_launches"])
    scan_ms = stats["scan_ms"] / nscan
    per_launch = algo / parts
    achieved = per_launch / (scan_ms / 1000.0) / 1e9 if scan_ms > 0 else None
    passed = job.sum_over_ranks(int(stats["rows_passed"]))
    groups = len(result_rows) if (world == 1 or stats["strategy"] in (1, 3)) else job.sum_over_ranks(len(result_rows))
    rec = {
        "workload": wl["desc"], "sql": sql, "value": total_rows * steps / (ms / 1000.0), "unit": "rows/s", "scaling": "weak",
        "ms_per_step": ms / steps, "steps": steps, "rows_per_gpu": rows_rank, "partitions_per_gpu": parts, "rows_per_partition": rows,
        "total_rows": total_rows, "rows_passed": passed, "groups": groups, "strategy": STRATEGY[stats["strategy"]],
        "gbs": algo * world * steps / (ms / 1000.0) / 1e9,
        "first_execute_ms": first_ms, "jit_ms_first_query": first["jit_ms"], "jit_from_disk_cache": bool(first["jit_disk_hits"]),
        "merge": "none (1 GPU)" if world == 1 else ("NCCL over NVLink inside every step: " +
                                                    ("dense state over peer-mapped NVLink buffers, merged + emitted by the tail kernel" if stats["strategy"] == 1 else
                                                     "ncclAllReduce(sum) of the direct-addressed group array" if stats["strategy"] == 3 else
                                                     "hash repartition (all-to-all) + owner-side insert")),
        "roofline": {"bound": "hbm",
                     # (strategy 4: the event pair of a table's scan also brackets the passes behind it)
                     "kernel": "evq_scan + evq_repart + evq_agg_smem (per table)" if stats["strategy"] == 4 else "evq_scan",
                     "achieved": achieved, "peak": job.peak, "unit": "GB/s",
                     "frac": (achieved / job.peak) if achieved else None, "traffic": None,
                     "traffic_note": "not measured in this run (needs ncu); per-kernel dram__bytes are in profiles/*ncu*.txt",
                     "peak_source": job.peak_src, "algorithmic_bytes_per_launch": per_launch, "launch_ms": scan_ms,
                     "launches_timed": nscan, "launch_timing": "CUDA event pair around every evq_scan launch, a second pass of K steps right after the timed region",
                     "bytes_per_row": algo / rows_rank,
                     "step_frac_of_aggregate_peak": algo * steps / (ms / 1000.0) / 1e9 / job.peak},
        "gpu_launches": int(launches),
    }
    # correctness guard inside the bench: count(1) over all groups == rows that passed WHERE on all ranks
    if name.startswith("c3_q1") and world == 1:
        cnt = sum(r[2] for r in result_rows)
        if cnt != passed:
            raise RuntimeError("bench: count(1) over all groups is %d, rows passed %d of %d" % (cnt, passed, total_rows))
    if keep_tables:
        return rec, tables, q, plan, wl
    q.close()
    for t in tables:
        t.close()
    return rec


def parity_multi_rank(job):
    """N > 1, before anything is timed: every rank synthesises a DIFFERENT slice of a small table, the ranks execute the
    partial plan and merge over NCCL, and the merged rows must equal what the reference's CPU engine (oracle/_ref/evqlref -
    the checker, as in cpu_baseline) returns on the whole table.  Covers both merge strategies: the dense all-gather merge
    (Q1; float min / max / mean; a NULL-able key) and the hash repartition merge (50 K keys)."""
    from eventql_b200 import plan as P
    from tests import common as T
    ctx, rank, world, dist = job.ctx, job.rank, job.world, job.dist
    cases = []
    spec = T.lineitem_spec()
    sql, plan = T.q1(spec)
    cases.append(("q1_dense", "lineitem", spec, sql, plan, 400_000))
    mixed = T.mixed_spec()
    pc = {c[0]: c for c in T.partial_cases()}
    for cname, label in (("pa_two_keys_all_aggs", "float_min_max_mean_dense"), ("pa_null_key", "null_key_dense")):
        _n, sql, plan = pc[cname]
        plan.flags = P.QUERY_GROUPBY
        cases.append((label, "t", mixed, sql, plan, 150_000))
    spec = T.events_spec(50_000)
    sql, plan = T.q_highcard(spec, expected_groups=50_000)
    cases.append(("highcard_hash", "events", spec, sql, plan, 300_000))
    done = []
    for label, alias, spec, sql, plan, n in cases:
        plan.flags |= P.QUERY_PARTIAL
        tbl = ctx.synthesize(n, spec, row_offset=rank * n)
        q = ctx.query(plan)
        q.execute([tbl])          # collective: prepare (slot assignment, layout) + scan
        q.merge()                 # collective: NCCL
        rows = q.rows()
        strategy = q.stats()["strategy"]
        q.close()
        tbl.close()
        gathered = [None] * world
        dist.all_gather_object(gathered, (strategy, rows))
        ok, why = True, ""
        if rank == 0:
            strategies = {g[0] for g in gathered}
            if strategy == 1:     # dense: every rank ends with the full result
                merged = gathered[0][1]
                for g in gathered[1:]:
                    e, w = T.rows_equal(g[1], merged)
                    ok, why = (ok and e), (why or w)
            else:                 # hash: every rank ends with its share of the groups
                merged = [r for g in gathered for r in g[1]]
            if len(strategies) != 1:
                ok, why = False, "ranks chose different strategies %r" % (strategies,)
            if not os.path.exists(EVQLREF):
                raise RuntimeError("bench: oracle/_ref/evqlref is missing - the multi-rank parity check needs the reference engine")
            tmp = tempfile.mkdtemp(prefix="evqpar")
            path = os.path.join(tmp, "whole.cst")
            whole = ctx.synthesize(n * world, spec, row_offset=0)   # the concatenation of the ranks' slices
            whole.write_file(path)
            whole.close()
            _ms, hdr, raw = run_evqlref(alias, path, sql)
            types = [h.rsplit(":", 1)[1] for h in hdr[1:].split(";")]
            e, w = T.rows_equal(merged, T.parse_ref_rows(raw, types))
            ok, why = (ok and e), (why or w)
            os.unlink(path)
            done.append({"case": label, "strategy": STRATEGY[strategy], "rows_per_rank": n, "groups": len(merged), "equal_to_reference_engine": bool(ok)})
        flag = [ok, why]
        dist.broadcast_object_list(flag, 0)
        if not flag[0]:
            raise RuntimeError("bench: multi-rank merge differs from the reference engine in case %s: %s" % (label, flag[1]))
    return done


def e2e_leg(job, tables, plan, steps, nparts):
    """End to end through the C ABI with HOST buffers: per step, for every partition, H2D of the encoded column streams from
    pinned host memory (evqgpu_table_create / add_stream: copy, row-tile index, statistics), the scan, the merge, D2H of the
    result rows - all inside the timed region (host clock, barrier on both sides, max over ranks)."""
    from eventql_b200 import plan as P
    ctx, world = job.ctx, job.world
    nparts = min(nparts, len(tables))
    host = []
    used = referenced_columns(plan)
    h2d = 0
    for t in tables[:nparts]:
        cols = []
        for info in t.columns():
            if info["name"] not in used:
                continue
            data, mx = t.read_stream(info["name"], P.STREAM_DATA)
            pin = ctx.host_alloc(max(1, data.nbytes))
            pin[: data.nbytes] = data
            cols.append((info, pin[: data.nbytes], mx))
            h2d += data.nbytes
        host.append((t.num_rows, cols))
    q2 = ctx.query(plan)

    def step():
        tbls = []
        for nrows_t, cols in host:
            t = ctx.create_table(nrows_t)
            for info, pin, mx in cols:
                t.add_column(info["name"], info["logical_type"], info["encoding"], info["dlevel_max"])
                t.add_stream(info["name"], P.STREAM_DATA, pin, mx)
            tbls.append(t)
        q2.execute(tbls)
        if world > 1:
            q2.merge()
        out = q2.fetch_packed()
        for t in tbls:
            t.close()
        return sum(len(c) for c in out)

    step()
    job.barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(steps):
        d2h = step()
    job.barrier()
    dt = job.max_over_ranks(time.perf_counter() - t0)
    q2.close()
    e2e_rows = sum(h[0] for h in host) * world
    return {"value": e2e_rows * steps / dt, "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h + 56,
            "rows_per_step": e2e_rows, "partitions_per_gpu_per_step": nparts, "steps": steps, "ms_per_step": 1000.0 * dt / steps,
            "h2d_gbs_per_gpu": h2d * steps / dt / 1e9,
            "bound": "PCIe: the step moves h2d_bytes_per_step over the host link; the scan itself is ~1 % of it",
            "path": "evqgpu_table_create/add_stream (pinned host -> HBM) + evqgpu_query_execute + merge + evqgpu_query_fetch"}


def evq_arm(args):
    job = Job(args)
    ctx, rank, world = job.ctx, job.rank, job.world
    name = args.workload
    parity = parity_multi_rank(job) if world > 1 else None

    sampler = ClockSampler(job.local)
    sampler.start()
    time.sleep(0.3)
    rec, tables, q, plan, wl = run_workload(job, name, args.steps, args.warmup, args.rows_per_partition, args.partitions_per_gpu,
                                            keep_tables=True, sampler=sampler)
    sampler.stop()
    rows, parts = rec["rows_per_partition"], rec["partitions_per_gpu"]

    # the exact BASELINE configs[2] split: 1 B rows over the N GPUs (8 / N partitions per GPU) - strong scaling
    strong = None
    if name == "c3_q1" and parts * world >= 8 and 8 % world == 0 and parts >= 8 // world:
        sub = tables[: 8 // world]
        q.prepare(sub)            # collective at N > 1: the table set changed
        for _ in range(3):
            q.enqueue(sub)
            if world > 1:
                q.merge()
        q.finish()
        sms, sst = job.timed(q, sub, args.steps)     # (no per-launch events here: they would sit between the kernels)
        strong = {"rows": rows * 8, "partitions_per_gpu": 8 // world, "ms_per_step": sms / args.steps, "scaling": "strong",
                  "value": rows * 8 * args.steps / (sms / 1000.0), "unit": "rows/s",
                  "frac_of_aggregate_peak": sst["algorithmic_bytes"] * args.steps / (sms / 1000.0) / 1e9 / job.peak,
                  # what a step costs beyond its scan launches (launch gaps, the merge + emit tail): step - partitions x the
                  # mean scan launch time measured in the main region
                  "fixed_ms_per_step": sms / args.steps - rec["roofline"]["launch_ms"] * (8 // world)}

    e2e = None
    if not args.no_e2e:
        nparts = args.e2e_partitions or (4 if world <= 2 else 2)
        e2e = e2e_leg(job, tables, plan, args.e2e_steps, nparts)
    q.close()
    for t in tables:
        t.close()

    # the other BASELINE configs and the nullable twin of C3, a few steps each (the default line; a named --workload runs alone)
    configs = None
    if name == "c3_q1" and not args.no_configs:
        configs = {}
        for sub_name, kw in (("c3_q1_null", dict(rows=125_000_000, parts=1)), ("c2_q6", {}), ("c4_highcard", {}), ("c5_timeseries", {})):
            configs[sub_name] = run_workload(job, sub_name, args.config_steps, 3, **kw)

    cpu = None
    if not args.no_cpu_baseline:
        if rank == 0:
            cpu = cpu_baseline(ctx, wl, args)
        if world > 1:
            job.dist.barrier()

    clocks = sampler.summary()
    if rank == 0:
        main_value, scaling = rec["value"], "weak"
        ms_per_step = rec["ms_per_step"]
        if strong and world > 1:
            # at N > 1 the headline is BASELINE's own split: 1 B rows over the N GPUs; the weak number stays as c3_weak
            main_value, scaling, ms_per_step = strong["value"], "strong", strong["ms_per_step"]
        out = {
            "metric": "scan+filter+GROUP BY throughput", "value": main_value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": rec["workload"], "sql": rec["sql"],
                       "rows_per_gpu": (strong["rows"] // world) if scaling == "strong" else rec["rows_per_gpu"],
                       "partitions_per_gpu": strong["partitions_per_gpu"] if scaling == "strong" else parts, "rows_per_partition": rows,
                       "total_rows": strong["rows"] if scaling == "strong" else rec["total_rows"], "groups": rec["groups"],
                       "strategy": rec["strategy"],
                       "l2": "inputs (%.1f GB per GPU) are larger than L2; no flush" % (rec["roofline"]["algorithmic_bytes_per_launch"] * (strong["partitions_per_gpu"] if scaling == "strong" else parts) / 1e9),
                       "merge": rec["merge"], "jit_ms_first_query": rec["jit_ms_first_query"], "first_execute_ms": rec["first_execute_ms"]},
            "gbs": rec["gbs"] if scaling == "weak" else strong["value"] * rec["roofline"]["bytes_per_row"] / 1e9,
            "roofline": rec["roofline"], "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": rec["gpu_launches"], "clocks": clocks,
        }
        if strong:
            out["c3_strong"] = strong
        out["c3_weak"] = {k: rec[k] for k in ("value", "unit", "ms_per_step", "rows_per_gpu", "total_rows", "gbs")}
        if parity is not None:
            out["parity_multi_rank"] = all(c["equal_to_reference_engine"] for c in parity)
            out["parity_multi_rank_cases"] = parity
        if configs:
            out["configs"] = configs
        _JSON_OUT.write(json.dumps(out) + "\n")
        _JSON_OUT.flush()
    job.close()


def main():
    args = parse_args()
    # libraries (NCCL's version banner, torchrun's notices) write to stdout: keep fd 1 for the ONE JSON line
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr",
               "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"), os.path.abspath(__file__)] + sys.argv[1:]
        os.dup2(_JSON_OUT.fileno(), 1)   # the ranks inherit the real stdout
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        reference_arm(args)
    else:
        evq_arm(args)


if __name__ == "__main__":
    main()
