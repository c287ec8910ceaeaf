"""Multi-process host logic on CPU: world_size-2 gloo job (no GPU).

Every rank runs the ORACLE on the partitions eventql_b200.sharding assigns to it, the partial rows are exchanged with
torch.distributed (gloo) and merged with the host-side statement of GroupByMergeExpression's semantics; the result must
equal the oracle on the whole table.  This pins (a) the partition assignment and (b) the merge semantics the NCCL path
(csrc/merge.cu) implements on the device.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eventql_b200 import plan as P
from eventql_b200 import sharding
from tests import common as T


def test_assign_partitions_covers_everything_once():
    for n in (0, 1, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                part = sharding.assign_partitions(n, r, world)
                assert part == sorted(part)
                seen += part
            assert seen == list(range(n))
            sizes = [len(sharding.assign_partitions(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.owner_of(5, 8, 2) == 1 and sharding.owner_of(3, 8, 2) == 0
    with pytest.raises(ValueError):
        sharding.assign_partitions(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out):
    sys.path.insert(0, T.ROOT)
    import numpy as np
    from oracle import evq_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nparts, rows = 5, 20_000
    if case == "lsm":
        # Partitions of an evqld table: each is a set of segments whose visibility filter (PartitionCursor, partition_cursor.cc:
        # 157-194) spans all of them, so a partition - not a segment - is the unit a rank owns.  Every rank filters and aggregates
        # its partitions; the merged partials equal the result over all visible rows of all partitions.
        import tempfile
        nparts, nseg, seg_rows = 3, 3, 1200
        d = tempfile.mkdtemp(prefix="evqlsm%d_" % rank)
        key, v = P.Col(0, P.UINT64), P.Col(1, P.UINT64)
        plan = P.QueryPlan(["key", "v"], [key % 7, P.call("count", P.lit(1)), P.call("sum", v), P.call("min", v), P.call("max", v)],
                           where=v >= 0, group=[key % 7])

        def partition(p):
            files, filt = [], []
            segs = []
            for sgi in range(nseg):
                path = os.path.join(d, "p%d_s%d.cst" % (p, sgi))
                T.write_lsm_segment(path, 10 * p + sgi, seg_rows, key_space=250)
                f = O.read_cstable(path)
                files.append(f)
                segs.append(O.LsmSegment(f, None, sgi == 0, True))
            for f, keep in zip(files, O.lsm_visibility(segs)):
                filt.append(np.ones(f.num_rows, dtype=bool) if keep is None else keep)
            return files, filt

        def run(parts):
            files, filt = [], []
            for p in parts:
                a, b = partition(p)
                files += a
                filt += b
            if not files:
                return []
            return O.run_query(files, plan, row_filter=np.concatenate(filt)).rows()

        partial = run(sharding.assign_partitions(nparts, rank, world))
        gathered = [None] * world
        dist.all_gather_object(gathered, partial)
        merged = sharding.merge_partial_rows(gathered, 1, ["sum", "sum", "min", "max"])
        whole = run(list(range(nparts)))
        ok, why = T.rows_equal(merged, whole)
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put((bool(flag.item()), why, len(whole)))
        dist.destroy_process_group()
        return
    if case == "q1":
        spec = T.lineitem_spec(null_every=7)
        _sql, plan = T.q1(spec, means=False)
        nk, ops = 2, ["sum"] * 6
    else:
        spec = T.events_spec(3000)
        c, names = T.cols_of(spec)
        plan = P.QueryPlan(names, [c["ekey"], P.call("count", P.lit(1)), P.call("sum", c["v"]), P.call("min", c["v"]),
                                   P.call("max", c["v"])], where=c["v"] >= 0, group=[c["ekey"]])
        nk, ops = 1, ["sum", "sum", "min", "max"]

    def inputs(parts):
        cols = []
        for s in spec:
            vs, ns = zip(*[T.synth_values(s, rows, row_offset=p * rows) for p in parts]) if parts else ((), ())
            v = np.concatenate(vs) if parts else np.zeros(0, dtype=np.uint64)
            n = np.concatenate(ns) if parts else np.zeros(0, dtype=bool)
            cols.append(O.Vec(T.sql_type_of(s), v, n.astype(np.uint8)))
        return cols, rows * len(parts)

    mine = sharding.assign_partitions(nparts, rank, world)
    cols, n = inputs(mine)
    partial = O.run_query_on(cols, n, plan).rows()
    gathered = [None] * world
    dist.all_gather_object(gathered, partial)
    merged = sharding.merge_partial_rows(gathered, nk, ops)
    cols, n = inputs(list(range(nparts)))
    whole = O.run_query_on(cols, n, plan).rows()
    ok, why = T.rows_equal(merged, whole)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put((bool(flag.item()), why, len(whole)))
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["q1", "highcard", "lsm"])
def test_partial_aggregates_merge_to_the_whole_table_result(case):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    ok, why, n = out.get(timeout=10)
    assert ok, why
    assert n > 0
