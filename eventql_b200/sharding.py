"""Partition -> GPU assignment for a one-process-per-GPU job (host-side logic, no device code).

The reference shards a query by table partition and ships one partial GROUP BY per partition to the host that owns it
(server/sql/scheduler.cc:117-162 buildPipelineGroupByExpression, :164-264 pipelineExpression); here the owner of a
partition is a GPU of the box.  Partitions are the independent units: no data-path collective is needed until the one
exchange step, the merge of the partial aggregates (evqgpu_query_merge, csrc/merge.cu).
"""
from typing import List, Sequence


def assign_partitions(num_partitions: int, rank: int, world: int) -> List[int]:
    """Contiguous blocks of partitions per rank (time-partitioned tables keep neighbouring partitions together, like
    the key-range order of pipelineExpression); the first `num_partitions % world` ranks take one more."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank %d of %d" % (rank, world))
    base, extra = divmod(num_partitions, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def owner_of(partition: int, num_partitions: int, world: int) -> int:
    for r in range(world):
        if partition in assign_partitions(num_partitions, r, world):
            return r
    raise ValueError(partition)


def merge_partial_rows(rows_per_rank: Sequence[Sequence[tuple]], num_keys: int, ops: Sequence[str]) -> List[tuple]:
    """Host-side statement of GroupByMergeExpression's semantics (sql/statements/select/groupby.cc:577-612) over
    fetched result rows: rows are (key..., aggregate...); `ops[i]` in {"sum", "min", "max"} says how aggregate i merges
    (count and sum merge by +, aggregate.cc:48-50,196-198).  Used by the multi-process tests as the statement of what
    the NCCL merge must produce; the device path never calls it."""
    acc = {}
    for rows in rows_per_rank:
        for r in rows:
            k = tuple(r[:num_keys])
            v = list(r[num_keys:])
            if k not in acc:
                acc[k] = v
                continue
            cur = acc[k]
            for i, op in enumerate(ops):
                if op == "sum":
                    cur[i] = (cur[i] + v[i]) if not isinstance(cur[i], int) else (cur[i] + v[i]) & 0xFFFFFFFFFFFFFFFF
                elif op == "min":
                    cur[i] = min(cur[i], v[i])
                elif op == "max":
                    cur[i] = max(cur[i], v[i])
                else:
                    raise ValueError(op)
    return [k + tuple(v) for k, v in acc.items()]
