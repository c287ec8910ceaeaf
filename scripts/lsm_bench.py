#!/usr/bin/env python3
"""Throughput of the string-column / visibility-filter kernels (csrc/strings.cu) on one B200 - a secondary measurement for
profiles/, not the bench.py headline.  Builds a synthetic partition of NSEG segments x ROWS rows with the reference's
bookkeeping columns (__lsm_id = 20 raw bytes, __lsm_is_update, __lsm_skip) from numpy streams, then times

  load     evqgpu_table_load_columns of the three columns (H2D + value index + statistics), host clock
  filters  evqgpu_lsm_build_filters over all segments (gather, radix sort, resolve, pack), host clock around the call after
           a device synchronise (the call synchronises itself before it returns)
  fetch    evqgpu_table_decode_string_column of one segment's ids (device gather + D2H), host clock

Usage: python scripts/lsm_bench.py [rows_per_segment] [segments] [reps]
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from eventql_b200 import capi, plan as P  # noqa: E402


def bitpack1(v):
    """1-bit libsimdcomp vertical blocks: 128 values -> 4 words, value i in word i & 3 at bit i >> 2."""
    n = (len(v) + 127) // 128 * 128
    b = np.zeros(n, dtype=np.uint32)
    b[:len(v)] = v
    b = b.reshape(-1, 32, 4)
    w = (b << np.arange(32, dtype=np.uint32)[None, :, None]).sum(axis=1, dtype=np.uint32)
    return w.reshape(-1).view(np.uint8)


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
    nseg = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    ctx = capi.Context(0)
    rng = np.random.default_rng(1)
    key_space = rows * nseg // 2
    pool = rng.integers(0, 256, size=(key_space, 20), dtype=np.uint8)
    tables = []
    t_load = 0.0
    for s in range(nseg):
        key = rng.integers(0, key_space, size=rows)
        stream = np.empty((rows, 21), dtype=np.uint8)
        stream[:, 0] = 20
        stream[:, 1:] = pool[key]
        upd = (rng.random(rows) < 0.3).astype(np.uint64)
        skip = (rng.random(rows) < 0.05).astype(np.uint64)
        t = ctx.create_table(rows) if hasattr(ctx, "create_table") else None
        if t is None:
            raise SystemExit("Context.create_table missing")
        t.add_column("__lsm_id", P.COL_STRING, P.ENC_STRING_PLAIN)
        t.add_column("__lsm_is_update", P.COL_BOOLEAN, P.ENC_BOOLEAN_BITPACKED)
        t.add_column("__lsm_skip", P.COL_BOOLEAN, P.ENC_BOOLEAN_BITPACKED)
        ub, sb = bitpack1(upd), bitpack1(skip)
        ctx.synchronize()
        t0 = time.perf_counter()
        t.add_stream("__lsm_id", P.STREAM_DATA, stream.reshape(-1))
        t.add_stream("__lsm_is_update", P.STREAM_DATA, ub, bitpack_max=1)
        t.add_stream("__lsm_skip", P.STREAM_DATA, sb, bitpack_max=1)
        ctx.synchronize()
        t_load += time.perf_counter() - t0
        tables.append(t)
    segs = [(t, None, True, True) for t in tables]
    times = []
    for r in range(reps + 1):
        ctx.synchronize()
        t0 = time.perf_counter()
        visible = ctx.lsm_build_filters(segs)
        ctx.synchronize()
        if r:
            times.append(time.perf_counter() - t0)
    total = rows * nseg
    tf = []
    for r in range(3):
        t0 = time.perf_counter()
        buf = tables[0].decode_string_column("__lsm_id")
        tf.append(time.perf_counter() - t0)
    assert len(buf) == rows * 25
    out = {"workload": "lsm_visibility", "segments": nseg, "rows_per_segment": rows, "reps": reps,
           "visible_rows": int(sum(visible)), "total_rows": total,
           "load_s": round(t_load, 4), "load_GBps": round(total * 21.25 / t_load / 1e9, 2),
           "filters_ms_median": round(float(np.median(times)) * 1e3, 3), "filters_ms_min": round(min(times) * 1e3, 3),
           "filters_Grows_per_s": round(total / float(np.median(times)) / 1e9, 3),
           "fetch_string_ms_min": round(min(tf) * 1e3, 3), "fetch_string_GBps_out": round(rows * 25 / min(tf) / 1e9, 2),
           "timing": "host clock around the C-ABI calls (each synchronises before returning); secondary measurement"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
