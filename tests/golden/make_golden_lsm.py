#!/usr/bin/env python3
"""Golden vectors for the LSM visibility filter, from the REFERENCE's own eventql::PartitionCursor.

Build container only (needs oracle/_ref/evqlref).  For every case of tests/common.py:LSM_CASES the partition's three
on-disk tables (tests/common.py:write_lsm_segment, written with the oracle's cstable writer, which is itself pinned to the
reference's reader) are scanned with `select v from t where v >= 0` through `evqlref sql -S`, i.e. the unmodified
PartitionCursor over a hand-built PartitionSnapshot: its needs_filter rule (server/sql/partition_cursor.cc:139-155), its
filter loop over __lsm_id / __lsm_is_update / __lsm_skip (:157-194) and one FastCSTableScan per table.  The rows it returns,
in order, are stored in tests/golden/ref_lsm.json; the generation asserts that the oracle's restatement
(oracle/evq_oracle.py:lsm_visibility) selects exactly those rows.

Usage: python tests/golden/make_golden_lsm.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import evq_oracle as O  # noqa: E402
from tests import common as T  # noqa: E402

EVQLREF = os.path.join(ROOT, "oracle", "_ref", "evqlref")


def main():
    tmp = tempfile.mkdtemp(prefix="evqlsm")
    cols = [T.write_lsm_segment(os.path.join(tmp, "seg%d.cst" % i), i, n, key_space=T.LSM_KEY_SPACE) for i, n in enumerate(T.LSM_SIZES)]
    out = {"generator": "tests/golden/make_golden_lsm.py", "reference": "17ai/eventql v0.5.0 eventql::PartitionCursor (oracle/_ref/evqlref sql -S)",
           "sql": "select v from t where v >= 0", "cases": {}}
    for case, meta in T.LSM_CASES.items():
        spec = ",".join("seg%d%s" % (i, (":" + ("s" if m[0] else "") + ("u" if m[1] else "")) if (m[0] or m[1]) else "") for i, m in enumerate(meta))
        r = subprocess.run([EVQLREF, "sql", "-S", tmp + ":" + spec, "-q", out["sql"]], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        lines = r.stdout.split("\n")
        assert r.returncode == 0 and "ERROR!" not in lines and lines[0].startswith("#"), (case, lines[:3])
        got = [int(x) for x in lines[1:] if x]
        segs = [O.LsmSegment(O.read_cstable(os.path.join(tmp, "seg%d.cst" % i)), None, m[0], None, m[1], m[2])
                for i, m in enumerate(T.lsm_case_segments(case))]
        vis = O.lsm_visibility(segs)
        want = []
        for i, n in enumerate(T.LSM_SIZES):
            keep = np.ones(n, dtype=bool) if vis[i] is None else vis[i]
            want += [int(x) for x in cols[i]["v"][keep]]
        assert got == want, case
        out["cases"][case] = {"spec": spec, "rows": got, "visible": [None if v is None else int(v.sum()) for v in vis]}
        print("case %-24s %5d rows, visible per table %s" % (case, len(got), out["cases"][case]["visible"]))
    with open(os.path.join(HERE, "ref_lsm.json"), "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print("wrote ref_lsm.json")


if __name__ == "__main__":
    main()
