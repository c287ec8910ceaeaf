#!/usr/bin/env python3
"""Per-CUDA-line instruction / stall-sample shares from `ncu --page source --csv --print-source cuda,sass` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
h = rows[hi]
ie = h.index("Instructions Executed"); sa = h.index("Warp Stall Sampling (All Samples)")
data = [r for r in rows[hi + 1:] if len(r) == len(h) and r[2] == "-"]     # CUDA-line rows have no SASS address
tot_i = sum(int(r[ie] or 0) for r in data); tot_s = sum(int(r[sa] or 0) for r in data)
print("warp instructions %d, samples %d" % (tot_i, tot_s))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
for r in sorted(data, key=lambda r: -int(r[ie] or 0))[:n]:
    print("%5.1f%% inst %5.1f%% samp | L%-5s %s" % (100 * int(r[ie] or 0) / tot_i, 100 * int(r[sa] or 0) / tot_s, r[0], r[1].strip()[:140]))
