#!/usr/bin/env python3
"""Per-function instruction / stall-sample shares from `ncu --page source --csv --print-source cuda,sass` output."""
import csv, re, sys, collections
txt = open(sys.argv[1]).read()
rows = list(csv.reader(txt.split("\n")))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
h = rows[hi]; ie = h.index("Instructions Executed"); sa = h.index("Warp Stall Sampling (All Samples)")
data = [r for r in rows[hi + 1:] if len(r) == len(h) and r[2] == "-"]
path = [l for l in txt.split("\n") if "File Path" in l][0].split('","')[1].rstrip('",')
src = open(path).read().split("\n")
fn_at, cur = {}, "?"
for i, l in enumerate(src, 1):
    m = re.match(r'\s*(?:template.*)?(?:__device__|extern "C" __global__).*?(\w+)\s*\(', l)
    if m and ("__device__" in l or "__global__" in l): cur = m.group(1)
    fn_at[i] = cur
agg = collections.defaultdict(lambda: [0, 0])
for r in data:
    f = fn_at.get(int(r[0]), "?"); agg[f][0] += int(r[ie] or 0); agg[f][1] += int(r[sa] or 0)
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("attributed warp instructions %d, samples %d" % (ti, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%5.1f%% inst %5.1f%% samp  %s" % (100 * v[0] / ti, 100 * v[1] / ts, k))
