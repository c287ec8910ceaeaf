# C4 after a change: parity tests of the hash tiers, the bench line, and the launch list (pass 1 / pass 2 durations)
[ -n "$SKIP_TESTS" ] || python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "partition or highcard or hash" 2>&1 | tail -3
python bench.py --workload c4_highcard --steps 5 --warmup 3 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/c4_bench.json'):
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('c4 %.2f Grows/s step %.3f ms launch_ms %s frac %.3f groups %s' % (d['value'] / 1e9, d['ms_per_step'], r.get('launch_ms'), r['frac'], d['config'].get('groups')), d['config'].get('strategy', '')[:60])
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:evq_ -c 12 --csv --log-file gpurun_out/c4_launches.csv \
  python bench.py --workload c4_highcard --steps 1 --warmup 1 --partitions-per-gpu 2 --no-configs --no-e2e --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open('gpurun_out/c4_launches.csv') if l.startswith('"'))]
h = rows[0]; ki, mi, vi = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value')
ui = h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault((r[0], r[ki]), {})[r[mi]] = (r[vi], r[ui])
for (i, k), m in list(agg.items())[-24:]:
    print(i, k[:30], {a: b for a, b in m.items()})
PY
