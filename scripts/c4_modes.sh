run() { python bench.py --workload c4_highcard --steps 5 --warmup 2 --no-configs --no-e2e --no-cpu-baseline --partitions-per-gpu 2 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('c4 $1', '%.1f Grows/s'%(d['value']/1e9), 'step %.3f ms, scan+agg %.3f ms frac %.3f'%(d['ms_per_step'], r['launch_ms'], r['frac']), d['config']['strategy'][:40], d['config']['groups'])"; }
for sl in 16 32; do for w in 2; do export EVQGPU_PART_SLICE_MB=$sl EVQGPU_AGG_WINDOW=$w; run "slice=$sl window=$w"; done; done
