# sweep aids of the shared-memory aggregation kernel: threads per CTA (env read by codegen.cc / query.cu)
run() { python bench.py --workload c4_highcard --steps 3 --warmup 2 --no-configs --no-e2e --no-cpu-baseline --partitions-per-gpu 4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('c4 $1', '%.2f Grows/s'%(d['value']/1e9), 'step %.3f ms, per table %.3f ms'%(d['ms_per_step'], r['launch_ms']))"; }
for t in 1024 512 256; do export EVQGPU_AGG_THREADS=$t; run "threads=$t"; done
