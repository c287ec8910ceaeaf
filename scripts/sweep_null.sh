#!/bin/bash
# launch-shape sweep of a workload's scan kernel: tiles per stage x CTAs/SM x stages
W=${1:-c3_q1_null}
for kt in 1 2; do for c in 2 3 4; do for st in 2 3; do
  r=$(EVQGPU_KT=$kt EVQGPU_MAX_CTAS=$c EVQGPU_NSTAGES=$st timeout 300 python bench.py --workload $W --partitions-per-gpu 2 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%.1f Grows/s launch %.3f ms frac %.3f' % (d['value']/1e9, d['roofline']['launch_ms'], d['roofline']['frac']))")
  echo "kt=$kt ctas=$c stages=$st : $r"
done; done; done
