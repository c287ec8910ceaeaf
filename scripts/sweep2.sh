#!/bin/bash
W=${1:-c3_q1}
for c in 3 4 5; do for st in 2 3 4; do
  r=$(EVQGPU_MAX_CTAS=$c EVQGPU_NSTAGES=$st timeout 300 python bench.py --workload $W --partitions-per-gpu 2 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%.1f Grows/s launch %.3f ms frac %.3f' % (d['value']/1e9, d['roofline']['launch_ms'], d['roofline']['frac']))")
  echo "ctas=$c stages=$st : $r"
done; done
