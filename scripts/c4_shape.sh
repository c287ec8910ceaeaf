#!/bin/bash
# launch-shape sweep of pass 1 of the partitioned hash tier (C4): tiles per stage x consumers x stages
for kt in 1 2; do for nc in 128 256; do for st in 2 3; do
  r=$(EVQGPU_KT=$kt EVQGPU_FAST_NCONS=$nc EVQGPU_NSTAGES=$st timeout 300 python bench.py --workload c4_highcard --partitions-per-gpu 2 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%.2f Grows/s per table %.3f ms' % (d['value']/1e9, d['roofline']['launch_ms']))")
  echo "kt=$kt ncons=$nc stages=$st : $r"
done; done; done
