#!/bin/bash
# one `ncu --set full` capture of the scan kernel of a bench workload + its summary (scripts/ncu_summary.py)
# usage: scripts/ncu_capture.sh <workload> <tag> [extra bench args]   -> gpurun_out/<tag>.ncu-rep, gpurun_out/<tag>.txt
W=$1; TAG=$2; shift 2
mkdir -p gpurun_out/jit
EVQGPU_JIT_DUMP_DIR=gpurun_out/jit EVQGPU_CACHE_DIR= ncu --set full --clock-control none --import-source on -k regex:evq_scan --launch-skip 3 -c 1 \
  -o gpurun_out/$TAG -f python bench.py --workload $W --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs "$@" > gpurun_out/$TAG.ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/$TAG.ncu-rep --source > gpurun_out/$TAG.txt 2>&1
tail -3 gpurun_out/$TAG.ncu.log
