#!/bin/bash
# one `ncu --set full` capture of a named kernel of a bench workload + its summary
# usage: scripts/ncu_kernel.sh <workload> <kernel regex> <tag> [launch-skip] [extra bench args]
W=$1; K=$2; TAG=$3; SKIP=${4:-1}; shift 4
mkdir -p gpurun_out/jit
EVQGPU_JIT_DUMP_DIR=gpurun_out/jit EVQGPU_CACHE_DIR= ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $SKIP -c 1 \
  -o gpurun_out/$TAG -f python bench.py --workload $W --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs "$@" > gpurun_out/$TAG.ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/$TAG.ncu-rep --source > gpurun_out/$TAG.txt 2>&1
tail -3 gpurun_out/$TAG.ncu.log
