#!/bin/bash
# ncu launch list (per-launch durations + grid) of one short bench run.  usage: bash scripts/gpu_launchlist.sh [name]
set -u
mkdir -p gpurun_out
NAME=${1:-launches}
CMD="python bench.py --steps 1 --warmup 3 --no-extras --in-flight 1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,launch__grid_size,sm__inst_executed_pipe_tensor.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${NAME}.csv $CMD > gpurun_out/ncu_${NAME}.log 2>&1
echo "launch list rc=$?"
