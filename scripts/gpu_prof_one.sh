#!/bin/bash
# usage: KERNEL=regex SKIP=n COUNT=n NAME=out bash scripts/gpu_prof_one.sh   (full ncu capture of one kernel)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-extras"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:${KERNEL} -s ${SKIP:-0} -c ${COUNT:-1} -o gpurun_out/${NAME} -f $CMD > gpurun_out/ncu_${NAME}.log 2>&1
echo "capture rc=$?"; tail -2 gpurun_out/ncu_${NAME}.log
