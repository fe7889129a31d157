set -u
mkdir -p gpurun_out
TAG=v4
CMD="python bench.py --steps 1 --warmup 3 --no-extras --in-flight 1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,launch__grid_size,sm__inst_executed_pipe_tensor.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
cap() { ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/prof_$1_$TAG -f $CMD > gpurun_out/ncu_$1_$TAG.log 2>&1; echo "$1 capture rc=$?"; }
capm() { ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$2 -s $3 -c $4 -o gpurun_out/prof_$1_$TAG -f $CMD > gpurun_out/ncu_$1_$TAG.log 2>&1; echo "$1 capture rc=$?"; }
cap grusmall gru_small_kernel 9 3
capm gemm conv_gemm_tc_kernelILi0E 71 13
capm highway conv_gemm_tc_kernelILi1E 28 1
capm split conv_gemm_tc_kernelILi2E 13 2
ls -la gpurun_out | grep v4
