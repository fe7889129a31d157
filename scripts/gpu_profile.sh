#!/bin/bash
# ncu evidence: launch list of a short bench run + full captures of the top kernels (1 GPU).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rnn_cluster -c 3 -o gpurun_out/prof_rnn -f $CMD > gpurun_out/ncu_rnn.log 2>&1
echo "rnn capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc -s 30 -c 8 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
CMD2="python bench.py --stft-only"
$CMD2 > gpurun_out/plain_stft.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stft_mel -s 3 -c 1 -o gpurun_out/prof_stft -f $CMD2 > gpurun_out/ncu_stft.log 2>&1
echo "stft capture rc=$?"
ls -la gpurun_out
