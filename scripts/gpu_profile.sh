#!/bin/bash
# ncu evidence (1 GPU): launch list + full captures of the top kernels of one short bench run (single lane, the
# launches of one generate() back to back), each only after the plain run exits 0.  Output stays small (<64 MiB).
# usage: bash scripts/gpu_profile.sh [tag]
set -u
mkdir -p gpurun_out
TAG=${1:-v4}
CMD="python bench.py --steps 1 --warmup 3 --no-extras --in-flight 1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,launch__grid_size,sm__inst_executed_pipe_tensor.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
cap() {  # name kernel-regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/prof_$1_$TAG -f $CMD > gpurun_out/ncu_$1_$TAG.log 2>&1
  echo "$1 capture rc=$?"
}
# Launch order of one generate() (serialised under ncu): dur_pred 4 x conv_gemm_tc_kernel<2>; pitch / energy 4 + 4 x <0>;
# prenet: bank, proj1, proj2, pre_highway (<0>), 4 x highway (<1>), GRU in-proj (<0>); LSTM in-proj (<0>); rnn_tc;
# lin; postnet: bank, proj1, proj2, pre_highway, 4 x <1>, GRU in-proj; post_proj.  => 21 x <0>, 8 x <1>, 4 x <2>,
# 2 x rnn_cluster_kernel, 3 x gru_small_kernel per generate; warm-up = 3 generates.
cap lstm rnn_tc_kernel 3 1
cap gru rnn_cluster_kernel 7 1
cap grusmall gru_small_kernel 9 3
capm() {  # as cap, matching the MANGLED name (template arguments are not part of ncu's default function name)
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$2 -s $3 -c $4 -o gpurun_out/prof_$1_$TAG -f $CMD > gpurun_out/ncu_$1_$TAG.log 2>&1
  echo "$1 capture rc=$?"
}
capm gemm conv_gemm_tc_kernelILi0E 71 13
capm highway conv_gemm_tc_kernelILi1E 28 1
capm split conv_gemm_tc_kernelILi2E 13 2
CMD="python bench.py --stft-only"
$CMD > gpurun_out/plain_stft.log 2>&1 && cap stft stft_mel 3 1
CMD="python scripts/fp_profile.py"
$CMD > gpurun_out/plain_fp.log 2>&1 && cap attn attention_tc_kernel 30 1
ls -la gpurun_out | head -40
