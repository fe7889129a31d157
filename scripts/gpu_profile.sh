#!/bin/bash
# ncu evidence (1 GPU): full captures of the top kernels of one short bench run, each only after the plain run exits 0.
# Output stays small (<64 MiB): one or two launches per capture.   usage: bash scripts/gpu_profile.sh [tag]
set -u
mkdir -p gpurun_out
TAG=${1:-v2}
CMD="python bench.py --steps 1 --warmup 3 --no-extras"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
cap() {  # name kernel-regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/prof_$1_$TAG -f $CMD > gpurun_out/ncu_$1_$TAG.log 2>&1
  echo "$1 capture rc=$?"
}
# per generate: 25 conv_gemm_tc launches (8 stage A, 9 prenet, LSTM in-proj = #17, lin, postnet bank, proj1 = #20 ...)
cap lstm rnn_tc_kernel 3 1
cap gru rnn_cluster_kernel 7 1
cap gemm conv_gemm_tc_kernel 92 4
CMD="python bench.py --stft-only"
$CMD > gpurun_out/plain_stft.log 2>&1 && cap stft stft_mel 3 1
ls -la gpurun_out
