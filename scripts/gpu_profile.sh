#!/bin/bash
# ncu evidence (1 GPU): launch list + full captures of the kernels of one short bench run (single lane, the launches of
# one generate() back to back), each only after the plain run exits 0.  Output stays small (<64 MiB).
# usage: [ONLY="lstm gru"] bash scripts/gpu_profile.sh [tag]
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
CMD="python bench.py --steps 1 --warmup 3 --no-extras --in-flight 1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,launch__grid_size,sm__inst_executed_pipe_tensor.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
summ() {  # the text summaries are what travels back (gpurun returns at most 64 MiB): reports above 5 MB are dropped
  python scripts/ncu_summary.py gpurun_out/prof_$1_$TAG.ncu-rep > gpurun_out/sum_$1_$TAG.txt 2>&1
  python scripts/ncu_hot.py gpurun_out/prof_$1_$TAG.ncu-rep 25 > gpurun_out/hot_$1_$TAG.txt 2>&1
  if [ $(stat -c %s gpurun_out/prof_$1_$TAG.ncu-rep) -gt 5000000 ]; then rm -f gpurun_out/prof_$1_$TAG.ncu-rep; fi
  rm -f gpurun_out/ncu_$1_$TAG.log
}
want() { [ -z "${ONLY:-}" ] || [[ " $ONLY " == *" $1 "* ]]; }
cap() {  # name kernel-regex skip count
  want $1 || return 0
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/prof_$1_$TAG -f $CMD > gpurun_out/ncu_$1_$TAG.log 2>&1
  echo "$1 capture rc=$?"
  summ $1
}
capm() {  # as cap, matching the MANGLED name (template arguments are not part of ncu's default function name)
  want $1 || return 0
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$2 -s $3 -c $4 -o gpurun_out/prof_$1_$TAG -f $CMD > gpurun_out/ncu_$1_$TAG.log 2>&1
  echo "$1 capture rc=$?"
  summ $1
}
# Launches of one generate() (serialised under ncu): dur_pred 4 x conv_gemm_tc_kernel<2>; pitch / energy 4 + 4 x <3>;
# prenet: bank, proj1, proj2 (<3>), cbhg_tail_kernel, GRU; LSTM in-proj (<3>); length_index; rnn_tc; lin (<3>);
# postnet: bank, proj1, proj2 (<3>), cbhg_tail_kernel, GRU; post_proj (<3>)  => 17 x <3>, 4 x <2>, 2 x cbhg_tail_kernel,
# 2 x rnn_cluster_kernel, 3 x gru_small_kernel, 1 x rnn_tc_kernel per generate; warm-up = 3 generates.
cap lstm rnn_tc_kernel 3 1
cap gru rnn_cluster_kernel 7 1
cap grusmall gru_small_kernel 9 3
cap tail cbhg_tail_kernel 6 2
cap lenidx length_index_kernel 3 1
capm gemm conv_gemm_tc_kernelILi3E 51 17
capm split conv_gemm_tc_kernelILi2E 13 2
CMD="python bench.py --stft-only"
want stft && $CMD > gpurun_out/plain_stft.log 2>&1 && cap stft stft_mel 3 1
CMD="python scripts/fp_profile.py"
# FastPitch.generate (cfg3), per generate: prenet + postnet = 8 x attention_umma_kernel<128>, 8 x <64> (pitch / energy),
# 4 x fp32 SIMT attention (duration predictor); the frame- and phoneme-rate transformer GEMMs run as CTA pairs:
# 16 x conv_gemm_tc_kernel<3, true> (qkv, conv1) and 16 x <4, true> (out_proj / conv2 with the fused LayerNorm): those and
# the LayerNorm launches are captured by scripts/gpu_profile_fp.sh (launch counts are taken from a first pass)
{ want attn || want attnf32; } && $CMD > gpurun_out/plain_fp.log 2>&1 && {
  capm attn attention_umma_kernelILi128E 20 1
  cap attnf32 attention_kernel 8 1
}
ls -la gpurun_out | grep $TAG; du -sh gpurun_out
