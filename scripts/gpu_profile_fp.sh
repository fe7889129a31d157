#!/bin/bash
# ncu --set full summaries of the FastPitch (cfg3) frame-rate GEMMs and LayerNorms: the last ~20 GEMM launches of the last
# generate() of scripts/fp_profile.py (postnet layers: qkv, out_proj + LayerNorm, conv1, conv2 + LayerNorm; lin).
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
CMD="python scripts/fp_profile.py"
$CMD > gpurun_out/plain_fp.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_fp.log; exit 1; }
N=$(ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_gemm_tc_kernel -c 2000 --csv $CMD 2>/dev/null | grep -c conv_gemm_tc_kernel)
echo "conv_gemm launches per run: $N"
SKIP=$((N - 18))
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc_kernel -s $SKIP -c 18 -o gpurun_out/prof_fpgemm_$TAG -f $CMD > gpurun_out/ncu_fpgemm_$TAG.log 2>&1
echo "fpgemm capture rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_fpgemm_$TAG.ncu-rep > gpurun_out/sum_fpgemm_$TAG.txt 2>&1
python scripts/ncu_hot.py gpurun_out/prof_fpgemm_$TAG.ncu-rep 25 > gpurun_out/hot_fpgemm_$TAG.txt 2>&1
rm -f gpurun_out/prof_fpgemm_$TAG.ncu-rep
M=$(ncu --metrics gpu__time_duration.sum --clock-control none -k regex:layernorm_kernel -c 2000 --csv $CMD 2>/dev/null | grep -c layernorm_kernel)
echo "layernorm launches per run: $M"
ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s $((M - 3)) -c 3 -o gpurun_out/prof_layernorm_$TAG -f $CMD > gpurun_out/ncu_layernorm_$TAG.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_layernorm_$TAG.ncu-rep > gpurun_out/sum_layernorm_$TAG.txt 2>&1
python scripts/ncu_hot.py gpurun_out/prof_layernorm_$TAG.ncu-rep 25 > gpurun_out/hot_layernorm_$TAG.txt 2>&1
ls -la gpurun_out | grep $TAG
