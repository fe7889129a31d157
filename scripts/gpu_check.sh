#!/bin/bash
# Runs every GPU test file in its own process (a faulting kernel must not poison the others),
# then smoke() and a short bench.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in ${TESTS:-test_gpu_length_regulator test_gpu_conv_gemm test_gpu_rnn test_gpu_dsp test_gpu_forward_tacotron test_gpu_fast_pitch test_gpu_attention test_gpu_full_size test_gpu_multi}; do
  echo "=== $f"
  timeout ${TEST_TIMEOUT:-420} python -m pytest tests/$f.py -q -m gpu -x -s --tb=short > gpurun_out/$f.log 2>&1
  echo "exit $?" | tee -a gpurun_out/$f.log
  tail -n 25 gpurun_out/$f.log
done
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  echo "=== smoke"
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit $?" | tee -a gpurun_out/smoke.log; tail -n 5 gpurun_out/smoke.log
  echo "=== bench"
  timeout 600 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit $?"; tail -n 3 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
fi
