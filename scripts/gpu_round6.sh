#!/usr/bin/env bash
# GPU visit: kernel + model tests, in-graph timeline with the row-stacked and the tap-major conv, bench.
set -u
out=${1:-gpurun_out/r}
export FILE_TIMEOUT=${FILE_TIMEOUT:-240} TEST_TIMEOUT=${TEST_TIMEOUT:-150}
bash scripts/gpu_tests.sh ${out}_k tests/test_gpu_conv.py tests/test_gpu_wgrad.py
if ! grep -q "test_gpu_conv rc=0" ${out}_k_summary.txt; then
  grep -E "^FAILED|Error|error:|assert " ${out}_k_test_gpu_conv.log | head -30
  echo "conv tests failed: rerun with DMD_TRS_GROUPS=1"; DMD_TRS_GROUPS=1 bash scripts/gpu_tests.sh ${out}_k1 tests/test_gpu_conv.py
  if grep -q "test_gpu_conv rc=0" ${out}_k1_summary.txt; then export DMD_TRS_GROUPS=1; else export DMD_CONV_TRS=0; fi
fi
bash scripts/gpu_tests.sh ${out}_m tests/test_gpu_denoiser.py tests/test_gpu_rew_end.py tests/test_actor_critic.py tests/test_gpu_training.py
timeout 200 python scripts/ktrace.py 32 ${out}_ktrace.csv > ${out}_ktrace.txt 2>&1; tail -32 ${out}_ktrace.txt
DMD_CONV_TRS=0 timeout 200 python scripts/ktrace.py 32 ${out}_ktrace_tapmajor.csv > ${out}_ktrace_tapmajor.txt 2>&1; grep "conv3x3\|by class\|traced" ${out}_ktrace_tapmajor.txt
timeout 500 python bench.py --skip-cpu-baseline --skip-gpu-baseline --steps 10 > ${out}_bench.json 2> ${out}_bench.err; tail -c 2200 ${out}_bench.json; tail -2 ${out}_bench.err
