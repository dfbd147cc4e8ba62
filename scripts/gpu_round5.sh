#!/usr/bin/env bash
# GPU visit: kernel tests (incl. the tap-row-stacked conv); model tests with it on (off if the kernel tests fail); in-graph timeline;
# bench; imagination-step diagnostic.
set -u
out=${1:-gpurun_out/r}
export FILE_TIMEOUT=${FILE_TIMEOUT:-240} TEST_TIMEOUT=${TEST_TIMEOUT:-150}
bash scripts/gpu_tests.sh ${out}_k tests/test_gpu_conv.py tests/test_gpu_wgrad.py
if ! grep -q "test_gpu_conv rc=0" ${out}_k_summary.txt; then
  grep -E "FAILED|PASSED" ${out}_k_test_gpu_conv.log | grep -c PASSED
  grep -E "^FAILED|Error|error:|assert " ${out}_k_test_gpu_conv.log | head -30
  echo "conv tests failed: model tests run with DMD_CONV_TRS=0"; export DMD_CONV_TRS=0
fi
bash scripts/gpu_tests.sh ${out}_m tests/test_gpu_denoiser.py tests/test_gpu_rew_end.py tests/test_actor_critic.py tests/test_gpu_training.py
grep -E "rel L2|one level off|max\|diff" ${out}_m_test_gpu_denoiser.log | cut -c1-160 | head -12
timeout 200 python scripts/ktrace.py 32 ${out}_ktrace.csv > ${out}_ktrace.txt 2>&1; tail -34 ${out}_ktrace.txt
timeout 500 python bench.py --skip-cpu-baseline --skip-gpu-baseline --steps 10 > ${out}_bench.json 2> ${out}_bench.err; tail -c 2200 ${out}_bench.json; tail -2 ${out}_bench.err
timeout 300 python scripts/diag_imag.py > ${out}_diag_imag.txt 2>&1; tail -14 ${out}_diag_imag.txt
