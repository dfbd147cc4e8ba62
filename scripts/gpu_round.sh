#!/usr/bin/env bash
# One GPU-box visit: conv/wgrad tests first (they decide whether the rest runs on the direct or the staged epilogue), then the
# model-level tests, the actor-critic gradient diagnostic and two short bench runs (direct vs staged epilogue).
set -u
out=${1:-gpurun_out/r}
export FILE_TIMEOUT=${FILE_TIMEOUT:-200} TEST_TIMEOUT=${TEST_TIMEOUT:-90}
bash scripts/gpu_tests.sh ${out}_k tests/test_gpu_conv.py tests/test_gpu_wgrad.py
if grep -q "test_gpu_conv rc=0" ${out}_k_summary.txt && grep -q "test_gpu_wgrad rc=0" ${out}_k_summary.txt; then
  echo "kernel tests green on the direct epilogue"
else
  echo "kernel tests FAILED on the direct epilogue: continuing with DMD_CONV_EPI=0"; export DMD_CONV_EPI=0
  bash scripts/gpu_tests.sh ${out}_k0 tests/test_gpu_conv.py tests/test_gpu_wgrad.py
fi
bash scripts/gpu_tests.sh ${out}_m tests/test_actor_critic.py tests/test_gpu_rew_end.py tests/test_gpu_denoiser.py tests/test_gpu_training.py
timeout 200 python scripts/diag_ac.py 4 > ${out}_diag_ac.log 2>&1; tail -40 ${out}_diag_ac.log
timeout 500 python bench.py --skip-cpu-baseline --skip-gpu-baseline --steps 10 > ${out}_bench.json 2> ${out}_bench.err; tail -c 2500 ${out}_bench.json; tail -3 ${out}_bench.err
DMD_CONV_EPI=0 timeout 200 python bench.py --skip-cpu-baseline --skip-gpu-baseline --skip-train --skip-imagination --steps 10 > ${out}_bench_epi0.json 2> ${out}_bench_epi0.err; tail -c 1500 ${out}_bench_epi0.json
