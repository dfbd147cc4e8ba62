#!/usr/bin/env bash
# GPU visit: kernel tests, model tests (fused small-conv path on; rerun with DMD_FUSE_SMALL=0 if it fails), in-graph timeline, bench.
set -u
out=${1:-gpurun_out/r}
export FILE_TIMEOUT=${FILE_TIMEOUT:-240} TEST_TIMEOUT=${TEST_TIMEOUT:-150}
bash scripts/gpu_tests.sh ${out}_k tests/test_gpu_conv.py tests/test_gpu_wgrad.py
bash scripts/gpu_tests.sh ${out}_m tests/test_gpu_denoiser.py tests/test_gpu_rew_end.py tests/test_actor_critic.py
if ! grep -q "test_gpu_denoiser rc=0" ${out}_m_summary.txt; then
  echo "model tests FAILED with the fused small-conv path: rerunning with DMD_FUSE_SMALL=0"
  grep -E "FAILED|Error|error|assert" ${out}_m_test_gpu_denoiser.log | head -20
  DMD_FUSE_SMALL=0 bash scripts/gpu_tests.sh ${out}_m0 tests/test_gpu_denoiser.py
  export DMD_FUSE_SMALL=0
fi
timeout 200 python scripts/ktrace.py 32 ${out}_ktrace.csv > ${out}_ktrace.txt 2>&1; tail -42 ${out}_ktrace.txt
timeout 500 python bench.py --skip-cpu-baseline --skip-gpu-baseline --steps 10 > ${out}_bench.json 2> ${out}_bench.err; tail -c 2600 ${out}_bench.json; tail -3 ${out}_bench.err
bash scripts/gpu_tests.sh ${out}_t tests/test_gpu_training.py
