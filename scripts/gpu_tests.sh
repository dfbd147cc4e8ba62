#!/usr/bin/env bash
# Runs the GPU test files one process each (a trapped kernel poisons its CUDA context, not the others), unbuffered, with a
# per-test watchdog.  Usage: scripts/gpu_tests.sh <out_prefix> [files...]
set -u
out=${1:-gpurun_out/gputests}; shift || true
files=("$@")
if [ ${#files[@]} -eq 0 ]; then files=(tests/test_gpu_wgrad.py tests/test_gpu_conv.py tests/test_actor_critic.py tests/test_gpu_rew_end.py tests/test_gpu_denoiser.py tests/test_gpu_training.py); fi
mkdir -p "$(dirname "$out")"
: > "${out}_summary.txt"
for f in "${files[@]}"; do
  name=$(basename "$f" .py)
  PYTHONUNBUFFERED=1 timeout ${FILE_TIMEOUT:-420} python -u -m pytest "$f" -v -s -m gpu -p no:cacheprovider --timeout ${TEST_TIMEOUT:-180} --timeout-method=thread > "${out}_${name}.log" 2>&1
  rc=$?
  echo "$name rc=$rc $(grep -E '^(=+ )?[0-9]+ (passed|failed)|passed|failed' "${out}_${name}.log" | tail -1)" >> "${out}_summary.txt"
done
cat "${out}_summary.txt"
