#!/usr/bin/env bash
# GPU visit: all GPU tests, in-graph timeline, bench.
set -u
out=${1:-gpurun_out/r}
export FILE_TIMEOUT=${FILE_TIMEOUT:-240} TEST_TIMEOUT=${TEST_TIMEOUT:-150}
bash scripts/gpu_tests.sh ${out}_t
timeout 200 python scripts/ktrace.py 32 ${out}_ktrace.csv > ${out}_ktrace.txt 2>&1; tail -36 ${out}_ktrace.txt
timeout 500 python bench.py --skip-cpu-baseline --skip-gpu-baseline --steps 10 > ${out}_bench.json 2> ${out}_bench.err; tail -c 2300 ${out}_bench.json; tail -2 ${out}_bench.err
