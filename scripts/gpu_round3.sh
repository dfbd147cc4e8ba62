#!/usr/bin/env bash
# GPU visit: all GPU tests, in-graph timeline, bench (+ roofline with 128-byte aligned slab bases), clock64 timeline of the
# dominant conv with aligned / unaligned operand addresses on a -DDMD_TIMELINE build.
set -u
out=${1:-gpurun_out/r}
export FILE_TIMEOUT=${FILE_TIMEOUT:-240} TEST_TIMEOUT=${TEST_TIMEOUT:-150}
bash scripts/gpu_tests.sh ${out}_t
timeout 200 python scripts/ktrace.py 32 ${out}_ktrace.csv > ${out}_ktrace.txt 2>&1; tail -36 ${out}_ktrace.txt
timeout 500 python bench.py --skip-cpu-baseline --skip-gpu-baseline --steps 10 > ${out}_bench.json 2> ${out}_bench.err; tail -c 2300 ${out}_bench.json; tail -2 ${out}_bench.err
DMD_CONV_PALLOC8=1 timeout 200 python bench.py --skip-cpu-baseline --skip-gpu-baseline --skip-train --skip-imagination --steps 10 > ${out}_bench_palloc8.json 2>/dev/null; python -c "
import json,sys
d=json.load(open('${out}_bench_palloc8.json')); print('PALLOC8: value', d['value'], 'conv us', d['roofline']['us_per_launch'])"
cp diamond_b200/libdiamond_b200.so /tmp/prod.so
DMD_EXTRA=-DDMD_TIMELINE bash diamond_b200/csrc/build.sh > /dev/null 2>&1
timeout 120 python scripts/timeline_conv.py > ${out}_timeline.txt 2>&1; grep -A1 "^noMMA\|^aligned\|^skip-stores" ${out}_timeline.txt | cut -c1-150
DMD_CONV_PALLOC8=1 timeout 120 python scripts/timeline_conv.py > ${out}_timeline_palloc8.txt 2>&1; grep -A1 "^noMMA\|^aligned\|^skip-stores" ${out}_timeline_palloc8.txt | cut -c1-150
cp /tmp/prod.so diamond_b200/libdiamond_b200.so
