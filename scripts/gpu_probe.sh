#!/usr/bin/env bash
# GPU visit: in-graph kernel timeline (ktrace), the env test, then the clock64 role timeline of the dominant conv on a
# -DDMD_TIMELINE build (rebuilt on the box, production library restored afterwards).
set -u
out=${1:-gpurun_out/p}
timeout 200 python scripts/ktrace.py 32 ${out}_ktrace.csv > ${out}_ktrace.txt 2>&1; tail -45 ${out}_ktrace.txt
timeout 300 python -m pytest tests/test_gpu_denoiser.py -q -m gpu -k "world_model or benchmarked" -p no:cacheprovider > ${out}_env.log 2>&1; tail -5 ${out}_env.log
cp diamond_b200/libdiamond_b200.so /tmp/prod.so
DMD_EXTRA=-DDMD_TIMELINE bash diamond_b200/csrc/build.sh > /dev/null 2>&1
DMD_CONV_EPI=0 DMD_CONV_GROUPS=1 timeout 120 python scripts/timeline_conv.py > ${out}_timeline_g1.txt 2>&1; tail -8 ${out}_timeline_g1.txt
DMD_CONV_EPI=1 timeout 120 python scripts/timeline_conv.py > ${out}_timeline_direct.txt 2>&1; tail -8 ${out}_timeline_direct.txt
cp /tmp/prod.so diamond_b200/libdiamond_b200.so
