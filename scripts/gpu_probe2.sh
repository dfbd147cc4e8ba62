#!/usr/bin/env bash
set -u
out=${1:-gpurun_out/p}
timeout 400 python scripts/diag_imag.py > ${out}_diag_imag.txt 2>&1; grep -v "^$" ${out}_diag_imag.txt | head -90 | cut -c1-170
cp diamond_b200/libdiamond_b200.so /tmp/prod.so
DMD_EXTRA=-DDMD_TIMELINE bash diamond_b200/csrc/build.sh > /dev/null 2>&1
timeout 120 python scripts/timeline_trs.py > ${out}_timeline_trs.txt 2>&1; cat ${out}_timeline_trs.txt | cut -c1-250 | tail -45
cp /tmp/prod.so diamond_b200/libdiamond_b200.so
