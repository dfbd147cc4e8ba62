#!/usr/bin/env bash
# GPU visit: pad/crop parity tests, then short A/B benches of the opt-in variants (tap-row-stacked conv, fused small convs).
set -u
out=${1:-gpurun_out/r8}
mkdir -p "$(dirname "$out")"
PYTHONUNBUFFERED=1 timeout 300 python -m pytest tests/test_gpu_denoiser.py -x -q -s -m gpu -k "padded" -p no:cacheprovider --timeout 120 > ${out}_pytest_padded.log 2>&1
echo "padded rc=$?"; grep -E "rel L2|one level|max\|diff|passed|failed|Error|error" ${out}_pytest_padded.log | cut -c1-200 | tail -12
B="--skip-cpu-baseline --skip-gpu-baseline --skip-train --skip-imagination --steps 10"
for v in "base" "DMD_CONV_TRS=1" "DMD_FUSE_SMALL=1" "DMD_CONV_TRS=1 DMD_TRS_GROUPS=1"; do
  tag=$(echo "$v" | tr ' =' '__')
  if [ "$v" = base ]; then timeout 200 python bench.py $B > ${out}_bench_${tag}.json 2> ${out}_bench_${tag}.err
  else env $v timeout 200 python bench.py $B > ${out}_bench_${tag}.json 2> ${out}_bench_${tag}.err; fi
  python - "$v" ${out}_bench_${tag}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    r = d["roofline"]
    print(f"{sys.argv[1]:36s} frames/s {d['value']:8.0f}  e2e {d['e2e']['value']:8.0f}  conv us {r['us_per_launch']:6.2f} frac {r['frac']:.3f}  prep us {r['prep_us_per_launch']:.2f}  launches {d['gpu_launches']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
