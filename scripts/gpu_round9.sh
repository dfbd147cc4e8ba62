#!/usr/bin/env bash
# GPU visit: the driver's test command, the full default bench, the imagination-step diagnostic.
set -u
out=${1:-gpurun_out/r9}
mkdir -p "$(dirname "$out")"
PYTHONUNBUFFERED=1 timeout 600 python -m pytest tests -x -q -s -m gpu -p no:cacheprovider --timeout 200 --timeout-method=thread > ${out}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 ${out}_pytest.log
timeout 420 python bench.py > ${out}_bench.json 2> ${out}_bench.err; echo "bench rc=$?"; tail -c 300 ${out}_bench.err
python - ${out}_bench.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("frames/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"], 4), "incomplete", d.get("incomplete"))
print("train", json.dumps(d.get("train_denoiser"))[:700])
print("imag", json.dumps(d.get("imagination_update"))[:500])
print("gpu_base", json.dumps(d.get("gpu_baseline"))[:400])
PY
timeout 300 python scripts/diag_imag.py > ${out}_diag_imag.txt 2>&1; tail -60 ${out}_diag_imag.txt
