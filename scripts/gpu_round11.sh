#!/usr/bin/env bash
# Final verification visit: the driver's test command, smoke(), both bench arms, refreshed ncu evidence for the training
# kernels, then compute-sanitizer memcheck over the code paths added this round (bounded).
set -u
out=${1:-gpurun_out/r11}
BUDGET=${BUDGET:-420}
mkdir -p "$(dirname "$out")"
t0=$SECONDS
stamp() { echo "[t+$((SECONDS - t0))s] $*"; }
lim() { local want=$1; local left=$((BUDGET - (SECONDS - t0))); if [ $left -lt 20 ]; then echo 0; elif [ $want -lt $left ]; then echo $want; else echo $left; fi; }
run() { local t; t=$(lim $1); shift; if [ "$t" = 0 ]; then echo "SKIPPED (budget): $*" | cut -c1-120; return 99; fi; timeout $t "$@"; }
PYTHONUNBUFFERED=1 run 300 python -m pytest tests -x -q -s -m gpu -p no:cacheprovider --timeout 200 --timeout-method=thread > ${out}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 ${out}_pytest.log; stamp tests
run 120 python -c "import __graft_entry__ as g; g.smoke()" > ${out}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 ${out}_smoke.log; stamp smoke
run 300 python bench.py > ${out}_bench.json 2> ${out}_bench.err; echo "bench rc=$?"; tail -c 300 ${out}_bench.err
python - ${out}_bench.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("frames/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic_source"], "incomplete", d.get("incomplete"))
print("train", json.dumps(d.get("train_denoiser"))[:400])
print("imag", json.dumps(d.get("imagination_update"))[:300])
PY
stamp bench
run 120 python bench.py --impl reference > ${out}_bench_ref.json 2> ${out}_bench_ref.err; echo "ref rc=$?"; head -c 400 ${out}_bench_ref.json; echo; stamp ref
run 120 ncu --clock-control none --set full --import-source on -k regex:'wgrad_tc_kernel|wgrad_reduce' --launch-skip 4 -c 4 -o ${out}_wgrad -f python scripts/prof_wgrad.py 64 > ${out}_ncu_wgrad.log 2>&1; tail -1 ${out}_ncu_wgrad.log; stamp ncu-wgrad
run 150 ncu --clock-control none --metrics gpu__time_duration.sum -c 2500 --csv --log-file ${out}_launches_train.csv python scripts/prof_train.py 64 > ${out}_launches_train.log 2>&1; wc -l ${out}_launches_train.csv; stamp train-launches
PYTHONUNBUFFERED=1 run 170 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_denoiser.py tests/test_gpu_training.py -q -m gpu -p no:cacheprovider -k "(padded and reference_golden and not sampler) or optimizer" > ${out}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" ${out}_memcheck.log | tail -5; stamp memcheck
