#!/usr/bin/env bash
# GPU visit (evidence run): the driver's own test command, the full default bench, ncu launch lists of the sampler bench and
# of one training step, `ncu --set full` captures of the dominant kernels, the in-graph timeline.  Every step is bounded by
# its own timeout AND by what is left of the overall budget ($BUDGET seconds), so the visit never runs into gpurun's limit.
set -u
out=${1:-gpurun_out/r7}
BUDGET=${BUDGET:-1080}
mkdir -p "$(dirname "$out")"
NCU="ncu --clock-control none"
t0=$SECONDS
stamp() { echo "[t+$((SECONDS - t0))s] $*"; }
lim() { local want=$1; local left=$((BUDGET - (SECONDS - t0))); if [ $left -lt 25 ]; then echo 0; elif [ $want -lt $left ]; then echo $want; else echo $left; fi; }
run() { local t; t=$(lim $1); shift; if [ "$t" = 0 ]; then echo "SKIPPED (budget): $*" | cut -c1-120; return 99; fi; timeout $t "$@"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
PYTHONUNBUFFERED=1 run 600 python -m pytest tests -x -q -s -m gpu -p no:cacheprovider --timeout 200 --timeout-method=thread > ${out}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 ${out}_pytest.log; stamp tests
run 420 python bench.py > ${out}_bench.json 2> ${out}_bench.err; echo "bench rc=$?"; tail -c 600 ${out}_bench.err; head -c 1800 ${out}_bench.json; echo; stamp bench
run 300 $NCU --metrics gpu__time_duration.sum -c 1100 --csv --log-file ${out}_launches_bench.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-gpu-baseline --skip-train --skip-imagination > ${out}_launches_bench.log 2>&1; wc -l ${out}_launches_bench.csv; stamp launches
run 240 $NCU --set full --import-source on -k regex:'conv_tc_kernel|prep_' --launch-skip 8 -c 6 -o ${out}_conv -f python scripts/prof_conv.py 32 > ${out}_ncu_conv.log 2>&1; tail -2 ${out}_ncu_conv.log; stamp ncu-conv
run 150 python scripts/ktrace.py 32 ${out}_ktrace.csv > ${out}_ktrace.txt 2>&1; tail -40 ${out}_ktrace.txt; stamp ktrace
run 240 $NCU --set full --import-source on -k regex:'wgrad_tc_kernel|wgrad_reduce' --launch-skip 4 -c 4 -o ${out}_wgrad -f python scripts/prof_wgrad.py 64 > ${out}_ncu_wgrad.log 2>&1; tail -2 ${out}_ncu_wgrad.log; stamp ncu-wgrad
run 300 $NCU --metrics gpu__time_duration.sum -c 2500 --csv --log-file ${out}_launches_train.csv python scripts/prof_train.py 64 > ${out}_launches_train.log 2>&1; wc -l ${out}_launches_train.csv; tail -2 ${out}_launches_train.log | cut -c1-400; stamp train-launches
ls -la gpurun_out | head -40
