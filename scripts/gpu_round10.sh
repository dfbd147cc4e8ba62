#!/usr/bin/env bash
# 2-GPU visit: the driver's test command, the driver's N=2 launch of bench.py (NCCL: gradient all-reduce of both training blocks),
# then a short N=1 bench.
set -u
out=${1:-gpurun_out/r10}
mkdir -p "$(dirname "$out")"
nvidia-smi --query-gpu=index,name --format=csv,noheader
PYTHONUNBUFFERED=1 timeout 600 python -m pytest tests -x -q -s -m gpu -p no:cacheprovider --timeout 200 --timeout-method=thread > ${out}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 ${out}_pytest.log
t0=$SECONDS
NCCL_DEBUG=WARN timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > ${out}_bench_n2.json 2> ${out}_bench_n2.err
echo "bench n2 rc=$? in $((SECONDS - t0)) s"; tail -c 400 ${out}_bench_n2.err
t0=$SECONDS
timeout 300 python bench.py --skip-gpu-baseline --skip-cpu-baseline > ${out}_bench_n1.json 2> ${out}_bench_n1.err; echo "bench n1 rc=$? in $((SECONDS - t0)) s"
python - ${out}_bench_n2.json ${out}_bench_n1.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "frames/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"], 4), "incomplete", d.get("incomplete"))
        print("  train", json.dumps(d.get("train_denoiser"))[:600])
        print("  imag", json.dumps(d.get("imagination_update"))[:500])
    except Exception as e:
        print(f, "FAILED", repr(e))
PY
