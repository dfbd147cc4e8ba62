"""GPU: in-graph kernel timeline of one DiffusionSampler.sample() (bench.py workload: 32 envs, 3 Euler steps).

Every conv / prep / attention / wrap launch stamps the GPU nanosecond timer when its inputs are ready (dmd_ktrace_*); the
difference of consecutive stamps is that kernel's duration INSIDE the CUDA graph (programmatic dependent launch, warm L2),
which is what ncu's serialised cold-cache list cannot show.  Prints the time per kernel class / problem size.
usage: python scripts/ktrace.py [envs] [out.csv]"""
import collections
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    from diamond_b200 import _lib
    from diamond_b200.models.diffusion import Denoiser, DenoiserConfig, DiffusionSampler, DiffusionSamplerConfig, InnerModelConfig
    from diamond_b200.synthetic import frame_stacks, randomize_module_

    envs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda:0")
    lib = _lib.lib()
    den = Denoiser(DenoiserConfig(InnerModelConfig(3, 4, 256, [2, 2, 2, 2], [64] * 4, [0] * 4, 4), 0.5, 0.3))
    randomize_module_(den.inner_model, 2024)
    den = den.to(dev).eval()
    sampler = DiffusionSampler(den, DiffusionSamplerConfig(3))
    obs, act, _ = frame_stacks(envs, 4, 3, 64, 64, 4, 100)
    obs, act = obs.to(dev), act.to(dev)
    cap = 4096
    _lib.check(lib.dmd_ktrace_begin(cap))      # the graph captured by the first sample() bakes the trace slots in
    for _ in range(4):                         # capture + replays: the stamps of the LAST replay survive
        sampler.sample(obs, act)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sampler.sample(obs, act); e1.record()
    torch.cuda.synchronize()
    buf = (C.c_longlong * cap)()
    n = lib.dmd_ktrace_end(buf, cap)
    names = [lib.dmd_ktrace_name(i).decode() for i in range(n)]
    t = [buf[i] for i in range(n)]
    print(f"traced launches: {n}; sample() incl. host copies {e0.elapsed_time(e1) * 1e3:.1f} us; first->last stamp {(t[-1] - t[0]) / 1e3:.1f} us")
    agg = collections.OrderedDict()
    rows = []
    for i in range(n - 1):
        d = (t[i + 1] - t[i]) / 1e3
        rows.append((i, names[i], d))
        k = names[i]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += d
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel (aux = K*1000+W for convs, mode*1000+W for preps)':64s} {'n':>4s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
    for k, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:64s} {c:4d} {s:10.1f} {s / c:8.2f} {100 * s / tot:5.1f}%")
    by = collections.defaultdict(float)
    for k, (c, s) in agg.items():
        by[k.split()[0]] += s
    print("by class:", {k: round(v, 1) for k, v in by.items()}, "total", round(tot, 1))
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            f.write("index,kernel,us_until_next_stamp\n")
            for i, k, d in rows:
                f.write(f"{i},{k},{d:.3f}\n")


if __name__ == "__main__":
    main()
