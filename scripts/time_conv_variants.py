"""GPU: time the dominant conv under debug switches to locate the bottleneck (graph replay, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diamond_b200 import ops
dev = torch.device("cuda:0")
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
wt = (torch.randn(64, 64, 3, 3) / 24).to(dev)
wpk, cp = ops.pack_conv_weight(wt, 64)
bias = torch.zeros(64, device=dev)
xs = [torch.randn(envs, 64, 64, 64, device=dev) for _ in range(4)]
film = torch.randn(envs, 128, device=dev) * 0.1
sts = [ops.gn_stats(x, 32) for x in xs]

def timeit(name, **kw):
    def launch(i):
        ops.conv2d_fprop(xs[i % 4], wpk, 64, cp, 64, bias=bias, **kw)
    for i in range(3): launch(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(12): launch(i)
        torch.cuda.synchronize(); g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(5): g.replay()
        e1.record(side)
    torch.cuda.synchronize()
    print(f"{name:60s} {e0.elapsed_time(e1)/60*1e3:8.1f} us", flush=True)

full = dict(prologue=1, silu=True, stats0=None, gs0=32, film=film, out_gs=32)
def F(i_kw):
    d = dict(full); d.update(i_kw); return d
# stats0 differs per buffer; use buffer 0's stats for all (values irrelevant for timing)
full["stats0"] = sts[0]
timeit("full (AdaGN+SiLU prologue, stats epilogue)", **full)
timeit("no out stats", **F(dict(out_gs=0)))
timeit("no prologue/silu", prologue=0, silu=False, out_gs=0)
timeit("dbg skip MMA", **F(dict(debug=2)))
timeit("dbg skip loads (zeros)", **F(dict(debug=4)))
timeit("dbg skip epilogue stores+stats", **F(dict(debug=8)))
timeit("dbg skip SiLU", **F(dict(debug=16)))
timeit("dbg skip loads+MMA", **F(dict(debug=6)))
timeit("dbg skip loads+epilogue", **F(dict(debug=12)))
timeit("dbg skip MMA+epilogue", **F(dict(debug=10)))
timeit("dbg skip loads+MMA+epilogue (pure pipeline overhead)", **F(dict(debug=14)))
timeit("skeleton + skip fence.proxy.async", **F(dict(debug=14 + 32)))
timeit("skeleton + skip STS", **F(dict(debug=14 + 64)))
timeit("skeleton + skip fence + STS", **F(dict(debug=14 + 96)))
timeit("skeleton + skip tmem_ld", **F(dict(debug=14 + 128)))
timeit("skeleton + skip fence+STS+tmem_ld", **F(dict(debug=14 + 224)))
timeit("skeleton, no stats, skip fence+STS+tmem_ld", **F(dict(debug=14 + 224, out_gs=0)))
timeit("skeleton no prologue, no stats, skip fence+STS+tmem_ld", prologue=0, silu=False, out_gs=0, debug=14 + 224)
w1 = (torch.randn(64, 64, 1, 1) / 8).to(dev); wpk1, cp1 = ops.pack_conv_weight(w1, 64)
def launch1(i): ops.conv2d_fprop(xs[i % 4], wpk1, 64, cp1, 64, 1, bias=bias)
for i in range(3): launch1(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream()
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
        for i in range(12): launch1(i)
    torch.cuda.synchronize(); g.replay(); e0.record(side)
    for _ in range(5): g.replay()
    e1.record(side)
torch.cuda.synchronize()
print(f"{'1x1 conv 64->64 plain (no halo, 1 tap)':60s} {e0.elapsed_time(e1)/60*1e3:8.1f} us")
