"""GPU: conv time vs number of images for skeleton / full variants (fixed cost vs per-tile slope)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diamond_b200 import ops
dev = torch.device("cuda:0")
wt = (torch.randn(64, 64, 3, 3) / 24).to(dev)
wpk, cp = ops.pack_conv_weight(wt, 64)
bias = torch.zeros(64, device=dev)

def timeit(envs, hw, **kw):
    xs = [torch.randn(envs, hw, hw, 64, device=dev) for _ in range(3)]
    film = torch.randn(envs, 128, device=dev) * 0.1
    st = ops.gn_stats(xs[0], 32)
    if kw.get("prologue"): kw = dict(kw, stats0=st, gs0=32, film=film)
    def launch(i): ops.conv2d_fprop(xs[i % 3], wpk, 64, cp, 64, bias=bias, **kw)
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
    return e0.elapsed_time(e1) / 60 * 1e3

for envs in (32, 64):
    a = timeit(envs, 64, prologue=0, silu=False, out_gs=0, debug=14 + 224)
    b = timeit(envs, 64, prologue=0, silu=False, out_gs=0, debug=14 + 224 + 256)
    c = timeit(envs, 64, prologue=0, silu=False, out_gs=0, debug=14 + 224 + 512)
    print(f"envs={envs} skeleton={a:6.1f}  no-empty-handshake={b:6.1f}  arrive-instead-of-commit={c:6.1f}", flush=True)
for hw in (64, 8):
    for envs in ((1, 2, 4, 8, 16, 32, 64) if hw == 64 else (32, 256)):
        tiles = (envs * (hw + 1) * (hw + 1) + 127) // 128
        sk = timeit(envs, hw, prologue=0, silu=False, out_gs=0, debug=14 + 224)
        pl = timeit(envs, hw, prologue=0, silu=False, out_gs=0)
        fu = timeit(envs, hw, prologue=1, silu=True, out_gs=32)
        print(f"hw={hw} envs={envs:3d} tiles={tiles:5d} tiles/cta={tiles/148:5.2f}  skeleton={sk:6.1f}us  plain={pl:6.1f}us  full={fu:6.1f}us", flush=True)
