"""GPU: clock64 role timeline of the tap-row-stacked conv (CTA 0), -DDMD_TIMELINE build.  MMA warp: per-slab wait / issue; epilogue
thread 64: tfull wait, first chunk loads, boundary barrier, tile end."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diamond_b200 import ops
dev = torch.device("cuda:0")
envs = 32
wt = (torch.randn(64, 64, 3, 3) / 24).to(dev); wpk, cp = ops.pack_conv_weight(wt, 64, trs=True)
bias = torch.zeros(64, device=dev)
x = torch.randn(envs, 64, 64, 64, device=dev)
film = torch.randn(envs, 128, device=dev) * 0.1
st = ops.gn_stats(x, 32)
opnd = ops.prep_act(x, mode=1, silu=True, stats0=st, gs0=32, film=film)[0]
for name, kw in (("plain", dict(out_gs=0)), ("stats", dict(out_gs=32)), ("stats+resid", dict(out_gs=32, residual=x))):
    buf = torch.zeros(3 * 16 * 16, dtype=torch.int64, device=dev)
    for rep in range(2):
        buf.zero_()
        ops.conv2d_operand(opnd, None, 64, 0, envs, 64, 64, wpk, 64, cp, bias=bias, debug_buf=buf, trs=True, **kw)
        torch.cuda.synchronize()
    b = buf.cpu().view(3, 16, 16)
    t0 = int(b[b > 0].min())
    print(f"==== {name}")
    for it in range(6):
        M = [int(v) - t0 if v > 0 else -1 for v in b[1, it, :15]]
        E = [int(v) - t0 if v > 0 else -1 for v in b[2, it, :10]]
        print(f"tile {it}: MMA slab(w_full,got,issued)x4 {M[:12]} tempty(wait,got) {M[12:14]} end {M[14]}")
        print(f"         EPI wait_tfull {E[0]} got {E[1]} chunk0_loaded {E[2]} after_barrier {E[8]} tile_end {E[3]}")
    bd = b.double()
    slab = [float((bd[1, 1:6, k * 3 + 2] - bd[1, 1:6, k * 3 + 1]).mean()) for k in range(4)]
    swait = [float((bd[1, 1:6, k * 3 + 1] - bd[1, 1:6, k * 3 + 0]).mean()) for k in range(4)]
    print(f"{name}: slab issue {[round(v) for v in slab]} slab wait {[round(v) for v in swait]} mma_tile {float((bd[1,1:6,14]-bd[1,1:6,13]).mean()):.0f} "
          f"tempty wait {float((bd[1,1:6,13]-bd[1,1:6,12]).mean()):.0f} | epi: tfull wait {float((bd[2,1:6,1]-bd[2,1:6,0]).mean()):.0f} "
          f"loads {float((bd[2,1:6,2]-bd[2,1:6,1]).mean()):.0f} barrier {float((bd[2,1:6,8]-bd[2,1:6,2]).mean()):.0f} rest {float((bd[2,1:6,3]-bd[2,1:6,8]).mean()):.0f} "
          f"tile period {float((bd[2,2:6,3]-bd[2,1:5,3]).mean()):.0f}")
