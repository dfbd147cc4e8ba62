"""GPU: clock64 timeline of CTA 0's roles for one conv launch (bring-up instrumentation).

Needs a library built with the instrumentation compiled in:  DMD_EXTRA=-DDMD_TIMELINE bash diamond_b200/csrc/build.sh
(the production build compiles it out: the MMA / producer warps are latency-bound on their scalar instructions).
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diamond_b200 import ops
dev = torch.device("cuda:0")
envs = 32
wt = (torch.randn(64, 64, 3, 3) / 24).to(dev); wpk, cp = ops.pack_conv_weight(wt, 64)
bias = torch.zeros(64, device=dev)
x = torch.randn(envs, 64, 64, 64, device=dev)
film = torch.randn(envs, 128, device=dev) * 0.1
st = ops.gn_stats(x, 32)
opnd = ops.prep_act(x, mode=1, silu=True, stats0=st, gs0=32, film=film)[0]
for name, kw in (("plain", dict(out_gs=0)), ("stats", dict(out_gs=32)), ("stats+resid", dict(out_gs=32, residual=x))):
    for rep in range(2):
        buf = torch.zeros(3 * 16 * 16, dtype=torch.int64, device=dev)
        ops.conv2d_operand(opnd, None, 64, 0, envs, 64, 64, wpk, 64, cp, bias=bias, debug_buf=buf, **kw)
        torch.cuda.synchronize()
    b = buf.cpu().view(3, 16, 16)
    t0 = int(b[b > 0].min())
    print(f"==== {name}: cycles relative to first stamp (CTA 0)")
    for it in range(8):
        L = [int(v) - t0 if v > 0 else -1 for v in b[0, it, :12]]
        M = [int(v) - t0 if v > 0 else -1 for v in b[1, it, :15]]
        E = [int(v) - t0 if v > 0 else -1 for v in b[2, it, :4]]
        print(f"tile {it}: PROD(w_empty,got,issued)x4 {L}")
        print(f"         MMA(w_full,got,committed)x4 {M[:12]}  tempty(wait,got) {M[12:14]} tfull_commit {M[14]}")
        print(f"         EPI(wait_tfull,got,pass1_done,pass2_done) {E}")

print("==== summary (cycles, mean over tiles 1..5 of CTA 0)")
for name, kw in (("plain", dict(out_gs=0)), ("stats", dict(out_gs=32)), ("resid", dict(out_gs=0, residual=x)), ("stats+resid", dict(out_gs=32, residual=x)),
                 ("skip-stores(dbg8)", dict(out_gs=0, debug=8)), ("noMMA(dbg2)", dict(out_gs=0, debug=8 | 2)),
                 ("aligned-taps(dbg16)", dict(out_gs=0, debug=8 | 16))):
    buf = torch.zeros(3 * 16 * 16, dtype=torch.int64, device=dev)
    for rep in range(2):
        buf.zero_()
        ops.conv2d_operand(opnd, None, 64, 0, envs, 64, 64, wpk, 64, cp, bias=bias, debug_buf=buf, **kw)
        torch.cuda.synchronize()
    b = buf.cpu().view(3, 16, 16).double()
    p1 = (b[2, 1:6, 2] - b[2, 1:6, 1]).mean(); p2 = (b[2, 1:6, 3] - b[2, 1:6, 2]).mean()
    mma = (b[1, 1:6, 14] - b[1, 1:6, 13]).mean(); tile = (b[2, 2:6, 3] - b[2, 1:5, 3]).mean()
    seg = lambda a, c: float((b[2, 1:6, a] - b[2, 1:6, c]).mean())
    slab = [float((b[1, 1:6, k * 3 + 2] - b[1, 1:6, k * 3 + 1]).mean()) for k in range(4)]
    swait = [float((b[1, 1:6, k * 3 + 1] - b[1, 1:6, k * 3 + 0]).mean()) for k in range(4)]
    print(f"{name:18s} slab issue {[round(v) for v in slab]} slab wait {[round(v) for v in swait]} tempty wait {float((b[1,1:6,13]-b[1,1:6,12]).mean()):.0f}")
    print(f"{name:18s} pass1={p1:7.0f} pass2={p2:7.0f} mma_tile={mma:7.0f} tile_period={tile:7.0f} | top->wait {seg(0, 7):6.0f} wait {seg(1, 0):6.0f} "
          f"loop {seg(4, 2):6.0f} shfl {seg(5, 4):6.0f} bar9 {seg(6, 5):6.0f} flush {seg(3, 6):6.0f}")
