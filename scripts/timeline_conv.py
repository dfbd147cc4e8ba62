"""GPU: clock64 timeline of CTA 0's roles for one conv launch (bring-up instrumentation)."""
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
for name, kw in (("skeleton", dict(prologue=0, silu=False, out_gs=0, debug=14 + 224)), ("plain", dict(prologue=0, silu=False, out_gs=0)),
                 ("full", dict(prologue=1, silu=True, out_gs=32, stats0=st, gs0=32, film=film))):
    for rep in range(2):
        buf = torch.zeros(3 * 16 * 16, dtype=torch.int64, device=dev)
        ops.conv2d_fprop(x, wpk, 64, cp, 64, bias=bias, debug_buf=buf, **kw)
        torch.cuda.synchronize()
    b = buf.cpu().view(3, 16, 16)
    t0 = int(b[b > 0].min())
    print(f"==== {name}: cycles relative to first stamp (CTA 0, 7-8 tiles)")
    for it in range(8):
        L = [int(v) - t0 if v > 0 else -1 for v in b[0, it, :12]]
        M = [int(v) - t0 if v > 0 else -1 for v in b[1, it, :15]]
        E = [int(v) - t0 if v > 0 else -1 for v in b[2, it, :3]]
        print(f"tile {it}: LOAD(w_empty,got,arrived)x4 {L}")
        print(f"         MMA(w_full,got,committed)x4 {M[:12]}  tempty(wait,got) {M[12:14]} tfull_commit {M[14]}")
        print(f"         EPI(wait_tfull,got,arrive_tempty) {E}")
