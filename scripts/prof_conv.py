"""Runs the dominant conv pair (prep_act + conv_tc, 3x3 64->64 @64x64, 32 images, AdaGN+SiLU, residual, stats) — ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diamond_b200 import ops
dev = torch.device("cuda:0")
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 64
wt = (torch.randn(64, 64, 3, 3) / 24).to(dev)
wpk, cp = ops.pack_conv_weight(wt, 64)
bias = torch.zeros(64, device=dev)
xs = [torch.randn(envs, hw, hw, 64, device=dev) for _ in range(4)]
film = torch.randn(envs, 128, device=dev) * 0.1
sts = [ops.gn_stats(x, 32) for x in xs]
for i in range(8):
    ops.conv2d_fprop(xs[i % 4], wpk, 64, cp, 64, bias=bias, prologue=1, silu=True, stats0=sts[i % 4], gs0=32, film=film, out_gs=32, residual=xs[(i + 1) % 4])
torch.cuda.synchronize()
print("done")
