"""Bring-up diagnostic: prep_act_kernel mode 1 (AdaGroupNorm + SiLU) output decoded from the PLC16 operand vs torch."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diamond_b200 import _lib, ops  # noqa: E402

if os.environ.get("DMD_LIB"):
    import ctypes
    _lib.LIB_PATH = os.environ["DMD_LIB"]
    h = ctypes.CDLL(_lib.LIB_PATH)
    _lib.SIGNATURES = {k: v for k, v in _lib.SIGNATURES.items() if hasattr(h, k) and k != "dmd_sampler_sample"}

dev = torch.device("cuda:0")
b, c, hh, ww = 2, 64, 64, 64
g = torch.Generator().manual_seed(3)
x = torch.randn(b, c, hh, ww, generator=g)
film = torch.randn(b, 2 * c + 5, generator=g) * 0.3
scale, shift = film[:, 5:5 + c], film[:, 5 + c:5 + 2 * c]
for mode, silu in ((0, False), (1, False), (1, True), (2, True)):
    gamma = 1 + 0.2 * torch.randn(c, generator=torch.Generator().manual_seed(9))
    beta = 0.1 * torch.randn(c, generator=torch.Generator().manual_seed(10))
    if mode == 0:
        ref = x
    elif mode == 1:
        ref = F.group_norm(x, 2, eps=1e-5) * (1 + scale[:, :, None, None]) + shift[:, :, None, None]
    else:
        ref = F.group_norm(x, 2, gamma, beta, eps=1e-5)
    if silu:
        ref = F.silu(ref)
    xs = ops.nchw_to_nhwc(x.to(dev))
    st = ops.gn_stats(xs, 32)
    out = ops.prep_act(xs, mode=mode, silu=silu, stats0=st if mode else None, gs0=32 if mode else 0, film=film.to(dev) if mode == 1 else None,
                       film_off=5, gamma=gamma.to(dev) if mode == 2 else None, beta=beta.to(dev) if mode == 2 else None)[0]
    torch.cuda.synchronize()
    pw, ph = ww + 1, hh + 1
    q = b * ph * pw
    G = pw + 1
    qalloc = G + ((q + 127) // 128) * 128 + pw + 1
    planes = out.view(torch.float16).reshape(c // 8, qalloc, 8)[:, G:G + q].reshape(c // 8, b, ph, pw, 8)
    got = planes[:, :, :hh, :ww].permute(1, 0, 4, 2, 3).reshape(b, c, hh, ww).float().cpu()
    err = (got - ref).abs()
    per_img = [float(err[i].max()) for i in range(b)]
    per_grp = [float(err[:, 32 * k:32 * (k + 1)].max()) for k in range(2)]
    pads = float(planes[:, :, hh, :].abs().max()), float(planes[:, :, :, ww].abs().max())
    print(f"mode={mode} silu={silu}: max err {float(err.max()):.3e} per image {per_img} per group {per_grp} pad max {pads}", flush=True)
    if float(err.max()) > 1e-2:
        idx = torch.nonzero(err > 1e-2)[:5]
        for i in idx:
            n, ch, y, xx = [int(v) for v in i]
            print(f"   [{n},{ch},{y},{xx}] got {float(got[n, ch, y, xx]):.4f} want {float(ref[n, ch, y, xx]):.4f} raw {float(x[n, ch, y, xx]):.4f}")
