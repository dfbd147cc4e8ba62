"""GPU diagnostic: per-trajectory-entry diff of the native sampler vs the CPU oracle for sampler variants (small net)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import torch_oracle as O
from oracle.make_golden import CASES
from diamond_b200.models.diffusion import Denoiser, DenoiserConfig, InnerModelConfig, DiffusionSampler, DiffusionSamplerConfig

dev = torch.device("cuda:0")
c = CASES["denoiser_small_heun"]; inner = c["inner"]
sd = O.seeded_state_dict(O.inner_model_shapes(inner), c["wseed"])
den = Denoiser(DenoiserConfig(InnerModelConfig(inner.img_channels, inner.num_steps_conditioning, inner.cond_channels, inner.depths, inner.channels, inner.attn_depths, inner.num_actions), 0.5, 0.3))
den.inner_model.load_state_dict(sd); den = den.to(dev).eval()
obs, act, _ = O.synthetic_inputs(c["b"], inner, c["h"], c["w"], c["iseed"])
g = np.load("tests/golden/denoiser_small_heun.npz")
x0 = torch.from_numpy(g["x0"]); eps = [torch.from_numpy(e) for e in g["eps"]]
for order, churn in ((1, 0.0), (1, 1.0), (2, 0.0), (2, 1.0)):
    sc = O.SamplerCfg(num_steps_denoising=4, order=order, s_churn=churn)
    with torch.no_grad():
        rx, rtraj = O.sample(obs, act, x0, sd, O.DenoiserCfg(inner=inner), sc, eps)
    s = DiffusionSampler(den, DiffusionSamplerConfig(4, order=order, s_churn=churn)); s.use_cuda_graph = False
    draws = [x0.to(dev)] + ([e.to(dev) for e in eps] if churn > 0 else [])
    orig = torch.randn; torch.randn = lambda *a, **k: draws.pop(0).clone()
    try:
        x, traj = s.sample(obs.to(dev), act.to(dev))
    finally:
        torch.randn = orig
    for i, (a, b) in enumerate(zip(traj, rtraj)):
        d = (a.cpu() - b).abs()
        print(f"order={order} churn={churn} traj[{i}] max={float(d.max()):.3e} frac>1e-3={float((d>1e-3).float().mean()):.3e} frac>1e-2={float((d>1e-2).float().mean()):.3e}")
