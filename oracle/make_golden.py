"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference/src with test-side
stub modules, oracle/ref_import.py) on seeded inputs and seeded 'de-zeroed' weights.  Run in the build container:

    python oracle/make_golden.py

The fixtures pin oracle/torch_oracle.py (CPU tests) and the CUDA path (GPU tests) to the reference itself.
Weights are NOT stored (17 MB); they are regenerated from the seed by torch_oracle.seeded_state_dict and guarded by a
checksum stored in the fixture.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # SURVEY.md 8d cfg 1 shape (default agent config), batch 2 so per-sample sigma / cond paths are exercised
    "denoiser_default": dict(
        inner=O.InnerCfg(), h=64, w=64, b=2, wseed=1234, iseed=77, sigmas=[0.7, 3.0],
        sampler=O.SamplerCfg(num_steps_denoising=3), rng_seed=0,
    ),
    # small net: 3 levels, attention inside a level, Heun + churn, odd batch, 6 actions, 2 conditioning frames
    "denoiser_small_heun": dict(
        inner=O.InnerCfg(img_channels=3, num_steps_conditioning=2, cond_channels=64, depths=[1, 2, 1],
                         channels=[32, 64, 32], attn_depths=[0, 0, 1], num_actions=6),
        h=32, w=32, b=3, wseed=4321, iseed=78, sigmas=[0.05, 1.0, 12.0],
        sampler=O.SamplerCfg(num_steps_denoising=4, order=2, s_churn=1.0), rng_seed=5,
    ),
}


def build_reference(ns, inner: O.InnerCfg, sd):
    D = ns.diffusion
    cfg = D.DenoiserConfig(
        D.InnerModelConfig(inner.img_channels, inner.num_steps_conditioning, inner.cond_channels, list(inner.depths),
                           list(inner.channels), list(inner.attn_depths), num_actions=inner.num_actions),
        sigma_data=0.5, sigma_offset_noise=0.3)
    den = D.Denoiser(cfg)
    ref_keys = list(den.inner_model.state_dict().keys())
    assert ref_keys == list(sd.keys()), "oracle key order differs from the reference's state_dict"
    for k, v in den.inner_model.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
    den.inner_model.load_state_dict(sd)
    return den.eval()


def main():
    ns = ref_import.load()
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    for name, c in CASES.items():
        inner = c["inner"]
        sd = O.seeded_state_dict(O.inner_model_shapes(inner), c["wseed"])
        den = build_reference(ns, inner, sd)
        obs, act, x_noisy = O.synthetic_inputs(c["b"], inner, c["h"], c["w"], c["iseed"])
        sig = torch.tensor(c["sigmas"], dtype=torch.float32)
        b, t, ch, h, w = obs.shape
        obs_flat = obs.reshape(b, t * ch, h, w)
        with torch.no_grad():
            cs = den.compute_conditioners(sig)
            mo = den.compute_model_output(x_noisy, obs_flat, act, cs)
            dn = den.wrap_model_output(x_noisy, mo, cs)
            dn2 = den.denoise(x_noisy, sig, obs_flat, act)
            assert torch.equal(dn, dn2)
            # full sampler, reference draws its own noise from the global torch RNG
            s = c["sampler"]
            sampler = ns.diffusion.DiffusionSampler(den, ns.diffusion.DiffusionSamplerConfig(
                s.num_steps_denoising, s.sigma_min, s.sigma_max, s.rho, s.order, s.s_churn, s.s_tmin, s.s_tmax, s.s_noise))
            torch.manual_seed(c["rng_seed"])
            x, traj = sampler.sample(obs, act)
            # replay the RNG stream to capture the noise tensors the reference consumed (diffusion_sampler.py:36,42)
            torch.manual_seed(c["rng_seed"])
            x0 = torch.randn(b, ch, h, w)
            assert torch.equal(x0, traj[0])
            n_sig = len(sampler.sigmas)
            gamma_ = min(s.s_churn / (n_sig - 1), 2**0.5 - 1)
            eps = np.zeros((n_sig - 1, b, ch, h, w), np.float32)
            for i, sg in enumerate(sampler.sigmas[:-1]):
                if gamma_ > 0 and s.s_tmin <= sg <= s.s_tmax:
                    eps[i] = (torch.randn(b, ch, h, w) * s.s_noise).numpy()
                    if s.order == 2 and sampler.sigmas[i + 1] != 0:
                        pass  # Heun draws no extra noise
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            weights_checksum=np.float64(O.state_checksum(sd)),
            sigmas_in=sig.numpy(), model_output=mo.numpy(), denoised=dn.numpy(),
            sampler_sigmas=sampler.sigmas.numpy(), x0=x0.numpy(), eps=eps, sample_x=x.numpy(),
            trajectory=torch.stack(traj).numpy(),
        )
        print(name, "model_output rms", float(mo.pow(2).mean().sqrt()), "sample rms", float(x.pow(2).mean().sqrt()),
              "size", os.path.getsize(os.path.join(OUT, name + ".npz")))


def make_actor_critic():
    """Reference ActorCritic.predict_act_value over 3 recurrent steps (hidden state carried), seeded de-zeroed weights."""
    ns = ref_import.load()
    AC = ns.actor_critic
    cfg = O.ActorCriticCfg()
    sd = O.seeded_actor_critic_state_dict(cfg, 555)
    ac = AC.ActorCritic(AC.ActorCriticConfig(cfg.lstm_dim, cfg.img_channels, cfg.img_size, list(cfg.channels), list(cfg.down), cfg.num_actions)).eval()
    assert [(k, tuple(v.shape)) for k, v in ac.state_dict().items()] == O.actor_critic_shapes(cfg)
    ac.load_state_dict(sd)
    rng = np.random.default_rng(91)
    b = 5
    obs = torch.from_numpy(rng.integers(0, 256, size=(3, b, 3, 64, 64)).astype(np.float32)).div(255).mul(2).sub(1)
    hx = torch.from_numpy(rng.standard_normal((b, 512)).astype(np.float32)) * 0.3
    cx = torch.from_numpy(rng.standard_normal((b, 512)).astype(np.float32)) * 0.3
    logits, vals = [], []
    h, c = hx, cx
    with torch.no_grad():
        for t in range(3):
            out = ac.predict_act_value(obs[t], (h, c))
            logits.append(out.logits_act); vals.append(out.val); h, c = out.hx_cx
    np.savez_compressed(os.path.join(OUT, "actor_critic_default.npz"), weights_checksum=np.float64(O.state_checksum(sd)),
                        hx0=hx.numpy(), cx0=cx.numpy(), logits=torch.stack(logits).numpy(), val=torch.stack(vals).numpy(),
                        hx=h.numpy(), cx=c.numpy())
    print("actor_critic_default logits rms", float(torch.stack(logits).pow(2).mean().sqrt()))


if __name__ == "__main__":
    main()
    make_actor_critic()
