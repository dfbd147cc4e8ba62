"""CPU, build container only: the oracle restatement against the LIVE unmodified reference on fresh seeds."""
import pytest
import torch

from oracle import ref_import
from oracle import torch_oracle as O

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree absent (GPU box)")


def test_blocks_and_denoise_match_live_reference():
    torch.set_num_threads(8)
    ns = ref_import.load()
    D = ns.diffusion
    inner = O.InnerCfg(depths=[1, 1, 1], channels=[32, 32, 64], attn_depths=[0, 1, 0], cond_channels=64, num_actions=5)
    sd = O.seeded_state_dict(O.inner_model_shapes(inner), 31337)
    den = D.Denoiser(D.DenoiserConfig(D.InnerModelConfig(3, 4, 64, [1, 1, 1], [32, 32, 64], [0, 1, 0], 5), 0.5, 0.3)).eval()
    den.inner_model.load_state_dict(sd)
    obs, act, x = O.synthetic_inputs(3, inner, 32, 32, 9)
    sig = torch.tensor([0.01, 0.5, 7.0])
    cfg = O.DenoiserCfg(inner=inner)
    with torch.no_grad():
        want = den.denoise(x, sig, obs.reshape(3, 12, 32, 32), act)
        got = O.denoise(x, sig, obs.reshape(3, 12, 32, 32), act, sd, cfg)
        cs = den.compute_conditioners(sig)
        want_mo = den.compute_model_output(x, obs.reshape(3, 12, 32, 32), act, cs)
        got_mo = O.model_output(x, sig, obs.reshape(3, 12, 32, 32), act, sd, cfg)
    assert torch.allclose(got_mo, want_mo, rtol=1e-5, atol=1e-5)
    assert float((got != want).float().mean()) < 1e-3
