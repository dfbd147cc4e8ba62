"""GPU: native RewEndModel.predict_rew_end against the reference's own outputs (tests/golden/rew_end_default.npz,
written by the unmodified reference: a 3-step burn-in call that returns the LSTM state, then two single-step calls carrying it —
the way WorldModelEnv uses the model, world_model_env.py:96-105,120-129)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-12))


def test_native_rew_end_matches_reference_golden(golden_dir):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    dev = torch.device("cuda:0")
    from diamond_b200.models.rew_end_model import RewEndModel, RewEndModelConfig
    from oracle import torch_oracle as O

    g = np.load(os.path.join(golden_dir, "rew_end_default.npz"))
    cfg = O.RewEndCfg()
    sd = O.seeded_state_dict(O.rew_end_shapes(cfg), 777)
    assert abs(O.state_checksum(sd) - float(g["weights_checksum"])) < 1e-6 * float(g["weights_checksum"])
    m = RewEndModel(RewEndModelConfig(cfg.lstm_dim, cfg.img_channels, cfg.img_size, cfg.cond_channels, list(cfg.depths), list(cfg.channels),
                                      list(cfg.attn_depths), cfg.num_actions))
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    frames, act = torch.from_numpy(g["frames"]).to(dev), torch.from_numpy(g["act"]).to(dev)
    lr, le, hc = m.predict_rew_end(frames[:, 0:3], act[:, 0:3], frames[:, 1:4])
    e = [_rel(lr.cpu(), torch.from_numpy(g["burn_rew"])), _rel(le.cpu(), torch.from_numpy(g["burn_end"]))]
    for k in (3, 4):
        lr, le, hc = m.predict_rew_end(frames[:, k:k + 1], act[:, k:k + 1], frames[:, k + 1:k + 2], hc)
        e += [_rel(lr.cpu(), torch.from_numpy(g[f"step{k}_rew"])), _rel(le.cpu(), torch.from_numpy(g[f"step{k}_end"]))]
    e += [_rel(hc[0].cpu(), torch.from_numpy(g["hx"])), _rel(hc[1].cpu(), torch.from_numpy(g["cx"]))]
    print("rew_end rel errors (burn rew/end, step3 rew/end, step4 rew/end, hx, cx):", ["%.2e" % v for v in e])
    assert hc[0].shape == (1, 3, cfg.lstm_dim)
    assert max(e) < 2e-3, e   # logits are small-magnitude sums of 512 terms; hidden state within 1e-3
    assert max(e[-2:]) < 1e-3, e
