"""Synthetic weights / inputs for benchmarks and smoke runs (there is no network for datasets or checkpoints).

`randomize_module_` fills EVERY parameter and buffer of a module from a numpy PCG64 stream ("de-zeroed" weights: the
reference zero-initialises conv2 / out_proj / conv_out / the actor-critic heads, blocks.py:59-60,139, inner_model.py:42,
actor_critic.py:50-53 — with zeros half the network would contribute nothing to a benchmark).  The stream and the scaling
rules are keyed only by the state_dict order and names, so the same seed gives the same numbers on every box; the test
oracle uses the identical rule, which lets the GPU tests compare against it.
"""
import math

import numpy as np
import torch


def _scaled(name: str, shape, a: np.ndarray) -> np.ndarray:
    if name.endswith("norm.weight"):            # GroupNorm gamma ~ 1
        return 1.0 + 0.2 * a
    if name.endswith(".bias") or "bias" in name.rsplit(".", 1)[-1]:
        return 0.1 * a
    if name == "noise_emb.weight" or "act_emb" in name or name.endswith("emb.weight"):
        return a                                  # N(0, 1) like torch.randn / nn.Embedding
    if "norm1.linear.weight" in name or "norm2.linear.weight" in name:
        return a * (0.5 / math.sqrt(shape[1]))   # FiLM projections: small scale / shift
    fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
    return a / math.sqrt(fan_in)


@torch.no_grad()
def randomize_module_(module: torch.nn.Module, seed: int) -> torch.nn.Module:
    rng = np.random.default_rng(seed)
    for name, t in module.state_dict().items():
        shape = tuple(t.shape)
        a = _scaled(name, shape, rng.standard_normal(shape))
        t.copy_(torch.from_numpy(np.ascontiguousarray(a)).to(t.dtype))
    return module


def frame_stacks(b: int, t: int, c: int, h: int, w: int, num_actions: int, seed: int):
    """(obs, act, x0): frames on the 1/255 grid in [-1, 1] (what Episode.load produces, episode.py:36-43), random actions,
    and a standard-normal tensor shaped like one frame."""
    rng = np.random.default_rng(seed)
    obs = torch.from_numpy(rng.integers(0, 256, size=(b, t, c, h, w)).astype(np.float32)).div(255).mul(2).sub(1)
    act = torch.from_numpy(rng.integers(0, num_actions, size=(b, t)).astype(np.int64))
    x0 = torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32))
    return obs, act, x0
