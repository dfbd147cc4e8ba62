"""Denoiser (reference: src/models/diffusion/denoiser.py): EDM preconditioning around the native InnerModel."""
from dataclasses import dataclass
from typing import Any, Dict, Tuple

import torch
from torch import Tensor
import torch.nn as nn

from ... import _lib
from .inner_model import InnerModel, InnerModelConfig

LossAndLogs = Tuple[Tensor, Dict[str, Any]]  # utils.py:53


def add_dims(input: Tensor, n: int) -> Tensor:  # denoiser.py:14-15
    return input.reshape(input.shape + (1,) * (n - input.ndim))


@dataclass
class Conditioners:  # denoiser.py:18-23
    c_in: Tensor
    c_out: Tensor
    c_skip: Tensor
    c_noise: Tensor


@dataclass
class SigmaDistributionConfig:  # denoiser.py:26-31
    loc: float
    scale: float
    sigma_min: float
    sigma_max: float


@dataclass
class DenoiserConfig:  # denoiser.py:34-38
    inner_model: InnerModelConfig
    sigma_data: float
    sigma_offset_noise: float


class Denoiser(nn.Module):
    def __init__(self, cfg: DenoiserConfig) -> None:
        super().__init__()
        self.cfg = cfg
        self.inner_model = InnerModel(cfg.inner_model)
        self.sample_sigma_training = None

    @property
    def device(self) -> torch.device:  # denoiser.py:48-50
        return self.inner_model.noise_emb.weight.device

    def setup_training(self, cfg: SigmaDistributionConfig) -> None:  # denoiser.py:52-59
        assert self.sample_sigma_training is None

        def sample_sigma(n: int, device: torch.device):
            s = torch.randn(n, device=device) * cfg.scale + cfg.loc
            return s.exp().clip(cfg.sigma_min, cfg.sigma_max)

        self.sample_sigma_training = sample_sigma

    def apply_noise(self, x: Tensor, sigma: Tensor, sigma_offset_noise: float) -> Tensor:  # denoiser.py:61-64
        b, c, _, _ = x.shape
        offset_noise = sigma_offset_noise * torch.randn(b, c, 1, 1, device=self.device)
        return x + offset_noise + torch.randn_like(x) * add_dims(sigma, x.ndim)

    def compute_conditioners(self, sigma: Tensor) -> Conditioners:  # denoiser.py:66-72 (host-side view; the fused path
        # recomputes the same fp32 expressions on device, csrc/aux_kernels.cuh edm_conditioners)
        sigma = (sigma**2 + self.cfg.sigma_offset_noise**2).sqrt()
        c_in = 1 / (sigma**2 + self.cfg.sigma_data**2).sqrt()
        c_skip = self.cfg.sigma_data**2 / (sigma**2 + self.cfg.sigma_data**2)
        c_out = sigma * c_skip.sqrt()
        c_noise = sigma.log() / 4
        return Conditioners(*(add_dims(c, n) for c, n in zip((c_in, c_out, c_skip, c_noise), (4, 4, 4, 1, 1))))

    def compute_model_output(self, noisy_next_obs: Tensor, obs: Tensor, act: Tensor, cs: Conditioners) -> Tensor:
        # denoiser.py:74-77
        rescaled_obs = obs / self.cfg.sigma_data
        rescaled_noise = noisy_next_obs * cs.c_in
        return self.inner_model(rescaled_noise, cs.c_noise, rescaled_obs, act)

    def _native_forward(self, noisy: Tensor, sigma: Tensor, obs: Tensor, act: Tensor, want_model: bool, want_denoised: bool):
        lib = _lib.lib()
        im = self.inner_model
        h = im.native(self.cfg.sigma_data, self.cfg.sigma_offset_noise)
        b, _, hh, ww = noisy.shape
        noisy_, obs_ = noisy.float().contiguous(), obs.float().contiguous()
        sig = sigma.float().contiguous().reshape(-1).to(noisy.device)
        if sig.numel() not in (1, b):
            raise ValueError("sigma must have 1 or B elements")
        act_ = act.long().contiguous()
        model = torch.empty_like(noisy_) if want_model else None
        den = torch.empty_like(noisy_) if want_denoised else None
        ws = im.workspace(lib.dmd_denoiser_workspace_bytes(h, b, hh, ww))
        _lib.check(lib.dmd_denoiser_forward(h, b, hh, ww, noisy_.data_ptr(), sig.data_ptr(), int(sig.numel() == 1),
                                            obs_.data_ptr(), act_.data_ptr(), _lib.ptr(model), _lib.ptr(den),
                                            ws.data_ptr(), ws.numel(), _lib.current_stream()))
        return model, den

    @torch.no_grad()
    def wrap_model_output(self, noisy_next_obs: Tensor, model_output: Tensor, cs: Conditioners) -> Tensor:
        # denoiser.py:79-84; elementwise torch ops on CUDA tensors (the fused path does this inside wrap_update_kernel)
        d = cs.c_skip * noisy_next_obs + cs.c_out * model_output
        return d.clamp(-1, 1).add(1).div(2).mul(255).byte().div(255).mul(2).sub(1)

    @torch.no_grad()
    def denoise(self, noisy_next_obs: Tensor, sigma: Tensor, obs: Tensor, act: Tensor) -> Tensor:  # denoiser.py:86-91
        _, den = self._native_forward(noisy_next_obs, sigma, obs, act, want_model=False, want_denoised=True)
        return den

    def forward(self, batch) -> LossAndLogs:  # denoiser.py:93-122
        """Training loss: for each autoregressive step, noise the target frame, run the native U-Net (autograd node
        `_InnerModelFn`), regress the EDM target, and write the quantised denoised frame back so that the next step is
        conditioned on the model's own output.  Same RNG draws, in the same order, as the reference (sigma, offset, noise)."""
        if self.sample_sigma_training is None:
            raise RuntimeError("call setup_training(SigmaDistributionConfig) first (denoiser.py:52)")
        n_cond = self.cfg.inner_model.num_steps_conditioning
        frames = batch.obs.clone()                     # (B, T, C, H, W); column n_cond + i is overwritten by step i
        steps = frames.size(1) - n_cond
        b, _, c, h, w = frames.shape
        total = 0
        for i in range(steps):
            target_frame = frames[:, n_cond + i]
            keep = batch.mask_padding[:, n_cond + i]
            stack = frames[:, i:n_cond + i].reshape(b, n_cond * c, h, w)
            sigma = self.sample_sigma_training(b, self.device)
            noisy = self.apply_noise(target_frame, sigma, self.cfg.sigma_offset_noise)
            cs = self.compute_conditioners(sigma)
            out = self.compute_model_output(noisy, stack, batch.act[:, i:n_cond + i], cs)
            wanted = (target_frame - cs.c_skip * noisy) / cs.c_out
            total = total + torch.nn.functional.mse_loss(out[keep], wanted[keep])
            frames[:, n_cond + i] = self.wrap_model_output(noisy, out, cs)
        loss = total / steps
        return loss, {"loss_denoising": loss.detach()}
