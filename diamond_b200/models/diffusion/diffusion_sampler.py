"""DiffusionSampler (reference: src/models/diffusion/diffusion_sampler.py), whole loop in one native call."""
import ctypes as C
from dataclasses import dataclass
from typing import List, Tuple

import torch
from torch import Tensor

from ... import _lib
from .denoiser import Denoiser


@dataclass
class DiffusionSamplerConfig:  # diffusion_sampler.py:10-20
    num_steps_denoising: int
    sigma_min: float = 2e-3
    sigma_max: float = 5
    rho: int = 7
    order: int = 1
    s_churn: float = 0
    s_tmin: float = 0
    s_tmax: float = float("inf")
    s_noise: float = 1


def build_sigmas(num_steps: int, sigma_min: float, sigma_max: float, rho: int, device: torch.device) -> Tensor:
    # diffusion_sampler.py:61-66 (Karras et al. schedule), same fp32 torch expressions
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    l = torch.linspace(0, 1, num_steps, device=device)
    sigmas = (max_inv_rho + l * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat((sigmas, sigmas.new_zeros(1)))


class DiffusionSampler:
    def __init__(self, denoiser: Denoiser, cfg: DiffusionSamplerConfig) -> None:
        self.denoiser = denoiser
        self.cfg = cfg
        # The schedule is host-known: computing it on CPU (bit-identical to the reference's CPU path) removes the two
        # device->host syncs per step the reference pays for `sigma <= ...` / `next_sigma == 0` (diffusion_sampler.py:39,47).
        self._sigmas_host = build_sigmas(cfg.num_steps_denoising, cfg.sigma_min, cfg.sigma_max, cfg.rho, torch.device("cpu"))
        self.sigmas = self._sigmas_host.to(denoiser.device)
        self.use_cuda_graph = True

    @torch.no_grad()
    def sample(self, prev_obs: Tensor, prev_act: Tensor) -> Tuple[Tensor, List[Tensor]]:  # diffusion_sampler.py:31-58
        lib = _lib.lib()
        den = self.denoiser
        im = den.inner_model
        device = prev_obs.device
        b, t, c, h, w = prev_obs.size()
        obs = prev_obs.reshape(b, t * c, h, w).float().contiguous()
        act = prev_act.long().contiguous()
        n_sig = int(self._sigmas_host.numel())
        gamma_ = min(self.cfg.s_churn / (n_sig - 1), 2**0.5 - 1)
        # RNG stream parity with the reference: x first, then one eps per step that churns (diffusion_sampler.py:36,42)
        x0 = torch.randn(b, c, h, w, device=device)
        eps = None
        if gamma_ > 0:
            eps = torch.zeros(n_sig - 1, b, c, h, w, device=device)
            for i, sigma in enumerate(self._sigmas_host[:-1].tolist()):
                if self.cfg.s_tmin <= sigma <= self.cfg.s_tmax:
                    eps[i] = torch.randn(b, c, h, w, device=device) * self.cfg.s_noise
        hnd = im.native(den.cfg.sigma_data, den.cfg.sigma_offset_noise)
        core = lib.dmd_denoiser_workspace_bytes(hnd, b, h, w)
        img_bytes = b * c * h * w * 4
        need = core + img_bytes * (n_sig + (n_sig - 1 if eps is not None else 0)) + 4096
        ws = im.workspace(need)
        sig_arr = (C.c_float * n_sig)(*self._sigmas_host.tolist())
        sc = _lib.SamplerConfigC(n_sig, sig_arr, int(self.cfg.order), float(self.cfg.s_churn), float(self.cfg.s_tmin),
                                 float(min(self.cfg.s_tmax, 3.0e38)), 1.0)
        out_x = torch.empty(b, c, h, w, device=device)
        traj = torch.empty(n_sig, b, c, h, w, device=device)
        _lib.check(lib.dmd_sampler_sample(hnd, C.byref(sc), b, h, w, obs.data_ptr(), act.data_ptr(), x0.data_ptr(),
                                          _lib.ptr(eps), out_x.data_ptr(), traj.data_ptr(), ws.data_ptr(), ws.numel(),
                                          int(self.use_cuda_graph), _lib.current_stream()))
        return out_x, list(traj.unbind(0))
