"""ActorCritic (reference: src/models/actor_critic.py) with a native sm_100a forward for predict_act_value."""
import ctypes as C
import math
from collections import namedtuple
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor
from torch.distributions.categorical import Categorical

from .. import _lib
from ..coroutines.env_loop import make_env_loop
from ..utils import LossAndLogs, init_lstm
from .blocks import SmallResBlock, _NativeOnly, conv3x3

ActorCriticOutput = namedtuple("ActorCriticOutput", "logits_act val hx_cx")


@dataclass
class ActorCriticLossConfig:  # actor_critic.py:21-27
    backup_every: int
    gamma: float
    lambda_: float
    weight_value_loss: float
    weight_entropy_loss: float


@dataclass
class ActorCriticConfig:  # actor_critic.py:30-37
    lstm_dim: int
    img_channels: int
    img_size: int
    channels: List[int]
    down: List[int]
    num_actions: Optional[int] = None


class ActorCriticEncoder(_NativeOnly):  # actor_critic.py:101-113 (parameter container; executed natively)
    def __init__(self, cfg: ActorCriticConfig) -> None:
        super().__init__()
        assert len(cfg.channels) == len(cfg.down)
        layers = [conv3x3(cfg.img_channels, cfg.channels[0])]
        for i in range(len(cfg.channels)):
            layers.append(SmallResBlock(cfg.channels[max(0, i - 1)], cfg.channels[i]))
            if cfg.down[i]:
                layers.append(nn.MaxPool2d(2))
        self.encoder = nn.Sequential(*layers)


class ActorCritic(nn.Module):
    def __init__(self, cfg: ActorCriticConfig) -> None:
        super().__init__()
        self.cfg = cfg
        self.encoder = ActorCriticEncoder(cfg)
        self.lstm_dim = cfg.lstm_dim
        input_dim_lstm = cfg.channels[-1] * (cfg.img_size // 2 ** (sum(cfg.down))) ** 2
        self.lstm = nn.LSTMCell(input_dim_lstm, cfg.lstm_dim)
        self.critic_linear = nn.Linear(cfg.lstm_dim, 1)
        self.actor_linear = nn.Linear(cfg.lstm_dim, cfg.num_actions)
        for lin in (self.actor_linear, self.critic_linear):  # actor_critic.py:50-53
            lin.weight.data.fill_(0)
            lin.bias.data.fill_(0)
        init_lstm(self.lstm)
        self.env_loop = None
        self.loss_cfg = None
        self._h = None
        self._wkey = None
        self._packed = None
        self._ws = None

    def __del__(self):
        try:
            if self._h is not None:
                _lib.lib().dmd_actor_critic_destroy(self._h)
        except Exception:
            pass

    @property
    def device(self) -> torch.device:  # actor_critic.py:59-61
        return self.lstm.weight_hh.device

    def setup_training(self, rl_env, loss_cfg: ActorCriticLossConfig) -> None:  # actor_critic.py:63-66
        assert self.env_loop is None and self.loss_cfg is None
        self.env_loop = make_env_loop(rl_env, self)
        self.loss_cfg = loss_cfg

    # ------------------------------------------------------------------ native plumbing
    def _native(self):
        lib = _lib.lib()
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("diamond_b200 runs on CUDA (sm_100a) only; move the model to a cuda device")
        if self._h is None:
            c = self.cfg
            cc = _lib.ActorCriticConfigC()
            cc.lstm_dim, cc.img_channels, cc.img_size, cc.num_levels = c.lstm_dim, c.img_channels, c.img_size, len(c.channels)
            for i in range(len(c.channels)):
                cc.channels[i], cc.down[i] = int(c.channels[i]), int(bool(c.down[i]))
            cc.num_actions = int(c.num_actions)
            h = lib.dmd_actor_critic_create(C.byref(cc))
            if not h:
                raise RuntimeError("diamond_b200: " + lib.dmd_last_error().decode())
            self._h = h
        tensors = list(self.state_dict(keep_vars=True).values())
        wkey = tuple((t.data_ptr(), t._version) for t in tensors)
        if wkey != self._wkey:
            n = lib.dmd_actor_critic_num_tensors(self._h)
            if n != len(tensors):
                raise RuntimeError(f"native actor-critic expects {n} tensors, module has {len(tensors)}")
            if self._packed is None:
                self._packed = torch.empty(lib.dmd_actor_critic_packed_bytes(self._h), dtype=torch.uint8, device=dev)
            arr = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
            _lib.check(lib.dmd_actor_critic_set_weights(self._h, arr, n, self._packed.data_ptr(), _lib.current_stream()))
            self._wkey = wkey
        return self._h

    # ------------------------------------------------------------------ reference surface
    def predict_act_value(self, obs: Tensor, hx_cx: Tuple[Tensor, Tensor]) -> ActorCriticOutput:  # actor_critic.py:68-73
        assert obs.ndim == 4
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "ActorCritic.predict_act_value with autograd (BPTT through the imagined rollout, SURVEY.md 8 a22/a23) is not "
                "built yet; call under torch.no_grad()"
            )
        lib = _lib.lib()
        h = self._native()
        b = obs.size(0)
        hx, cx = hx_cx
        obs_, hx_, cx_ = obs.float().contiguous(), hx.float().contiguous(), cx.float().contiguous()
        logits = torch.empty(b, self.cfg.num_actions, device=obs.device)
        val = torch.empty(b, device=obs.device)
        hx_o, cx_o = torch.empty_like(hx_), torch.empty_like(cx_)
        need = lib.dmd_actor_critic_workspace_bytes(h, b)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=obs.device)
        _lib.check(lib.dmd_actor_critic_forward(h, b, obs_.data_ptr(), hx_.data_ptr(), cx_.data_ptr(), logits.data_ptr(),
                                                val.data_ptr(), hx_o.data_ptr(), cx_o.data_ptr(), self._ws.data_ptr(),
                                                self._ws.numel(), _lib.current_stream()))
        return ActorCriticOutput(logits, val, (hx_o, cx_o))

    def forward(self) -> LossAndLogs:  # actor_critic.py:75-98
        c = self.loss_cfg
        _, act, rew, end, trunc, logits_act, val, val_bootstrap, _ = self.env_loop.send(c.backup_every)
        d = Categorical(logits=logits_act)
        entropy = d.entropy().mean()
        lambda_returns = compute_lambda_returns(rew, end, trunc, val_bootstrap, c.gamma, c.lambda_)
        loss_actions = (-d.log_prob(act) * (lambda_returns - val).detach()).mean()
        loss_values = c.weight_value_loss * F.mse_loss(val, lambda_returns)
        loss_entropy = -c.weight_entropy_loss * entropy
        loss = loss_actions + loss_entropy + loss_values
        metrics = {
            "policy_entropy": entropy.detach() / math.log(2),
            "loss_actions": loss_actions.detach(),
            "loss_entropy": loss_entropy.detach(),
            "loss_values": loss_values.detach(),
            "loss_total": loss.detach(),
        }
        return loss, metrics


@torch.no_grad()
def compute_lambda_returns(rew: Tensor, end: Tensor, trunc: Tensor, val_bootstrap: Tensor, gamma: float, lambda_: float) -> Tensor:
    """actor_critic.py:116-143: lambda-returns with sign-clipped rewards, episode ends and truncations."""
    assert rew.ndim == 2 and rew.size() == end.size() == trunc.size() == val_bootstrap.size()
    rew = rew.sign()
    end_or_trunc = (end + trunc).clip(max=1)
    not_end, not_trunc = 1 - end, 1 - trunc
    returns = rew + not_end * gamma * (not_trunc * (1 - lambda_) + trunc) * val_bootstrap
    if lambda_ == 0:
        return returns
    last = val_bootstrap[:, -1]
    for t in reversed(range(rew.size(1))):
        returns[:, t] += end_or_trunc[:, t].logical_not() * gamma * lambda_ * last
        last = returns[:, t]
    return returns
