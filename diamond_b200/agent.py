"""Agent container (reference: src/agent.py:28-62): owns the three models under the reference's attribute names.

`rew_end_model` is NOT part of this hot path: it stays the reference's
`models.rew_end_model.RewEndModel`, resolved at construction time from the overlaid reference tree."""
from collections import OrderedDict
from dataclasses import dataclass
from pathlib import Path
from typing import Any

import torch
import torch.nn as nn

from .models.actor_critic import ActorCritic, ActorCriticConfig, ActorCriticLossConfig
from .models.diffusion import Denoiser, DenoiserConfig, SigmaDistributionConfig


@dataclass
class AgentConfig:  # agent.py:16-25
    denoiser: DenoiserConfig
    rew_end_model: Any
    actor_critic: ActorCriticConfig
    num_actions: int

    def __post_init__(self) -> None:
        self.denoiser.inner_model.num_actions = self.num_actions
        self.rew_end_model.num_actions = self.num_actions
        self.actor_critic.num_actions = self.num_actions


def _reference_rew_end_model():
    try:
        from models.rew_end_model import RewEndModel  # the reference's module, on sys.path in an overlaid tree
    except Exception as e:  # pragma: no cover - depends on the deployment
        raise RuntimeError("Agent needs the reference's models.rew_end_model.RewEndModel on sys.path (INTEGRATION.md)") from e
    return RewEndModel


class Agent(nn.Module):
    def __init__(self, cfg: AgentConfig, rew_end_model_cls=None) -> None:
        super().__init__()
        self.denoiser = Denoiser(cfg.denoiser)
        self.rew_end_model = (rew_end_model_cls or _reference_rew_end_model())(cfg.rew_end_model)
        self.actor_critic = ActorCritic(cfg.actor_critic)

    @property
    def device(self):
        return self.denoiser.device

    def setup_training(self, sigma_distribution_cfg: SigmaDistributionConfig, actor_critic_loss_cfg: ActorCriticLossConfig, rl_env) -> None:
        self.denoiser.setup_training(sigma_distribution_cfg)
        self.actor_critic.setup_training(rl_env, actor_critic_loss_cfg)

    def load(self, path_to_ckpt: Path, load_denoiser: bool = True, load_rew_end_model: bool = True, load_actor_critic: bool = True) -> None:
        sd = torch.load(Path(path_to_ckpt), map_location=self.device)
        parts = {name: OrderedDict((k.split(".", 1)[1], v) for k, v in sd.items() if k.startswith(name))
                 for name in ("denoiser", "rew_end_model", "actor_critic")}  # utils.extract_state_dict (utils.py:173-174)
        if load_denoiser:
            self.denoiser.load_state_dict(parts["denoiser"])
        if load_rew_end_model:
            self.rew_end_model.load_state_dict(parts["rew_end_model"])
        if load_actor_critic:
            self.actor_critic.load_state_dict(parts["actor_critic"])
