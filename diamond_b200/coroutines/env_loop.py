"""Policy/environment stepping coroutine (reference: src/coroutines/env_loop.py:12-74).

Host logic only; the arithmetic lives in `model.predict_act_value` (native actor-critic) and `env.step` (native
sampler).  Semantics kept exactly: bootstrap values of dead envs come from the final observation, hidden state of dead
envs is zeroed and burnt in on the fresh segment's frames, `.send(num_steps)` returns stacked (B, T, ...) tensors."""
import random
from typing import Generator, Tuple

import torch
import torch.nn as nn
from torch.distributions.categorical import Categorical

from . import coroutine


@coroutine
def make_env_loop(env, model: nn.Module, epsilon: float = 0.0) -> Generator[Tuple[torch.Tensor, ...], int, None]:
    num_steps = yield

    hx = torch.zeros(env.num_envs, model.lstm_dim, device=model.device)
    cx = torch.zeros(env.num_envs, model.lstm_dim, device=model.device)

    seed = random.randint(0, 2**31 - 1)  # env_loop.py:21 (python RNG stream kept)
    obs, _ = env.reset(seed=[seed + i for i in range(env.num_envs)])

    while True:
        hx, cx = hx.detach(), cx.detach()
        records, infos = [], []
        for n in range(num_steps):
            logits_act, val, (hx, cx) = model.predict_act_value(obs, (hx, cx))
            act = Categorical(logits=logits_act).sample()
            if random.random() < epsilon:
                act = torch.randint(low=0, high=env.num_actions, size=(obs.size(0),), device=obs.device)

            next_obs, rew, end, trunc, info = env.step(act)

            if n > 0:  # value of THIS step's obs bootstraps the previous transition (env_loop.py:39-43)
                val_bootstrap = val.detach().clone()
                if dead.any():
                    val_bootstrap[dead] = val_final_obs
                records[-1][-1] = val_bootstrap

            dead = torch.logical_or(end, trunc)
            if dead.any():
                with torch.no_grad():
                    _, val_final_obs, _ = model.predict_act_value(info["final_observation"], (hx[dead], cx[dead]))
                keep = 1 - dead.float().unsqueeze(1)
                hx, cx = hx * keep, cx * keep
                if "burnin_obs" in info:
                    burnin = info["burnin_obs"]
                    for i in range(burnin.size(1)):
                        _, _, (hx[dead], cx[dead]) = model.predict_act_value(burnin[:, i], (hx[dead], cx[dead]))

            records.append([obs, act, rew, end, trunc, logits_act, val, None])
            infos.append(info)
            obs = next_obs

        with torch.no_grad():
            _, val_bootstrap, _ = model.predict_act_value(next_obs, (hx, cx))  # hx/cx not advanced (env_loop.py:64-65)
        if dead.any():
            val_bootstrap[dead] = val_final_obs
        records[-1][-1] = val_bootstrap

        stacked = tuple(torch.stack(x, dim=1) for x in zip(*records))
        num_steps = yield (*stacked, infos)
