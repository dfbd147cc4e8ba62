"""GPU diagnostic: where one imagined environment step (bench.py cfg 3: 32 envs) spends its time -- sampler, reward/termination
model, policy (with and without autograd node), the Python of WorldModelEnv.step / the env loop.  Wall time with a device
synchronisation around each component, median of `reps`."""
import os
import statistics
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def timed(fn, reps=15, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts)


def main():
    from diamond_b200.envs import WorldModelEnv, WorldModelEnvConfig
    from diamond_b200.models.actor_critic import ActorCritic, ActorCriticConfig, ActorCriticLossConfig
    from diamond_b200.models.diffusion import Denoiser, DenoiserConfig, DiffusionSamplerConfig, InnerModelConfig
    from diamond_b200.models.rew_end_model import RewEndModel, RewEndModelConfig
    from diamond_b200.synthetic import frame_stacks, randomize_module_

    dev = torch.device("cuda:0")
    envs = 32
    den = Denoiser(DenoiserConfig(InnerModelConfig(3, 4, 256, [2, 2, 2, 2], [64] * 4, [0] * 4, 4), 0.5, 0.3))
    randomize_module_(den.inner_model, 2024)
    rem = RewEndModel(RewEndModelConfig(512, 3, 64, 128, [2, 2, 2, 2], [32] * 4, [0] * 4, 4))
    randomize_module_(rem, 2025)
    ac = ActorCritic(ActorCriticConfig(512, 3, 64, [32, 32, 64, 64], [1, 1, 1, 1], 4))
    randomize_module_(ac, 2026)
    den, rem, ac = den.to(dev).eval(), rem.to(dev).eval(), ac.to(dev).train()

    class Loader:
        batch_sampler = types.SimpleNamespace(batch_size=envs)

        def __iter__(self):
            k = 0
            while True:
                obs, act, _ = frame_stacks(envs, 4, 3, 64, 64, 4, 1000 + k)
                k += 1
                yield types.SimpleNamespace(obs=obs, act=act)

    env = WorldModelEnv(den, rem, Loader(), WorldModelEnvConfig(15, 4, DiffusionSamplerConfig(3)))
    obs, _ = env.reset()
    act = torch.randint(0, 4, (envs,), device=dev)
    hx = torch.zeros(envs, 512, device=dev); cx = torch.zeros(envs, 512, device=dev)
    print("predict_next_obs (sampler on the ring)   ms:", round(timed(lambda: env.predict_next_obs()), 3))
    nxt = env.predict_next_obs()[0]
    print("predict_rew_end (native rew/end + 2 draws) ms:", round(timed(lambda: env.predict_rew_end(nxt.unsqueeze(1))), 3))
    with torch.no_grad():
        print("policy predict_act_value, no grad        ms:", round(timed(lambda: ac.predict_act_value(obs, (hx, cx))), 3))
    print("policy predict_act_value, autograd node  ms:", round(timed(lambda: ac.predict_act_value(obs, (hx, cx))), 3))

    def fwd_bwd():
        logits, val, (h2, c2) = ac.predict_act_value(obs, (hx, cx))
        (logits.sum() + val.sum() + h2.sum()).backward()
    print("policy forward + backward (one node)      ms:", round(timed(fwd_bwd), 3))
    print("env.step (sampler + rew/end + bookkeeping) ms:", round(timed(lambda: env.step(act)), 3))
    ac.setup_training(env, ActorCriticLossConfig(15, 0.985, 0.95, 1.0, 0.001))

    def update():
        loss, _ = ac()
        loss.backward()
    print("ActorCritic.forward() + backward, 15 steps ms:", round(timed(update, reps=3, warm=1), 3))

    def fwd_only():
        with torch.no_grad():
            pass
        loss, _ = ac()
        return loss
    print("ActorCritic.forward() alone, 15 steps       ms:", round(timed(fwd_only, reps=3, warm=1), 3))
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    update()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
    # python-side cost of the native plumbing
    t0 = time.perf_counter()
    for _ in range(100):
        ac._native()
    print("ActorCritic._native() (weight-change check) ms/call:", round((time.perf_counter() - t0) * 10, 4))
    t0 = time.perf_counter()
    for _ in range(100):
        ac.grad_layout()
    print("ActorCritic.grad_layout() ms/call:", round((time.perf_counter() - t0) * 10, 4))
    t0 = time.perf_counter()
    for _ in range(100):
        den.inner_model.native(0.5, 0.3)
    print("InnerModel.native() ms/call:", round((time.perf_counter() - t0) * 10, 4))


if __name__ == "__main__":
    main()
