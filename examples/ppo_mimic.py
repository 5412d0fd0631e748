#!/usr/bin/env python
"""A PPO loop over the batched mimic env (the consumer side of the hot path).

The reference imports brax's PPO (`agents_env/agent_mimic.py:14-15`) and never trains; brax / JAX are not installable in
this image, so this is the same loop written against the env's own facade: `wrappers.wrap(HumanoidTrain(...))` (brax's
EpisodeWrapper + AutoResetWrapper), a Gaussian MLP policy whose action is the PD target of the 28 actuated joints as an
offset from the clip's pose, GAE(lambda), the clipped surrogate.  Everything stays on the device: the env hands out torch
tensors that alias the kernel's buffers, the rollout never touches the host.

    python examples/ppo_mimic.py --envs 4096 --iters 20

It is an example and a smoke test (tests/test_gpu_env_api.py::test_ppo_loop_consumes_the_env), not a tuned trainer.
"""
import argparse
import importlib
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
PKG = "physics-based-rl_b200"


class ActorCritic(torch.nn.Module):
    def __init__(self, obs_dim, act_dim, hidden=256):
        super().__init__()
        mlp = lambda out: torch.nn.Sequential(torch.nn.Linear(obs_dim, hidden), torch.nn.Tanh(), torch.nn.Linear(hidden, hidden),
                                              torch.nn.Tanh(), torch.nn.Linear(hidden, out))
        self.pi, self.v = mlp(act_dim), mlp(1)
        self.log_std = torch.nn.Parameter(torch.full((act_dim,), -1.5))
        with torch.no_grad():   # start at the clip's pose: zero offset
            self.pi[-1].weight.mul_(0.01)
            self.pi[-1].bias.zero_()

    def dist(self, obs):
        return torch.distributions.Normal(self.pi(obs), self.log_std.exp())


def make_env(num_envs, episode_length=64, device=0, clips=None):
    envs = importlib.import_module(PKG + ".envs")
    wrappers = importlib.import_module(PKG + ".wrappers")
    util_data = importlib.import_module(PKG + ".util_data")
    mjcf = importlib.import_module(PKG + ".mjcf")
    mocap = importlib.import_module(PKG + ".mocap")
    model = mjcf.load_model(ROOT / PKG / "assets" / "final_humanoid.model.json")
    table = mocap.MotionTable.load(ROOT / PKG / "assets" / "motions.npz")
    kp, kd = util_data.generate_kp_kd_gains()
    kw = dict(num_envs=num_envs, device=device, reward_mode="mimic", motion_table=table)
    if clips:
        kw["clips"] = clips
    base = envs.HumanoidTrain(None, None, 0.0, {}, model, kp, kd, **kw)
    return wrappers.wrap(base, episode_length=episode_length), model


def features(state):
    """Policy input: the env's observation [qpos, qvel] without the root's horizontal position, plus the clip phase."""
    obs = state.obs
    phase = (state.info["ref_t0"] + state.pipeline_state.time).unsqueeze(1)
    return torch.cat([obs[:, 2:], torch.sin(phase), torch.cos(phase)], dim=1)


def train(num_envs=2048, iters=10, horizon=16, epochs=2, minibatches=4, gamma=0.95, lam=0.95, clip=0.2, lr=3e-4, seed=0, log=print):
    torch.manual_seed(seed)
    env, model = make_env(num_envs)
    dev = env.unwrapped.device
    lo = torch.as_tensor(model.jnt_range[1:, 0], dtype=torch.float32, device=dev)
    hi = torch.as_tensor(model.jnt_range[1:, 1], dtype=torch.float32, device=dev)
    state = env.reset(seed)
    x = features(state)
    net = ActorCritic(x.shape[1], 28).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    history = []
    for it in range(iters):
        t0 = time.time()
        buf = dict(x=[], a=[], logp=[], v=[], r=[], d=[])
        with torch.no_grad():
            for _ in range(horizon):
                dist = net.dist(x)
                a = dist.sample()
                # PD target = the pose the env currently holds + a bounded offset, inside the joint ranges
                target = torch.minimum(torch.maximum(state.pipeline_state.qpos[:, 7:] + 0.3 * torch.tanh(a), lo), hi)
                buf["x"].append(x); buf["a"].append(a); buf["logp"].append(dist.log_prob(a).sum(1)); buf["v"].append(net.v(x).squeeze(1))
                state = env.step(state, target.contiguous())
                buf["r"].append(state.reward[:, 0].clone()); buf["d"].append(state.done.clone())
                x = features(state)
            last_v = net.v(x).squeeze(1)
        T = horizon
        adv = torch.zeros(T, num_envs, device=dev)
        gae = torch.zeros(num_envs, device=dev)
        for t in reversed(range(T)):
            nv = last_v if t == T - 1 else buf["v"][t + 1]
            nd = 1.0 - buf["d"][t]
            delta = buf["r"][t] + gamma * nv * nd - buf["v"][t]
            gae = delta + gamma * lam * nd * gae
            adv[t] = gae
        ret = adv + torch.stack(buf["v"])
        flat = lambda k: torch.stack(buf[k]).reshape(T * num_envs, *buf[k][0].shape[1:])
        X, A, LP, ADV, RET = flat("x"), flat("a"), flat("logp"), adv.reshape(-1), ret.reshape(-1)
        ADV = (ADV - ADV.mean()) / (ADV.std() + 1e-6)
        for _ in range(epochs):
            perm = torch.randperm(X.shape[0], device=dev)
            for mb in perm.chunk(minibatches):
                dist = net.dist(X[mb])
                ratio = (dist.log_prob(A[mb]).sum(1) - LP[mb]).exp()
                pg = -torch.min(ratio * ADV[mb], ratio.clamp(1 - clip, 1 + clip) * ADV[mb]).mean()
                vl = 0.5 * (net.v(X[mb]).squeeze(1) - RET[mb]).pow(2).mean()
                loss = pg + 0.5 * vl - 1e-3 * dist.entropy().sum(1).mean()
                opt.zero_grad(set_to_none=True)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
                opt.step()
        torch.cuda.synchronize()
        dt = time.time() - t0
        rec = dict(iter=it, mean_reward=float(torch.stack(buf["r"]).mean()), done_frac=float(torch.stack(buf["d"]).mean()),
                   loss=float(loss.detach()), env_steps_per_s=horizon * num_envs / dt)
        history.append(rec)
        log(rec)
    return history, net


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--horizon", type=int, default=16)
    a = ap.parse_args()
    train(a.envs, a.iters, a.horizon)
