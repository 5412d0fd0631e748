"""Training wrappers around the batched envs, with brax's names and semantics (brax 0.10.3
`brax/envs/wrappers/training.py`: `wrap`, `EpisodeWrapper`, `AutoResetWrapper`; SURVEY.md section 8f N3 -- the
reference imports brax's PPO, `agents_env/agent_mimic.py:14-15`, whose `train` applies exactly these).

`VmapWrapper` is the identity here: the envs of this package are batched natively (one warp per env); it only checks the batch size.

  EpisodeWrapper   info['steps'] / info['truncation'], `done` forced at `episode_length`, `action_repeat` inner steps
                   with summed reward (elementwise bookkeeping on the device);
  AutoResetWrapper envs whose episode ended restart inside one kernel (`pbrl_autoreset`): qpos0 in reference mode, the
                   reference pose at a fresh clip phase in mimic mode (reference-state initialisation); `obs` becomes the
                   first observation of the new episode while `reward` / `done` / `metrics` keep the terminal values.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .envs import State


class Wrapper:
    """brax.envs.base.Wrapper: forwards everything to the wrapped env."""

    def __init__(self, env):
        self.env = env

    def reset(self, rng=None) -> State:
        return self.env.reset(rng)

    def step(self, state: State, action) -> State:
        return self.env.step(state, action)

    @property
    def unwrapped(self):
        return getattr(self.env, "unwrapped", self.env)

    def __getattr__(self, name):
        if name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)


class EpisodeWrapper(Wrapper):
    """Maintains episode step count and sets done at episode end (brax training.py: EpisodeWrapper)."""

    def __init__(self, env, episode_length: int, action_repeat: int = 1):
        super().__init__(env)
        self.episode_length = int(episode_length)
        self.action_repeat = int(action_repeat)

    def reset(self, rng=None) -> State:
        state = self.env.reset(rng)
        n, dev = state.done.shape[0], state.done.device
        state.info["steps"] = torch.zeros(n, dtype=torch.int32, device=dev)
        state.info["truncation"] = torch.zeros(n, dtype=torch.float32, device=dev)
        return state

    def step(self, state: State, action) -> State:
        total = None
        for _ in range(self.action_repeat):
            state = self.env.step(state, action)
            total = state.reward.clone() if total is None else total + state.reward
        if self.action_repeat > 1:
            state.reward.copy_(total)
        steps = state.info["steps"]
        steps += self.action_repeat
        over = steps >= self.episode_length
        state.info["truncation"] = torch.where(over, 1.0 - state.done, torch.zeros_like(state.done))
        state.done.copy_(torch.where(over, torch.ones_like(state.done), state.done))
        return state


class VmapWrapper(Wrapper):
    """brax training.py: VmapWrapper(env, batch_size) vectorises a single-env step with jax.vmap.  The envs here take and
    return batched states already, so this wrapper forwards everything; a `batch_size` that is not the env's `num_envs` is refused."""

    def __init__(self, env, batch_size=None):
        super().__init__(env)
        n = getattr(env, "num_envs", None)
        if batch_size is not None and n is not None and int(batch_size) != int(n):
            raise ValueError(f"VmapWrapper: batch_size {batch_size} != num_envs {n} (the batch is fixed when the env is built)")
        self.batch_size = batch_size


class AutoResetWrapper(Wrapper):
    """Automatically resets envs that are done (brax training.py: AutoResetWrapper).

    brax copies `info['first_pipeline_state']` / `info['first_obs']` over the finished envs; here the restart runs in a
    kernel.  In mimic mode every restart draws a fresh clip phase (reference-state initialisation) unless
    `resample_phase=False`, which restarts at the phase the env was first reset to (brax's exact behaviour)."""

    def __init__(self, env, resample_phase: bool = True, seed: int = 0):
        super().__init__(env)
        self.resample_phase = resample_phase
        self._gen = None
        self._seed = seed

    def reset(self, rng=None) -> State:
        state = self.env.reset(rng)
        state.info["first_obs"] = state.obs.clone()
        return state

    def step(self, state: State, action) -> State:
        if "steps" in state.info:  # the count of an env that finished in the previous call restarts now
            steps = state.info["steps"]
            steps.mul_((state.done == 0).to(steps.dtype))
        state.done.zero_()
        state = self.env.step(state, action)
        base = self.unwrapped
        nxt = None
        if base.reward_mode == 1 and self.resample_phase:
            if self._gen is None:
                self._gen = torch.Generator(device=state.done.device)
                self._gen.manual_seed(self._seed)
            dur = torch.tensor(base.clip_durations, device=state.done.device)[state.info["motion_id"].long()]
            nxt = torch.rand(state.done.shape[0], generator=self._gen, device=state.done.device) * dur
        b = base._buffers(state)
        _lib.check(_lib.lib().pbrl_autoreset(base.model.handle, ctypes.c_int(base.num_envs), ctypes.c_int(base.reward_mode),
                                             _lib._ptr(state.done), _lib._ptr(nxt), None, ctypes.byref(b), base._stream()))
        return state


def wrap(env, episode_length: int = 1000, action_repeat: int = 1, **autoreset_kwargs):
    """brax.envs.wrappers.training.wrap: EpisodeWrapper, then AutoResetWrapper (the vmap layer is the env itself)."""
    return AutoResetWrapper(EpisodeWrapper(env, episode_length, action_repeat), **autoreset_kwargs)
