"""Batched env classes with the reference's interface, running on libpbrl (CUDA, sm_100a).

Mirrors, for N environments at once:
  HumanoidDiff    /root/reference/agents_env/agent_env_template.py:33-155
  HumanoidTrain   /root/reference/agents_env/agent_mimic.py:34-271
  HumanoidReplay  /root/reference/agents_env/agent_replay_motion.py:34-83
and brax's `State(pipeline_state, obs, reward, done, metrics, info)` container.  Same method names,
argument meaning and return layout; arrays are torch CUDA tensors with a leading batch dimension
(the reference never batches; `jax.vmap` of its step is what this replaces).

Differences that are deliberate and documented:
  * `step` updates the state buffers in place (the State passed in is donated); use `State.clone()`
    to keep a copy.
  * `pipeline_state.qM` / `.qfrc_bias` are evaluated on demand from the lagged state (they are the
    fields of the last `mjx.forward`) instead of being stored per env.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import numpy as np
import torch

from . import _lib
from . import pds_controllers
from .mjcf import ModelConst, load_model

PD_TORQUE, PD_STANDARD, PD_FEEDBACK, PD_STABLE = 0, 1, 2, 3


class PipelineState:
    """The subset of brax.mjx `State` the reference reads."""

    def __init__(self, env, qpos, qvel, time, warm, ctrl, qpos_lag, qvel_lag):
        self._env = env
        self.qpos, self.qvel, self.time, self.qacc_warmstart, self.ctrl = qpos, qvel, time, warm, ctrl
        self.qpos_lag, self.qvel_lag = qpos_lag, qvel_lag

    q = property(lambda self: self.qpos)
    qd = property(lambda self: self.qvel)

    @property
    def qfrc_applied(self):
        return torch.zeros_like(self.qvel)

    def _lag_dynamics(self):
        N = self.qpos.shape[0]
        env = self._env
        tau = torch.empty((N, 28), device=self.qpos.device)
        qM = torch.empty((N, 34, 34), device=self.qpos.device)
        bias = torch.empty((N, 34), device=self.qpos.device)
        zeros = torch.zeros((N, 28), device=self.qpos.device)
        _lib.check(_lib.lib().pbrl_pd_torque(
            env.model.handle, ctypes.c_int(N), ctypes.c_int(PD_TORQUE), ctypes.c_int(1), _lib._ptr(zeros),
            _lib._ptr(self.qpos), _lib._ptr(self.qvel), _lib._ptr(self.qpos_lag), _lib._ptr(self.qvel_lag),
            _lib._ptr(env._kp0), _lib._ptr(env._kp0), _lib._ptr(self.time), ctypes.c_float(0.0), _lib._ptr(tau),
            _lib._ptr(qM), _lib._ptr(bias), _lib.current_stream_ptr(self.qpos.device)))
        return qM, bias

    @property
    def qM(self):
        return self._lag_dynamics()[0]

    @property
    def qfrc_bias(self):
        return self._lag_dynamics()[1]

    def clone(self):
        c = lambda t: t.clone()
        return PipelineState(self._env, c(self.qpos), c(self.qvel), c(self.time), c(self.qacc_warmstart), c(self.ctrl),
                             c(self.qpos_lag), c(self.qvel_lag))


@dataclass
class State:
    pipeline_state: PipelineState
    obs: torch.Tensor
    reward: torch.Tensor
    done: torch.Tensor
    metrics: Dict[str, torch.Tensor] = field(default_factory=dict)
    info: Dict[str, Any] = field(default_factory=dict)

    def replace(self, **kw):
        d = dict(pipeline_state=self.pipeline_state, obs=self.obs, reward=self.reward, done=self.done,
                 metrics=self.metrics, info=self.info)
        d.update(kw)
        return State(**d)

    def clone(self):
        """An independent copy: every tensor (state, info buffers the kernels write -- metrics, status, cost, env_order,
        motion_id, ref_t0, stats, the wrappers' step counters) is duplicated, and `metrics` is re-bound to views of the
        copy's own metrics buffer, so stepping either state leaves the other untouched."""
        info = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.info.items()}
        st = State(self.pipeline_state.clone(), self.obs.clone(), self.reward.clone(), self.done.clone(), {}, info)
        if "_metrics" in info:
            HumanoidDiff._bind_metrics(st)
        else:
            st.metrics = {k: v.clone() for k, v in self.metrics.items()}
        return st


class HumanoidDiff:
    """agent_env_template.py:33 -- loads the model, fixes timestep 0.002 and n_frames = int((1/60)/0.002) = 8."""

    def __init__(self, reference_trajectory_qpos, reference_trajectory_qvel, duration_trajectory, dict_duration,
                 model_path, **kwargs):
        self.num_envs = int(kwargs.pop("num_envs", 1))
        dev = kwargs.pop("device", 0)
        self.device = torch.device("cuda", dev if isinstance(dev, int) else torch.device(dev).index or 0)
        const = model_path if isinstance(model_path, ModelConst) else load_model(model_path)
        self._dt = 1 / 60
        const.timestep = float(kwargs.pop("timestep", 0.002))          # agent_env_template.py:65
        self.n_frames = int(kwargs.pop("n_frames", int(self._dt / 0.002)))  # agent_env_template.py:68
        self.sys = const
        self.sys.dt = const.timestep
        # mimic mode (build-defined): reward_mode="mimic" + a MotionTable
        self.reward_mode = {"reference": 0, "mimic": 1}[kwargs.pop("reward_mode", "reference")]
        self.motion_table = kwargs.pop("motion_table", None)
        self._mimic_keep = None
        mdesc = None
        if self.reward_mode == 1:
            from .mimic import make_mimic_desc
            if self.motion_table is None:
                raise ValueError("reward_mode='mimic' needs motion_table=MotionTable")
            mdesc, frames = make_mimic_desc(const, self.motion_table, kwargs.pop("clips", None))
            self._mimic_keep = (mdesc, frames)
            self.clip_durations = [float(mdesc.clip_duration[c]) for c in range(mdesc.nclip)]
        self.model = _lib.Model(const, device=self.device.index, mimic=mdesc)
        self.reference_trajectory_qpos = reference_trajectory_qpos
        self.reference_trajectory_qvel = reference_trajectory_qvel
        self.duration_trajectory = duration_trajectory
        self.dict_duration = dict_duration
        self._kp0 = torch.zeros(28, device=self.device)

    dt = property(lambda self: self.sys.timestep * self.n_frames)
    action_size = property(lambda self: self.sys.nu)
    observation_size = property(lambda self: self.sys.nq + self.sys.nv)

    def set_new_trajectory(self, trajectory):
        self.reference_trajectory_qpos = np.asarray(trajectory.data_pos)
        self.reference_trajectory_qvel = np.asarray(trajectory.data_vel)
        self.duration_trajectory = trajectory.total_time
        self.dict_duration = trajectory.duration_dict

    # ---- helpers ----------------------------------------------------------------------------------
    def _alloc_state(self):
        N, d = self.num_envs, self.device
        z = lambda *s: torch.zeros(s, device=d, dtype=torch.float32)
        ps = PipelineState(self, z(N, 35), z(N, 34), z(N), z(N, 34), z(N, 28), z(N, 35), z(N, 34))
        metrics = z(N, 3)
        st = State(ps, z(N, 69), z(N, 3), z(N), {}, {})
        st.info["_metrics"] = metrics
        st.info["status"] = torch.zeros(N, device=d, dtype=torch.int32)
        # scheduling hints of pbrl_step: work estimate of the last step and the processing order derived from it
        st.info["cost"] = torch.zeros(N, device=d, dtype=torch.int32)
        st.info["env_order"] = torch.arange(N, device=d, dtype=torch.int32)
        if self.reward_mode == 1:
            st.info["motion_id"] = torch.zeros(N, device=d, dtype=torch.int32)
            st.info["ref_t0"] = z(N)
            st.info["stats"] = z(8)
        self._bind_metrics(st)
        return st

    @staticmethod
    def _bind_metrics(st):
        m = st.info["_metrics"]
        st.metrics = {"step_index": m[:, 0], "pose_error": m[:, 1], "fall": m[:, 2]}

    def _buffers(self, st: State):
        ps = st.pipeline_state
        return _lib.make_buffers(qpos=ps.qpos, qvel=ps.qvel, time=ps.time, warm=ps.qacc_warmstart, ctrl=ps.ctrl,
                                 qpos_lag=ps.qpos_lag, qvel_lag=ps.qvel_lag, obs=st.obs, reward=st.reward, done=st.done,
                                 metrics=st.info["_metrics"], status=st.info["status"], motion_id=st.info.get("motion_id"),
                                 ref_t0=st.info.get("ref_t0"), stats=st.info.get("stats"), cost=st.info.get("cost"),
                                 env_order=st.info.get("env_order"))

    def _stream(self):
        return _lib.current_stream_ptr(self.device)

    # ---- API --------------------------------------------------------------------------------------
    def reset(self, rng=None) -> State:
        """agent_env_template.py:88-112: qpos0, zero velocity, pipeline_init; `rng` is unused there too."""
        st = self._alloc_state()
        if self.reward_mode == 1:  # reference-state initialisation: clip = env id mod nclip, phase ~ U[0,1)
            g = torch.Generator(device="cpu")
            g.manual_seed(int(rng) if isinstance(rng, (int, np.integer)) else 0)
            nclip = len(self.clip_durations)
            mid = torch.arange(self.num_envs, dtype=torch.int32) % nclip
            t0 = torch.rand(self.num_envs, generator=g) * torch.tensor(self.clip_durations)[mid.long()]
            st.info["motion_id"].copy_(mid)
            st.info["ref_t0"].copy_(t0)
        b = self._buffers(st)
        _lib.check(_lib.lib().pbrl_reset(self.model.handle, ctypes.c_int(self.num_envs), ctypes.c_int(self.reward_mode),
                                         ctypes.byref(b), self._stream()))
        return st

    def pipeline_init(self, qpos, qvel) -> PipelineState:
        c = lambda t: torch.as_tensor(t, dtype=torch.float32, device=self.device).contiguous().clone()
        qpos, qvel = c(qpos), c(qvel)
        N = qpos.shape[0]
        z = lambda *s: torch.zeros(s, device=self.device)
        return PipelineState(self, qpos, qvel, z(N), z(N, 34), z(N, 28), qpos.clone(), qvel.clone())

    def _get_obs(self, data: PipelineState, action=None):
        return torch.cat([data.qpos, data.qvel], dim=-1)  # agent_env_template.py:151-155

    def _check_state(self, st: State):
        """The kernels read N rows of every state buffer through raw pointers: refuse a state of another batch size."""
        ps, N = st.pipeline_state, self.num_envs
        for name, t, w in (("qpos", ps.qpos, (35,)), ("qvel", ps.qvel, (34,)), ("time", ps.time, ()), ("qacc_warmstart", ps.qacc_warmstart, (34,)),
                           ("ctrl", ps.ctrl, (28,)), ("qpos_lag", ps.qpos_lag, (35,)), ("qvel_lag", ps.qvel_lag, (34,)),
                           ("obs", st.obs, (69,)), ("reward", st.reward, (3,)), ("done", st.done, ())):
            if tuple(t.shape) != (N, *w) or t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                raise ValueError(f"state.{name}: expected a contiguous float32 tensor of shape {(N, *w)} on {self.device}, got {tuple(t.shape)} {t.dtype} on {t.device}")

    def _launch_step(self, st: State, action, pd_mode, pd_dt, pd_time=None, ctrl_mask=0x0FFFFFFF, kp=None, kd=None):
        N = self.num_envs
        action = torch.as_tensor(action, dtype=torch.float32, device=self.device)
        if tuple(action.shape) in ((28,), (28, 4)):  # one action for every env (the reference's unbatched call)
            action = action.expand(N, *action.shape)
        if tuple(action.shape) not in ((N, 28), (N, 28, 4)):
            raise ValueError(f"action: expected shape ({N}, 28) [targets / torques] or ({N}, 28, 4) [cubic coefficients], got {tuple(action.shape)}")
        action = action.contiguous()
        cols = 1 if action.dim() == 2 else 4
        if pd_mode != PD_TORQUE:
            for name, g in (("kp", kp), ("kd", kd)):
                if g is None or g.numel() != 28 or g.dtype != torch.float32 or g.device != self.device:
                    raise ValueError(f"{name}: expected 28 float32 gains on {self.device}")
        if pd_time is not None:
            pd_time = torch.as_tensor(pd_time, dtype=torch.float32, device=self.device)
            if pd_time.numel() not in (1, N):
                raise ValueError(f"time: expected a scalar or {N} values, got {tuple(pd_time.shape)}")
            pd_time = pd_time.reshape(-1).expand(N).contiguous()
        self._check_state(st)
        b = self._buffers(st)
        _lib.check(_lib.lib().pbrl_step(
            self.model.handle, ctypes.c_int(self.num_envs), ctypes.c_int(self.n_frames), ctypes.c_int(pd_mode),
            ctypes.c_int(cols), _lib._ptr(action), _lib._ptr(kp), _lib._ptr(kd), ctypes.c_float(pd_dt),
            _lib._ptr(pd_time), ctypes.c_uint(ctrl_mask), ctypes.c_int(self.reward_mode), ctypes.byref(b), self._stream()))
        if st.info.get("cost") is not None and st.info.get("env_order") is not None:  # next step: similar-cost envs share a CTA
            _lib.check(_lib.lib().pbrl_balance_order(self.model.handle, ctypes.c_int(self.num_envs), _lib._ptr(st.info["cost"]),
                                                     _lib._ptr(st.info["env_order"]), self._stream()))
        return st


class HumanoidTrain(HumanoidDiff):
    """agent_mimic.py:34 -- PD-driven env."""

    def __init__(self, reference_trajectory_qpos, reference_trajectory_qvel, duration_trajectory, dict_duration,
                 model_path, kp_gains, kd_gains, **kwargs):
        super().__init__(reference_trajectory_qpos, reference_trajectory_qvel, duration_trajectory, dict_duration,
                         model_path, **kwargs)
        self.kp__gains = torch.as_tensor(np.asarray(kp_gains), dtype=torch.float32, device=self.device).contiguous()
        self.kd__gains = torch.as_tensor(np.asarray(kd_gains), dtype=torch.float32, device=self.device).contiguous()
        self.pd_function = pds_controllers.stable_pd_controller_custom_trajectory

    def set_pd_callback(self, pd_control):
        self.pd_function = pd_control

    def _pd_step(self, st, target, time, ctrl_mask=0x0FFFFFFF):
        mode = getattr(self.pd_function, "pd_mode", None)
        if mode is not None:  # one of the built-in laws: fused into the step kernel
            return self._launch_step(st, target, mode, self.sys.dt, pd_time=time, ctrl_mask=ctrl_mask,
                                     kp=self.kp__gains, kd=self.kd__gains)
        ps = st.pipeline_state  # arbitrary callable: evaluate it, then run the torque-driven step
        torque = self.pd_function(target, self.sys, st, ps.q, ps.qd, self.kp__gains, self.kd__gains,
                                  ps.time if time is None else time, self.sys.dt)
        return self._launch_step(st, torque, PD_TORQUE, self.sys.dt, ctrl_mask=ctrl_mask)

    def step(self, state: State, action) -> State:
        """agent_mimic.py:58-91"""
        return self._pd_step(state, action, None)

    def step_custom_target(self, state: State, action, custom_target, time) -> State:
        """agent_mimic.py:200-236"""
        return self._pd_step(state, custom_target, time)

    def step_custom_target_and_outside(self, state: State, action) -> State:
        """agent_mimic.py:239-271: the torque was computed outside."""
        return self._launch_step(state, action, PD_TORQUE, self.sys.dt)

    @staticmethod
    def _mask(*joints):
        m = 0
        for j in joints:
            m |= 1 << int(j)
        return m

    def step_selected_joints_custom_target_and_joints(self, state, action, custom_target, joints1, joints2, joints3,
                                                      joints4, joints5, joints6, time) -> State:
        """agent_mimic.py:94-139: only joints1..4 take the new torque, the rest keep the previous ctrl."""
        return self._pd_step(state, custom_target, time, self._mask(joints1, joints2, joints3, joints4))

    def step_selected_joints_custom_target_and_joints6(self, state, action, custom_target, joints1, joints2, joints3,
                                                       joints4, joints5, joints6, time) -> State:
        """agent_mimic.py:145-190"""
        return self._pd_step(state, custom_target, time, self._mask(joints1, joints2, joints3, joints4, joints5, joints6))


class HumanoidReplay(HumanoidDiff):
    """agent_replay_motion.py:34 -- kinematic replay: `pipeline_init` at the state's qpos/qvel."""

    def step(self, state: State, action=None, want_poses=False) -> State:
        ps = state.pipeline_state
        N = ps.qpos.shape[0]
        xpos = torch.empty((N, 14, 3), device=self.device) if want_poses else None
        xquat = torch.empty((N, 14, 4), device=self.device) if want_poses else None
        _lib.check(_lib.lib().pbrl_replay(self.model.handle, ctypes.c_int(N), _lib._ptr(ps.qpos), _lib._ptr(ps.qvel),
                                          _lib._ptr(state.obs), _lib._ptr(state.reward), _lib._ptr(state.done),
                                          _lib._ptr(xpos), _lib._ptr(xquat), self._stream()))
        ps.time.zero_()  # pipeline_init builds fresh data: time = 0
        ps.qpos_lag.copy_(ps.qpos)
        ps.qvel_lag.copy_(ps.qvel)
        if want_poses:
            state.info["xpos"], state.info["xquat"] = xpos, xquat
        return state


    def rollout(self, frame0=None, num_steps=None, want_poses=True):
        """The replay loop of the reference's notebooks (testing_forward_env.ipynb cell 10: write frame i of the reference
        trajectory into the state, step, next frame) as ONE launch: env e shows frame (frame0[e] + k) mod F at step k.
        Returns step-major tensors: obs (K, N, 69), reward (K, N, 3) zeros, done (K, N) ones and, with want_poses,
        xpos (K, N, 14, 3), xquat (K, N, 14, 4)."""
        if getattr(self, "_traj_dev", None) is None:
            qp = torch.as_tensor(np.asarray(self.reference_trajectory_qpos), dtype=torch.float32, device=self.device).contiguous()
            qv = torch.as_tensor(np.asarray(self.reference_trajectory_qvel), dtype=torch.float32, device=self.device).contiguous()
            self._traj_dev = (qp, qv)
        qp, qv = self._traj_dev
        F, N = int(qp.shape[0]), self.num_envs
        K = F if num_steps is None else int(num_steps)
        f0 = None if frame0 is None else torch.as_tensor(frame0, dtype=torch.int32, device=self.device).contiguous()
        e = lambda *s: torch.empty(s, device=self.device, dtype=torch.float32)
        out = dict(obs=e(K, N, 69), reward=e(K, N, 3), done=e(K, N))
        if want_poses:
            out.update(xpos=e(K, N, 14, 3), xquat=e(K, N, 14, 4))
        _lib.check(_lib.lib().pbrl_replay_rollout(self.model.handle, ctypes.c_int(N), ctypes.c_int(K), _lib._ptr(qp), _lib._ptr(qv),
                                                  ctypes.c_int(F), _lib._ptr(f0), _lib._ptr(out["obs"]), _lib._ptr(out["reward"]),
                                                  _lib._ptr(out["done"]), _lib._ptr(out.get("xpos")), _lib._ptr(out.get("xquat")),
                                                  self._stream()))
        return out


_REGISTRY = {}


def register_environment(name, cls):
    """brax.envs.register_environment, as the notebooks use it (testing_pds_env.ipynb cell 2)."""
    _REGISTRY[name] = cls


def get_environment(env_name, **kwargs):
    return _REGISTRY[env_name](**kwargs)
