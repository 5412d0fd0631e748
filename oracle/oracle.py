"""ctypes front end of the CPU oracle -- TEST INFRASTRUCTURE (see pbrl_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under physics-based-rl_b200/ does.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def build(force=False):
    """Compile the C restatement (both precisions) with the committed Makefile."""
    libs = [HERE / "liboracle_f64.so", HERE / "liboracle_f32.so"]
    srcs = [HERE / "pbrl_oracle.c", HERE / "pbrl_oracle_env.c", HERE / "pbrl_oracle.h"]
    stale = force or any(not l.exists() or l.stat().st_mtime < max(s.stat().st_mtime for s in srcs) for l in libs)
    if stale:
        subprocess.run(["make", "-C", str(HERE), "-B"], check=True, capture_output=True)
    return libs


class Oracle:
    """One precision of the oracle bound to one model."""

    def __init__(self, model, precision="f64", tolerance=None, iterations=None, self_collision=None):
        """self_collision: evaluate the body-vs-body candidates (default: the model's `self_collision` switch)."""
        if self_collision is None:
            self_collision = bool(getattr(model, "self_collision", 1))
        build()
        self.lib = ctypes.CDLL(str(HERE / f"liboracle_{precision}.so"))
        self.real = np.float64 if precision == "f64" else np.float32
        assert self.lib.orc_real_size() == np.dtype(self.real).itemsize
        self.m = model
        self.nq, self.nv, self.nu, self.nbody = model.nq, model.nv, model.nu, model.nbody
        self._model_buf = np.zeros(self.lib.orc_model_size(), dtype=np.uint8)
        self._fill_model(model, tolerance, iterations)
        if self_collision and getattr(model, "pair_geom1", None) is not None:
            self.set_model_field("pair_geom1", model.pair_geom1)
            self.set_model_field("pair_geom2", model.pair_geom2)
            self.set_model_field("npair", len(model.pair_geom1))
        self._dsize = self.lib.orc_data_size()
        self._dfields = {}
        for fn in ("orc_reset_data", "orc_forward", "orc_step"):
            getattr(self.lib, fn).restype = None

    # ---- model --------------------------------------------------------------------------------
    def _field(self, name):
        off, cnt, isint = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        if self.lib.orc_model_field(name.encode(), ctypes.byref(off), ctypes.byref(cnt), ctypes.byref(isint)):
            raise KeyError(name)
        return off.value, cnt.value, isint.value

    def set_model_field(self, name, value):
        off, cnt, isint = self._field(name)
        dt = np.int32 if isint else np.float64
        v = np.asarray(value, dtype=dt).ravel()
        assert v.size <= cnt, (name, v.size, cnt)
        view = self._model_buf[off:off + cnt * np.dtype(dt).itemsize].view(dt)
        view[:v.size] = v

    def _fill_model(self, m, tolerance, iterations):
        scalars = ["nbody", "nq", "nv", "nu", "njnt", "ngeom", "iterations", "ls_iterations", "timestep",
                   "impratio", "tolerance", "ls_tolerance", "meaninertia"]
        arrays = ["gravity", "solref", "solimp", "body_parent", "body_pos", "body_mass", "body_ipos",
                  "body_inertia", "body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum", "body_invweight0",
                  "jnt_type", "jnt_bodyid", "jnt_qposadr", "jnt_dofadr", "jnt_axis", "jnt_pos", "jnt_limited",
                  "jnt_range", "jnt_margin", "dof_bodyid", "dof_parentid", "dof_armature", "dof_damping",
                  "dof_invweight0", "qpos0", "geom_type", "geom_bodyid", "geom_size", "geom_pos", "geom_quat",
                  "geom_friction", "geom_margin", "geom_gap", "act_dofadr", "act_gear", "act_ctrllimited",
                  "act_ctrlrange"]
        for n in scalars + arrays:
            self.set_model_field(n, getattr(m, n))
        if tolerance is not None:
            self.set_model_field("tolerance", tolerance)
        if iterations is not None:
            self.set_model_field("iterations", iterations)

    @property
    def model_ptr(self):
        return self._model_buf.ctypes.data_as(ctypes.c_void_p)

    # ---- single-env data ------------------------------------------------------------------------
    def new_data(self):
        d = OracleData(self)
        self.lib.orc_reset_data(self.model_ptr, d.ptr)
        return d

    def forward(self, d):
        self.lib.orc_forward(self.model_ptr, d.ptr)

    def step(self, d):
        self.lib.orc_step(self.model_ptr, d.ptr)

    def pd_torque(self, lag, mode, action, q, qd, kp, kd, time, dt):
        action = np.ascontiguousarray(action, dtype=self.real)
        cols = 1 if action.ndim == 1 else action.shape[1]
        tau = np.zeros(self.nu, dtype=self.real)
        c = lambda a: np.ascontiguousarray(a, dtype=self.real).ctypes.data_as(ctypes.c_void_p)
        R = ctypes.c_double if self.real == np.float64 else ctypes.c_float
        self.lib.orc_pd_torque.restype = None
        self.lib.orc_pd_torque(self.model_ptr, lag.ptr, ctypes.c_int(mode), ctypes.c_int(cols), c(action), c(q),
                               c(qd), c(kp), c(kd), R(time), R(dt), tau.ctypes.data_as(ctypes.c_void_p))
        return tau


class OracleData:
    def __init__(self, orc):
        self._orc = orc
        self.buf = np.zeros(orc._dsize // np.dtype(orc.real).itemsize, dtype=orc.real)

    @property
    def ptr(self):
        return self.buf.ctypes.data_as(ctypes.c_void_p)

    def __getattr__(self, name):
        orc = self.__dict__["_orc"]
        if name not in orc._dfields:
            off, cnt = ctypes.c_int(), ctypes.c_int()
            if orc.lib.orc_data_field(name.encode(), ctypes.byref(off), ctypes.byref(cnt)):
                raise AttributeError(name)
            isz = np.dtype(orc.real).itemsize
            orc._dfields[name] = (off.value // isz, cnt.value)
        o, c = orc._dfields[name]
        return self.__dict__["buf"][o:o + c]

    def qM_dense(self):
        nv = self._orc.nv
        return self.qM.reshape(36, 36)[:nv, :nv]

    def copy(self):
        d = OracleData(self._orc)
        d.buf[:] = self.buf
        return d


class _EnvBuffers(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("qpos", "qvel", "time", "warm", "ctrl", "qpos_lag", "qvel_lag", "obs", "reward", "done", "metrics")]


class OracleEnv:
    """Batched env state + reset/step/replay through the oracle's env layer (pbrl_oracle_env.c)."""

    def __init__(self, orc: Oracle, N: int, nthreads: int = 0):
        self.orc, self.N, self.nthreads = orc, N, nthreads
        r = orc.real
        self.qpos = np.zeros((N, orc.nq), r)
        self.qvel = np.zeros((N, orc.nv), r)
        self.time = np.zeros(N, r)
        self.warm = np.zeros((N, orc.nv), r)
        self.ctrl = np.zeros((N, orc.nu), r)
        self.qpos_lag = np.zeros((N, orc.nq), r)
        self.qvel_lag = np.zeros((N, orc.nv), r)
        self.obs = np.zeros((N, orc.nq + orc.nv), r)
        self.reward = np.zeros((N, 3), r)
        self.done = np.zeros(N, r)
        self.metrics = np.zeros((N, 3), r)

    def _bufs(self):
        b = _EnvBuffers()
        for n, _ in _EnvBuffers._fields_:
            setattr(b, n, getattr(self, n).ctypes.data)
        return b

    def _p(self, a):
        return None if a is None else np.ascontiguousarray(a, dtype=self.orc.real).ctypes.data_as(ctypes.c_void_p)

    def reset(self):
        b = self._bufs()
        self.orc.lib.orc_env_reset.restype = None
        self.orc.lib.orc_env_reset(self.orc.model_ptr, ctypes.c_int(self.N), ctypes.byref(b))

    def set_state(self, qpos, qvel, time=None):
        """Reference-state initialisation: state and lag state both at (qpos, qvel)."""
        self.qpos[:] = qpos
        self.qvel[:] = qvel
        self.qpos_lag[:] = qpos
        self.qvel_lag[:] = qvel
        self.warm[:] = 0
        self.ctrl[:] = 0
        self.time[:] = 0 if time is None else time

    def step(self, action, kp, kd, pd_mode=3, pd_dt=0.002, n_frames=8, pd_time=None, ctrl_mask=0x0FFFFFFF):
        r = self.orc.real
        action = np.ascontiguousarray(action, dtype=r)
        cols = 1 if action.ndim == 2 else action.shape[2]
        R = ctypes.c_double if r == np.float64 else ctypes.c_float
        b = self._bufs()
        keep = [np.ascontiguousarray(kp, dtype=r), np.ascontiguousarray(kd, dtype=r),
                None if pd_time is None else np.ascontiguousarray(pd_time, dtype=r)]
        self.orc.lib.orc_env_step.restype = None
        self.orc.lib.orc_env_step(self.orc.model_ptr, ctypes.c_int(self.N), ctypes.c_int(n_frames),
                                  ctypes.c_int(pd_mode), ctypes.c_int(cols), action.ctypes.data_as(ctypes.c_void_p),
                                  keep[0].ctypes.data_as(ctypes.c_void_p), keep[1].ctypes.data_as(ctypes.c_void_p),
                                  R(pd_dt), None if keep[2] is None else keep[2].ctypes.data_as(ctypes.c_void_p),
                                  ctypes.c_uint(ctrl_mask), ctypes.byref(b), ctypes.c_int(self.nthreads))

    def replay(self, qpos, qvel, want_poses=False):
        r = self.orc.real
        qpos = np.ascontiguousarray(qpos, dtype=r)
        qvel = np.ascontiguousarray(qvel, dtype=r)
        N = qpos.shape[0]
        obs = np.zeros((N, self.orc.nq + self.orc.nv), r)
        reward = np.zeros((N, 3), r)
        done = np.zeros(N, r)
        xpos = np.zeros((N, self.orc.nbody, 3), r) if want_poses else None
        xquat = np.zeros((N, self.orc.nbody, 4), r) if want_poses else None
        self.orc.lib.orc_env_replay.restype = None
        self.orc.lib.orc_env_replay(self.orc.model_ptr, ctypes.c_int(N), self._p(qpos), self._p(qvel),
                                    obs.ctypes.data_as(ctypes.c_void_p), reward.ctypes.data_as(ctypes.c_void_p),
                                    done.ctypes.data_as(ctypes.c_void_p),
                                    None if xpos is None else xpos.ctypes.data_as(ctypes.c_void_p),
                                    None if xquat is None else xquat.ctypes.data_as(ctypes.c_void_p),
                                    ctypes.c_int(self.nthreads))
        return obs, reward, done, xpos, xquat

    def pd_torque(self, action, kp, kd, pd_mode, pd_dt, time=None, want_qM=False):
        r = self.orc.real
        action = np.ascontiguousarray(action, dtype=r)
        cols = 1 if action.ndim == 2 else action.shape[2]
        tau = np.zeros((self.N, self.orc.nu), r)
        qM = np.zeros((self.N, self.orc.nv, self.orc.nv), r) if want_qM else None
        bias = np.zeros((self.N, self.orc.nv), r) if want_qM else None
        t = np.ascontiguousarray(self.time if time is None else time, dtype=r)
        R = ctypes.c_double if r == np.float64 else ctypes.c_float
        self.orc.lib.orc_env_pd_torque.restype = None
        self.orc.lib.orc_env_pd_torque(self.orc.model_ptr, ctypes.c_int(self.N), ctypes.c_int(pd_mode),
                                       ctypes.c_int(cols), self._p(action), self._p(self.qpos), self._p(self.qvel),
                                       self._p(self.qpos_lag), self._p(self.qvel_lag), self._p(kp), self._p(kd),
                                       self._p(t), R(pd_dt), tau.ctypes.data_as(ctypes.c_void_p),
                                       None if qM is None else qM.ctypes.data_as(ctypes.c_void_p),
                                       None if bias is None else bias.ctypes.data_as(ctypes.c_void_p))
        return (tau, qM, bias) if want_qM else tau


    # ---- mimic mode (build-defined; pbrl_oracle_env.c) -------------------------------------------------------
    def reset_mimic(self, mimic_desc, motion_id, ref_t0):
        self.motion_id = np.ascontiguousarray(motion_id, np.int32)
        self.ref_t0 = np.ascontiguousarray(ref_t0, self.orc.real)
        self.mimic_desc = mimic_desc
        b = self._bufs()
        self.orc.lib.orc_env_reset_mimic.restype = None
        self.orc.lib.orc_env_reset_mimic(self.orc.model_ptr, ctypes.byref(mimic_desc), ctypes.c_int(self.N),
                                         self.motion_id.ctypes.data_as(ctypes.c_void_p), self._p(self.ref_t0), ctypes.byref(b))

    def step_mimic(self, action, kp, kd, pd_mode=3, pd_dt=0.002, n_frames=8):
        r = self.orc.real
        action = np.ascontiguousarray(action, dtype=r)
        cols = 1 if action.ndim == 2 else action.shape[2]
        R = ctypes.c_double if r == np.float64 else ctypes.c_float
        b = self._bufs()
        stats = np.zeros(8, np.float64)
        kp_, kd_ = np.ascontiguousarray(kp, dtype=r), np.ascontiguousarray(kd, dtype=r)
        self.orc.lib.orc_env_step_mimic.restype = None
        self.orc.lib.orc_env_step_mimic(self.orc.model_ptr, ctypes.byref(self.mimic_desc), ctypes.c_int(self.N),
                                        ctypes.c_int(n_frames), ctypes.c_int(pd_mode), ctypes.c_int(cols),
                                        action.ctypes.data_as(ctypes.c_void_p), kp_.ctypes.data_as(ctypes.c_void_p),
                                        kd_.ctypes.data_as(ctypes.c_void_p), R(pd_dt),
                                        self.motion_id.ctypes.data_as(ctypes.c_void_p), self._p(self.ref_t0), ctypes.byref(b),
                                        stats.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(self.nthreads))
        return stats
