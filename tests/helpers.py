"""Shared test plumbing: assets, the oracle, and two back ends with one interface --
`EmuBackend` (host emulation of the kernel bodies, no GPU) and `GpuBackend` (libpbrl through the C ABI)."""
import ctypes
import importlib
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
PKG = "physics-based-rl_b200"
mjcf = importlib.import_module(PKG + ".mjcf")
desc = importlib.import_module(PKG + "._desc")
util_data = importlib.import_module(PKG + ".util_data")
math_utils = importlib.import_module(PKG + ".math_utils")
mocap = importlib.import_module(PKG + ".mocap")
mimic = importlib.import_module(PKG + ".mimic")
ASSETS = ROOT / PKG / "assets"
GOLDEN = ROOT / "tests" / "golden"

from oracle.oracle import Oracle, OracleEnv  # noqa: E402  (test infrastructure)


def load_model_asset(name):
    return mjcf.load_model(ASSETS / f"{name}.model.json")


def start_trajectory():
    """`start_trajec` of the notebooks (testing_pds_env3.ipynb cells 8-10)."""
    g = math_utils.generate_trajectory
    t = {7: g(1, 3, 0, -1.5), 10: g(1, 3, 0, 1.5), 13: g(1, 3, 0, 1.5), 9: g(1, 3, 0, 1.5), 17: g(1, 3, 0, 0), 24: g(1, 3, 0, 0)}
    return math_utils.start_trajectories(t)


def random_states(model, N, seed=0, gravity_heights=False, vel_scale=0.5, joint_frac=0.3):
    rng = np.random.default_rng(seed)
    lo, hi = model.jnt_range[1:, 0], model.jnt_range[1:, 1]
    qpos = np.tile(model.qpos0, (N, 1))
    qpos[:, 7:] = rng.uniform(joint_frac * lo, joint_frac * hi, (N, 28))
    q = np.array([1.0, 0, 0, 0]) + 0.3 * rng.normal(size=(N, 4))
    qpos[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    if gravity_heights:
        qpos[:, 2] = rng.uniform(0.7, 0.95, N)
    qvel = vel_scale * rng.normal(size=(N, 34))
    return qpos, qvel


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-12, np.abs(b).max()))


def row_relerr(a, b):
    """Per-env max-abs error relative to that env's max-abs value."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).reshape(a.shape[0], -1).max(1) / np.maximum(1e-12, np.abs(b).reshape(b.shape[0], -1).max(1))


STATE_SHAPES = [("qpos", 35), ("qvel", 34), ("time", 0), ("warm", 34), ("ctrl", 28), ("qpos_lag", 35), ("qvel_lag", 34),
                ("obs", 69), ("reward", 3), ("done", 0), ("metrics", 3)]
DBG = dict(xpos=(0, 42), xquat=(42, 56), com=(98, 3), qM=(101, 1156), bias=(1257, 34), qfrc_smooth=(1291, 34),
           qacc_smooth=(1325, 34), qacc=(1359, 34), qfrc_constraint=(1393, 34), ncon_active=(1427, 1), con_dist=(1428, 32),
           con_pos=(1460, 96), niter=(1556, 1), con_force=(1557, 128), lim_force=(1685, 28), nself=(1713, 1),
           self_dist=(1714, 8), self_force=(1722, 32))
DBG_SIZE = 1756


class _HostState:
    def __init__(self, N):
        self.N = N
        for n, w in STATE_SHAPES:
            setattr(self, n, np.zeros((N, w) if w else (N,), np.float32))
        self.status = np.zeros(N, np.int32)
        self.cost = np.zeros(N, np.int32)
        self.motion_id = np.zeros(N, np.int32)
        self.ref_t0 = np.zeros(N, np.float32)
        self.stats = np.zeros(8, np.float32)

    def set_state(self, qpos, qvel):
        self.qpos[:] = qpos
        self.qvel[:] = qvel
        self.qpos_lag[:] = qpos
        self.qvel_lag[:] = qvel
        self.warm[:] = 0
        self.ctrl[:] = 0
        self.time[:] = 0


def build_emu():
    src = ROOT / "tests" / "emu" / "emu.cpp"
    so = ROOT / "tests" / "emu" / "libpbrl_emu.so"
    deps = [src] + list((ROOT / PKG / "csrc").glob("*.h")) + list((ROOT / PKG / "csrc").glob("*.cuh")) + [ROOT / "include" / "pbrl.h"]
    if not so.exists() or so.stat().st_mtime < max(d.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(so), str(src)], check=True)
    return so


class EmuBackend:
    """Kernel bodies compiled for the host (tests/emu/emu.cpp).  Logic check only."""
    name = "emu"

    def __init__(self, model):
        self.lib = ctypes.CDLL(str(build_emu()))
        self.lib.emu_last_error.restype = ctypes.c_char_p
        self._desc = desc.build_model_desc(model, self.lib.emu_model_desc_size, self.lib.emu_model_desc_field)
        self._select()

    def _select(self):
        """The emulation library holds one model at a time: (re)select ours before every call."""
        if self.lib.emu_set_model(self._desc.ctypes.data_as(ctypes.c_void_p)):
            raise RuntimeError(self.lib.emu_last_error().decode())
        if getattr(self, "_mimic", None) is not None and self.lib.emu_set_mimic(ctypes.byref(self._mimic)):
            raise RuntimeError(self.lib.emu_last_error().decode())

    def new_state(self, N):
        return _HostState(N)

    @staticmethod
    def _bufs(s):
        b = desc.PbrlBuffers()
        for n, _ in desc.PbrlBuffers._fields_:
            if hasattr(s, n):
                setattr(b, n, getattr(s, n).ctypes.data)
        return b

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data_as(ctypes.c_void_p)

    def set_mimic(self, mdesc):
        self._mimic = mdesc
        self._select()

    def reset(self, s, reward_mode=0):
        self._select()
        b = self._bufs(s)
        self.lib.emu_reset(s.N, reward_mode, ctypes.byref(b))

    def autoreset(self, s, done, next_ref_t0=None, steps=None, reward_mode=0):
        self._select()
        b = self._bufs(s)
        done = np.ascontiguousarray(done, np.float32)
        nxt = None if next_ref_t0 is None else np.ascontiguousarray(next_ref_t0, np.float32)
        self.lib.emu_autoreset(s.N, reward_mode, self._p(done), self._p(nxt), self._p(steps), ctypes.byref(b))

    def step(self, s, action, kp, kd, pd_mode=3, pd_dt=0.002, n_frames=8, pd_time=None, ctrl_mask=0x0FFFFFFF, reward_mode=0):
        self._select()
        a = np.ascontiguousarray(action, np.float32)
        cols = 1 if a.ndim == 2 else a.shape[2]
        kp, kd = np.ascontiguousarray(kp, np.float32), np.ascontiguousarray(kd, np.float32)
        t = None if pd_time is None else np.ascontiguousarray(pd_time, np.float32)
        b = self._bufs(s)
        self.lib.emu_step(s.N, n_frames, pd_mode, cols, self._p(a), self._p(kp), self._p(kd), ctypes.c_float(pd_dt), self._p(t),
                          ctypes.c_uint(ctrl_mask), reward_mode, ctypes.byref(b))

    def debug_forward(self, qpos, qvel, ctrl, warm):
        self._select()
        A = [np.ascontiguousarray(x, np.float32) for x in (qpos, qvel, ctrl, warm)]
        out = np.zeros((A[0].shape[0], DBG_SIZE), np.float32)
        self.lib.emu_debug_forward(A[0].shape[0], *[self._p(x) for x in A], self._p(out))
        return out

    def pd_torque(self, s, action, kp, kd, pd_mode, pd_dt, time=None, want_qM=False):
        self._select()
        a = np.ascontiguousarray(action, np.float32)
        cols = 1 if a.ndim == 2 else a.shape[2]
        kp, kd = np.ascontiguousarray(kp, np.float32), np.ascontiguousarray(kd, np.float32)
        t = np.ascontiguousarray(s.time if time is None else time, np.float32)
        tau = np.zeros((s.N, 28), np.float32)
        qM = np.zeros((s.N, 34, 34), np.float32) if want_qM else None
        bias = np.zeros((s.N, 34), np.float32) if want_qM else None
        self.lib.emu_pd_torque(s.N, pd_mode, cols, self._p(a), self._p(s.qpos), self._p(s.qvel), self._p(s.qpos_lag),
                               self._p(s.qvel_lag), self._p(kp), self._p(kd), self._p(t), ctypes.c_float(pd_dt), self._p(tau),
                               self._p(qM), self._p(bias))
        return (tau, qM, bias) if want_qM else tau

    def replay_item_fk(self, qpos):
        """The per-item kinematics of the replay-rollout kernel (one thread per item on the GPU)."""
        self._select()
        qpos = np.ascontiguousarray(qpos, np.float32)
        xpos, xquat = np.zeros((qpos.shape[0], 14, 3), np.float32), np.zeros((qpos.shape[0], 14, 4), np.float32)
        for i in range(qpos.shape[0]):
            self.lib.emu_replay_item_fk(self._p(qpos[i]), self._p(xpos[i]), self._p(xquat[i]))
        return xpos, xquat

    def replay(self, qpos, qvel, want_poses=True):
        self._select()
        qpos, qvel = np.ascontiguousarray(qpos, np.float32), np.ascontiguousarray(qvel, np.float32)
        N = qpos.shape[0]
        obs, reward, done = np.zeros((N, 69), np.float32), np.ones((N, 3), np.float32), np.zeros(N, np.float32)
        xpos, xquat = np.zeros((N, 14, 3), np.float32), np.zeros((N, 14, 4), np.float32)
        self.lib.emu_replay(N, self._p(qpos), self._p(qvel), self._p(obs), self._p(reward), self._p(done),
                            self._p(xpos) if want_poses else None, self._p(xquat) if want_poses else None)
        return obs, reward, done, xpos, xquat


class GpuBackend:
    """libpbrl.so through the C ABI, device buffers held as torch tensors."""
    name = "gpu"

    def __init__(self, model, device=0, mimic_desc=None):
        import torch
        self.torch = torch
        self.L = importlib.import_module(PKG + "._lib")
        self.model = self.L.Model(model, device=device, mimic=mimic_desc)
        self.dev = torch.device("cuda", device)

    class _DevState:
        pass

    def new_state(self, N):
        s = _HostState(N)
        return s

    def _up(self, s):
        t = self.torch
        d = {n: t.from_numpy(getattr(s, n)).to(self.dev) for n, _ in STATE_SHAPES}
        for n in ("status", "cost", "motion_id", "ref_t0", "stats"):
            d[n] = t.from_numpy(getattr(s, n)).to(self.dev)
        return d

    def _down(self, s, d):
        for n in d:
            getattr(s, n)[...] = d[n].cpu().numpy()

    def _dp(self, a, dtype=np.float32):
        return None if a is None else self.torch.from_numpy(np.ascontiguousarray(a, dtype)).to(self.dev)

    def _stream(self):
        return self.L.current_stream_ptr(self.dev)

    def reset(self, s, reward_mode=0):
        d = self._up(s)
        b = self.L.make_buffers(**d)
        self.L.check(self.L.lib().pbrl_reset(self.model.handle, s.N, reward_mode, ctypes.byref(b), self._stream()))
        self.torch.cuda.synchronize()
        self._down(s, d)

    def autoreset(self, s, done, next_ref_t0=None, steps=None, reward_mode=0):
        d = self._up(s)
        b = self.L.make_buffers(**d)
        dn, nx = self._dp(done), self._dp(next_ref_t0)
        st = None if steps is None else self.torch.from_numpy(steps).to(self.dev)
        P = self.L._ptr
        self.L.check(self.L.lib().pbrl_autoreset(self.model.handle, s.N, reward_mode, P(dn), P(nx), P(st), ctypes.byref(b), self._stream()))
        self.torch.cuda.synchronize()
        self._down(s, d)
        if steps is not None:
            steps[...] = st.cpu().numpy()

    def step(self, s, action, kp, kd, pd_mode=3, pd_dt=0.002, n_frames=8, pd_time=None, ctrl_mask=0x0FFFFFFF, reward_mode=0):
        a = np.ascontiguousarray(action, np.float32)
        cols = 1 if a.ndim == 2 else a.shape[2]
        d = self._up(s)
        b = self.L.make_buffers(**d)
        ad, kpd, kdd, td = self._dp(a), self._dp(kp), self._dp(kd), self._dp(pd_time)
        P = self.L._ptr
        self.L.check(self.L.lib().pbrl_step(self.model.handle, s.N, n_frames, pd_mode, cols, P(ad), P(kpd), P(kdd),
                                            ctypes.c_float(pd_dt), P(td), ctypes.c_uint(ctrl_mask), reward_mode, ctypes.byref(b),
                                            self._stream()))
        self.torch.cuda.synchronize()
        self._down(s, d)

    def debug_forward(self, qpos, qvel, ctrl, warm):
        A = [self._dp(x) for x in (qpos, qvel, ctrl, warm)]
        N = A[0].shape[0]
        out = self.torch.zeros((N, DBG_SIZE), device=self.dev)
        P = self.L._ptr
        self.L.check(self.L.lib().pbrl_debug_forward(self.model.handle, N, P(A[0]), P(A[1]), P(A[2]), P(A[3]), P(out), self._stream()))
        self.torch.cuda.synchronize()
        return out.cpu().numpy()

    def pd_torque(self, s, action, kp, kd, pd_mode, pd_dt, time=None, want_qM=False):
        a = np.ascontiguousarray(action, np.float32)
        cols = 1 if a.ndim == 2 else a.shape[2]
        t = self.torch
        tau = t.zeros((s.N, 28), device=self.dev)
        qM = t.zeros((s.N, 34, 34), device=self.dev) if want_qM else None
        bias = t.zeros((s.N, 34), device=self.dev) if want_qM else None
        P = self.L._ptr
        args = [self._dp(x) for x in (a, s.qpos, s.qvel, s.qpos_lag, s.qvel_lag, kp, kd, s.time if time is None else time)]
        self.L.check(self.L.lib().pbrl_pd_torque(self.model.handle, s.N, pd_mode, cols, *[P(x) for x in args],
                                                 ctypes.c_float(pd_dt), P(tau), P(qM), P(bias), self._stream()))
        t.cuda.synchronize()
        return (tau.cpu().numpy(), qM.cpu().numpy(), bias.cpu().numpy()) if want_qM else tau.cpu().numpy()

    def replay_rollout(self, data_pos, data_vel, frame0, N, K):
        t = self.torch
        dp, dv, f0 = self._dp(data_pos), self._dp(data_vel), self._dp(frame0, np.int32)
        z = lambda *s_: t.full(s_, -7.0, device=self.dev)
        out = dict(obs=z(K, N, 69), reward=z(K, N, 3), done=z(K, N), xpos=z(K, N, 14, 3), xquat=z(K, N, 14, 4))
        P = self.L._ptr
        self.L.check(self.L.lib().pbrl_replay_rollout(self.model.handle, N, K, P(dp), P(dv), dp.shape[0], P(f0), P(out["obs"]),
                                                      P(out["reward"]), P(out["done"]), P(out["xpos"]), P(out["xquat"]), self._stream()))
        t.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items()}

    def replay(self, qpos, qvel, want_poses=True):
        t = self.torch
        qp, qv = self._dp(qpos), self._dp(qvel)
        N = qp.shape[0]
        obs, reward, done = t.zeros((N, 69), device=self.dev), t.ones((N, 3), device=self.dev), t.zeros(N, device=self.dev)
        xpos, xquat = t.zeros((N, 14, 3), device=self.dev), t.zeros((N, 14, 4), device=self.dev)
        P = self.L._ptr
        self.L.check(self.L.lib().pbrl_replay(self.model.handle, N, P(qp), P(qv), P(obs), P(reward), P(done),
                                              P(xpos) if want_poses else None, P(xquat) if want_poses else None, self._stream()))
        t.cuda.synchronize()
        return tuple(x.cpu().numpy() for x in (obs, reward, done, xpos, xquat))
