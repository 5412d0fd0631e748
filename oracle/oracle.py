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

    def __init__(self, model, precision="f64", tolerance=None, iterations=None):
        build()
        self.lib = ctypes.CDLL(str(HERE / f"liboracle_{precision}.so"))
        self.real = np.float64 if precision == "f64" else np.float32
        assert self.lib.orc_real_size() == np.dtype(self.real).itemsize
        self.m = model
        self.nq, self.nv, self.nu, self.nbody = model.nq, model.nv, model.nu, model.nbody
        self._model_buf = np.zeros(self.lib.orc_model_size(), dtype=np.uint8)
        self._fill_model(model, tolerance, iterations)
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
