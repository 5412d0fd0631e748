"""Fill the C-ABI `PbrlModelDesc` (include/pbrl.h) from a ModelConst, by field name."""
import ctypes

import numpy as np

MODEL_SCALARS = ["nbody", "nq", "nv", "nu", "njnt", "ngeom", "iterations", "ls_iterations", "timestep",
                 "impratio", "tolerance", "ls_tolerance", "meaninertia", "npair", "solver_rule"]
MODEL_ARRAYS = ["gravity", "solref", "solimp", "body_parent", "body_pos", "body_mass", "body_ipos",
                "body_inertia", "body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum", "body_invweight0",
                "jnt_type", "jnt_bodyid", "jnt_qposadr", "jnt_dofadr", "jnt_axis", "jnt_pos", "jnt_limited",
                "jnt_range", "jnt_margin", "dof_bodyid", "dof_parentid", "dof_armature", "dof_damping",
                "dof_invweight0", "qpos0", "geom_type", "geom_bodyid", "geom_size", "geom_pos", "geom_quat",
                "geom_friction", "geom_margin", "geom_gap", "act_dofadr", "act_gear", "act_ctrllimited",
                "act_ctrlrange", "pair_geom1", "pair_geom2"]


def build_model_desc(model, size_fn, field_fn):
    """Returns a uint8 numpy buffer holding a PbrlModelDesc.  `size_fn`/`field_fn` are the library's
    `*_model_desc_size` / `*_model_desc_field` entry points."""
    buf = np.zeros(size_fn(), dtype=np.uint8)
    for name in MODEL_SCALARS + MODEL_ARRAYS:
        off, cnt, isint = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        if field_fn(name.encode(), ctypes.byref(off), ctypes.byref(cnt), ctypes.byref(isint)):
            raise KeyError(name)
        dt = np.int32 if isint.value else np.float64
        if name == "npair":  # collision candidates: all of them unless the model opts out (self_collision = 0 keeps the floor pairs only)
            val = len(model.pair_geom1) if getattr(model, "pair_geom1", None) is not None and getattr(model, "self_collision", 1) else 0
        elif name in ("pair_geom1", "pair_geom2"):
            val = getattr(model, name, None)
            val = np.zeros(0, np.int32) if val is None else val
        elif name == "solver_rule":  # 0 exact active set (default), 1 MuJoCo / MJX tolerance rule (include/pbrl.h)
            val = int(getattr(model, "solver_rule", 0))
        else:
            val = getattr(model, name)
        v = np.asarray(val, dtype=dt).ravel()
        if v.size > cnt.value:
            raise ValueError(f"{name}: {v.size} values exceed the capacity {cnt.value}")
        buf[off.value:off.value + cnt.value * np.dtype(dt).itemsize].view(dt)[:v.size] = v
    return buf


class PbrlBuffers(ctypes.Structure):
    """Mirror of `PbrlBuffers` in include/pbrl.h (device pointers)."""
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("qpos", "qvel", "time", "warm", "ctrl", "qpos_lag", "qvel_lag", "obs", "reward", "done", "metrics",
                 "status", "motion_id", "ref_t0", "stats", "env_order", "cost")]
