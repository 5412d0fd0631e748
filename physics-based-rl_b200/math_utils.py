"""Cubic point-to-point joint trajectories used as stable-PD targets.

Same functions as /root/reference/some_math/math_utils.py:75-121 (numpy instead of jax).  The mocap half of that module
(`align_position`, `align_rotation`, `calc_rot_vel`, :15-65) lives in mocap.py and is re-exported here, so that
`from math_utils import *` finds what the reference's module offers.
"""
import numpy as np

from .mocap import align_position, align_rotation, calc_rot_vel  # noqa: F401  (re-exports)


def generate_trajectory(t0, tf, q0, qf):
    """Coefficients [a0,a1,a2,a3] of the cubic with q(t0)=q0, q(tf)=qf and zero end velocities."""
    d = float(tf - t0) ** 3
    a0 = (qf * t0 * t0 * (3 * tf - t0) + q0 * tf * tf * (tf - 3 * t0)) / d
    a1 = 6.0 * t0 * tf * (q0 - qf) / d
    a2 = 3.0 * (t0 + tf) * (qf - q0) / d
    a3 = 2.0 * (q0 - qf) / d
    return np.array([a0, a1, a2, a3], dtype=np.float32)


def start_trajectories(trajectory_dict, use_dummy=True, nu=28):
    """(nu,4) table: all-zero cubics, rows in `trajectory_dict` overwritten (math_utils.py:94-111)."""
    traj = np.tile(generate_trajectory(1, 3, 0, 0), (nu, 1))
    for k, v in trajectory_dict.items():
        traj[int(k)] = np.asarray(v, dtype=np.float32)
    return traj


def compute_cubic_trajectory(time, trajectory):
    """q_hat_i(t) = a0 + a1 t + a2 t^2 + a3 t^3 for every actuator row (math_utils.py:114-121)."""
    t = np.asarray(time, dtype=trajectory.dtype)
    return trajectory[..., 0] + trajectory[..., 1] * t + trajectory[..., 2] * t ** 2 + trajectory[..., 3] * t ** 3
