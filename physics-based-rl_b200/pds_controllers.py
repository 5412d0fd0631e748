"""PD controllers with the reference's signatures, evaluated by the CUDA library.

  *_mjx variants  /root/reference/agents_env/pds_controllers_mjx.py:17-32, 40-80, 156-191
                  fn(target, mj_model, mj_data, mjx_data, q, qdot, kp_array, kd_array, time, dt) -> tau[N,28]
  brax variants   /root/reference/agents_env/pds_controllers_agents.py:17-34, 42-78, 137-172
                  fn(target, sys, state, q, qdot, kp_array, kd_array, time, dt) -> tau[N,28]

Semantics follow pds_controllers_agents.py (live M / bias of the last forward); the *_mjx twins keep
their signature but not the stale-host-M quirk.  Each function carries a
`pd_mode` attribute so that HumanoidTrain can fuse the law into the step kernel.
"""
import ctypes

import torch

from . import _lib


def _evaluate(mode, target, env_model, pipeline_state, q, qdot, kp, kd, time, dt):
    dev = q.device
    N = q.shape[0]
    target = torch.as_tensor(target, dtype=torch.float32, device=dev)
    if tuple(target.shape) in ((28,), (28, 4)):
        target = target.expand(N, *target.shape)
    # the kernel reads N x 28 x cols floats through a raw pointer: refuse anything else before the launch
    if tuple(target.shape) not in ((N, 28), (N, 28, 4)):
        raise ValueError(f"target: expected shape ({N}, 28) or ({N}, 28, 4) [cubic coefficients], got {tuple(target.shape)}")
    target = target.contiguous()
    cols = 1 if target.dim() == 2 else 4
    kp = torch.as_tensor(kp, dtype=torch.float32, device=dev).contiguous()
    kd = torch.as_tensor(kd, dtype=torch.float32, device=dev).contiguous()
    if kp.numel() != 28 or kd.numel() != 28:
        raise ValueError("kp_array / kd_array: expected 28 gains each")
    time = torch.as_tensor(time, dtype=torch.float32, device=dev)
    if time.numel() not in (1, N):
        raise ValueError(f"time: expected a scalar or {N} values, got {tuple(time.shape)}")
    time = time.reshape(-1).expand(N).contiguous()
    if tuple(q.shape) != (N, 35) or tuple(qdot.shape) != (N, 34) or tuple(pipeline_state.qpos_lag.shape) != (N, 35) or \
            tuple(pipeline_state.qvel_lag.shape) != (N, 34):
        raise ValueError("q / qdot / lagged state: expected (N, 35) / (N, 34) rows for the same N")
    tau = torch.empty((N, 28), device=dev)
    _lib.check(_lib.lib().pbrl_pd_torque(
        env_model.handle, ctypes.c_int(N), ctypes.c_int(mode), ctypes.c_int(cols), _lib._ptr(target),
        _lib._ptr(q.contiguous()), _lib._ptr(qdot.contiguous()), _lib._ptr(pipeline_state.qpos_lag),
        _lib._ptr(pipeline_state.qvel_lag), _lib._ptr(kp), _lib._ptr(kd), _lib._ptr(time), ctypes.c_float(dt),
        _lib._ptr(tau), None, None, _lib.current_stream_ptr(dev)))
    return tau


def _brax(mode):
    def fn(target, sys, state, q, qdot, kp_array, kd_array, time, dt):
        ps = state.pipeline_state
        return _evaluate(mode, target, ps._env.model, ps, q, qdot, kp_array, kd_array, time, dt)
    fn.pd_mode = mode
    return fn


def _mjx(mode):
    def fn(target, mj_model, mj_data, mjx_data, q, qdot, kp_array, kd_array, time, dt):
        return _evaluate(mode, target, mjx_data._env.model, mjx_data, q, qdot, kp_array, kd_array, time, dt)
    fn.pd_mode = mode
    return fn


standard_pd_controller = _brax(1)
feedback_pd_controller = _brax(2)
stable_pd_controller_custom_trajectory = _brax(3)
standard_pd_controller_mjx = _mjx(1)
feedback_pd_controller_mjx = _mjx(2)
stable_pd_controller_custom_trajectory_mjx = _mjx(3)


# ------------------------------------------------------------------------------------------------------------
# The building blocks of the reference's stable PD, for callers that assemble the law themselves
# (/root/reference/agents_env/pds_controllers_agents.py:83-134, pds_controllers_mjx.py:85-153).  They are
# host-side conveniences over device tensors (batched: a leading env dimension everywhere), not the product
# path: stable_pd_controller_custom_trajectory above evaluates the whole law inside one kernel, with the
# tree-sparse factorisation instead of a dense one.  Same argument order and return order as the reference.
def init_corolis_mass_external_diagonal(*args):
    """(state, kp_array, kd_array) as pds_controllers_agents.py:83, or (mj_model, mj_data, mjx_data, kp_array, kd_array)
    as pds_controllers_mjx.py:85 -> C [N,34] (qfrc_bias), M [N,34,34] (dense qM), KP, KD (diagonal matrices of the gains as
    given: the reference passes them padded to 34), tau_ext [N,34] (qfrc_applied, zeros)."""
    if len(args) == 3:
        ps, kp_array, kd_array = args[0].pipeline_state, args[1], args[2]
    elif len(args) == 5:
        ps, kp_array, kd_array = args[2], args[3], args[4]
    else:
        raise TypeError("init_corolis_mass_external_diagonal(state, kp, kd) or (mj_model, mj_data, mjx_data, kp, kd)")
    M, C = ps._lag_dynamics()   # one launch for both
    dev = C.device
    KP = torch.diag(torch.as_tensor(kp_array, dtype=torch.float32, device=dev).reshape(-1))
    KD = torch.diag(torch.as_tensor(kd_array, dtype=torch.float32, device=dev).reshape(-1))
    return C, M, KP, KD, ps.qfrc_applied


def calculate_new_mass(M, KD, dt):
    """M + KD dt (pds_controllers_agents.py:99-104; the reference also prints M)."""
    return M + KD * dt


def calculate_new_mass_check_if_invertible(M, KD, dt, time):
    """pds_controllers_mjx.py:104-127: M + KD dt, plus 1e-6 I when time == 1, and a warning when the condition number
    reaches 1 / eps of the dtype."""
    new_mass = M + KD * dt
    t = torch.as_tensor(time)
    if t.numel() == 1 and float(t) == 1:
        new_mass = new_mass + 1e-6 * torch.eye(new_mass.shape[-1], dtype=new_mass.dtype, device=new_mass.device)
    cond = torch.linalg.cond(new_mass)
    if bool((cond >= 1.0 / torch.finfo(new_mass.dtype).eps).any()):
        print(f"Warning: High condition number ({float(cond.max())}) indicates potential numerical instability. current time {time}")
    return new_mass


def compute_acceleration(q_error, qdot, C, tau_ext, new_mass, KP, KD):
    """new_mass^-1 (-C - KP q_error - KD qdot + tau_ext) by a Cholesky solve (pds_controllers_agents.py:106-134); vectors
    are [N,34] or [34], matrices [N,34,34] or [34,34]."""
    rhs = -C - q_error @ KP.transpose(-1, -2) - qdot @ KD.transpose(-1, -2) + tau_ext
    L = torch.linalg.cholesky(new_mass)
    return torch.cholesky_solve(rhs.unsqueeze(-1), L).squeeze(-1)
