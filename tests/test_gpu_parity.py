"""Parity tests proper: the CUDA path (libpbrl.so through the C ABI) against the oracle on a B200."""
import pytest

import parity_cases as pc
from helpers import GpuBackend

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu_nograv(model_nograv):
    return GpuBackend(model_nograv)


@pytest.fixture(scope="module")
def gpu_grav(model_grav):
    return GpuBackend(model_grav)


def test_forward_stages_limits(gpu_nograv, model_nograv):
    pc.check_forward_stages(gpu_nograv, model_nograv, gravity=False, N=64)


def test_forward_stages_contacts(gpu_grav, model_grav):
    pc.check_forward_stages(gpu_grav, model_grav, gravity=True, N=64)


def test_one_substep(gpu_nograv, model_nograv, gpu_grav, model_grav):
    pc.check_one_substep(gpu_nograv, model_nograv, gravity=False, N=256)
    pc.check_one_substep(gpu_grav, model_grav, gravity=True, N=256)


@pytest.mark.parametrize("pd_mode", [0, 1, 2, 3])
def test_32_contact_free_substeps(gpu_nograv, model_nograv, pd_mode):
    pc.check_32_substeps(gpu_nograv, model_nograv, pd_mode=pd_mode, N=128)


def test_pd_laws(gpu_nograv, model_nograv):
    pc.check_pd_torque(gpu_nograv, model_nograv, N=64)


def test_golden_T1(gpu_nograv, model_nograv):
    pc.check_golden_T1(gpu_nograv, model_nograv)


def test_golden_T2_T4(gpu_nograv, model_nograv):
    pc.check_golden_T2_T4(gpu_nograv, model_nograv)


def test_replay(gpu_nograv, model_nograv):
    pc.check_replay(gpu_nograv, model_nograv, N=4096)


def test_reset(gpu_nograv, model_nograv):
    pc.check_reset(gpu_nograv, model_nograv)


def test_hands_against_thighs(gpu_nograv, model_nograv):
    """Body-vs-body contacts (sphere / capsule candidates) on the CUDA path, dense Hessian branch included."""
    pc.check_hands_against_thighs(gpu_nograv, model_nograv, N=48, steps=12)


def test_contact_rollout(gpu_grav, model_grav):
    pc.check_contact_rollout(gpu_grav, model_grav, N=64, steps=40)


def test_mimic_mode(model_grav):
    table, mdesc, keep = pc.mimic_setup(model_grav)
    b = GpuBackend(model_grav, mimic_desc=mdesc)
    pc.check_mimic(b, model_grav, table, mdesc, N=112)


def test_autoreset(model_grav):
    table, mdesc, keep = pc.mimic_setup(model_grav)
    pc.check_autoreset(GpuBackend(model_grav, mimic_desc=mdesc), model_grav, table, mdesc, N=133)


def test_edge_sizes_and_full_size_batch_invariance(gpu_grav, model_grav):
    """Empty, single, ragged and maximum batches.  At BASELINE's full size (65 536 envs) the result of an env must not
    depend on the batch around it: 1 024 copies of 64 distinct envs give 1 024 bit-identical groups, equal to the
    64-env launch and -- with a permuted processing order (`env_order`) -- to themselves."""
    import numpy as np
    import torch
    import ctypes
    kp, kd = pc.util_data.generate_kp_kd_gains()
    be = gpu_grav
    L = be.L
    lo, hi = model_grav.jnt_range[1:, 0], model_grav.jnt_range[1:, 1]
    # empty batch: a no-op that reports success
    assert L.lib().pbrl_step(be.model.handle, 0, 8, 3, 1, None, None, None, ctypes.c_float(0.002), None, ctypes.c_uint(0), 0, None, None) == 0
    assert L.lib().pbrl_reset(be.model.handle, 0, 0, None, None) == 0
    base_q, base_v = pc.random_states(model_grav, 64, 41, gravity_heights=True)
    rng = np.random.default_rng(41)
    act = rng.uniform(lo, hi, (64, 28)).astype(np.float32)

    def run(N, order=None):
        reps = -(-N // 64)
        s = be.new_state(N)
        s.set_state(np.tile(base_q, (reps, 1))[:N].astype(np.float32), np.tile(base_v, (reps, 1))[:N].astype(np.float32))
        if order is None:
            for _ in range(2):
                be.step(s, np.tile(act, (reps, 1))[:N], kp, kd, pd_mode=3, n_frames=8)
            return s
        d = be._up(s)
        d["env_order"] = torch.from_numpy(order.astype(np.int32)).to(be.dev)
        a = torch.from_numpy(np.tile(act, (reps, 1))[:N]).to(be.dev)
        kpd, kdd = be._dp(kp), be._dp(kd)
        b = L.make_buffers(**d)
        for _ in range(2):
            L.check(L.lib().pbrl_step(be.model.handle, N, 8, 3, 1, L._ptr(a), L._ptr(kpd), L._ptr(kdd), ctypes.c_float(0.002), None,
                                      ctypes.c_uint(0x0FFFFFFF), 0, ctypes.byref(b), be._stream()))
        torch.cuda.synchronize()
        be._down(s, {k: v for k, v in d.items() if k != "env_order"})
        return s

    ref = run(64)
    assert np.isfinite(ref.qpos).all() and (ref.status & 7 == 0).all()
    for N in (1, 17, 2369):                                  # single env, ragged CTA, one more than a full wave (148 x 16)
        s = run(N)
        n = min(N, 64)
        assert np.array_equal(s.qpos[:n], ref.qpos[:n]) and np.array_equal(s.qvel[:n], ref.qvel[:n]), N
    N = 65536
    s = run(N)
    q = s.qpos.reshape(N // 64, 64, 35)
    assert (q == q[:1]).all() and np.array_equal(q[0], ref.qpos)
    assert np.allclose(np.linalg.norm(s.qpos[:, 3:7], axis=1), 1.0, atol=1e-5)
    sp = run(N, order=np.random.default_rng(5).permutation(N))
    assert np.array_equal(sp.qpos, s.qpos) and np.array_equal(sp.qvel, s.qvel) and np.array_equal(sp.warm, s.warm)


def test_replay_rollout(gpu_nograv, model_nograv):
    pc.check_replay_rollout(gpu_nograv, model_nograv)
    pc.check_replay_rollout(gpu_nograv, model_nograv, F=227, N=4096, K=3, seed=38)   # several CTAs per step, ragged tail
    pc.check_replay_rollout(gpu_nograv, model_nograv, F=5, N=1, K=1, seed=39)


def test_contact_rich_statistics(model_grav):
    table, mdesc, keep = pc.mimic_setup(model_grav)
    pc.check_contact_rich_statistics(GpuBackend(model_grav, mimic_desc=mdesc), model_grav, table, mdesc)


def test_sphere_box_contacts(gpu_grav, model_grav):
    pc.check_sphere_box_contacts(gpu_grav, model_grav, want=30)
