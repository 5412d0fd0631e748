"""Kernel LOGIC vs the oracle without a GPU: the kernel bodies of pbrl_core.cuh compiled for the host
(tests/emu/emu.cpp).  The same cases run on the real CUDA path in test_gpu_parity.py."""
import pytest

import parity_cases as pc
from helpers import EmuBackend


@pytest.fixture(scope="module")
def emu_nograv(model_nograv):
    return EmuBackend(model_nograv)


@pytest.fixture(scope="module")
def emu_grav(model_grav):
    return EmuBackend(model_grav)


def test_forward_stages_limits(emu_nograv, model_nograv):
    pc.check_forward_stages(emu_nograv, model_nograv, gravity=False)


def test_forward_stages_contacts(emu_grav, model_grav):
    pc.check_forward_stages(emu_grav, model_grav, gravity=True)


def test_one_substep(emu_nograv, model_nograv, emu_grav, model_grav):
    pc.check_one_substep(emu_nograv, model_nograv, gravity=False)
    pc.check_one_substep(emu_grav, model_grav, gravity=True)


@pytest.mark.parametrize("pd_mode", [0, 1, 2, 3])
def test_32_contact_free_substeps(emu_nograv, model_nograv, pd_mode):
    pc.check_32_substeps(emu_nograv, model_nograv, pd_mode=pd_mode, N=8)


def test_pd_laws(emu_nograv, model_nograv):
    pc.check_pd_torque(emu_nograv, model_nograv)


def test_golden_T1(emu_nograv, model_nograv):
    pc.check_golden_T1(emu_nograv, model_nograv)


def test_golden_T2_T4(emu_nograv, model_nograv):
    pc.check_golden_T2_T4(emu_nograv, model_nograv)


def test_replay(emu_nograv, model_nograv):
    pc.check_replay(emu_nograv, model_nograv)


def test_reset(emu_nograv, model_nograv):
    pc.check_reset(emu_nograv, model_nograv)


def test_contact_rollout(emu_grav, model_grav):
    pc.check_contact_rollout(emu_grav, model_grav, N=8, steps=30)


def test_mimic_mode(model_grav):
    table, mdesc, keep = pc.mimic_setup(model_grav)
    b = EmuBackend(model_grav)
    b.set_mimic(mdesc)
    pc.check_mimic(b, model_grav, table, mdesc)


def test_autoreset(model_grav):
    table, mdesc, keep = pc.mimic_setup(model_grav)
    b = EmuBackend(model_grav)
    b.set_mimic(mdesc)
    pc.check_autoreset(b, model_grav, table, mdesc)


def test_replay_item_kinematics(emu_nograv, model_nograv):
    """The thread-serial FK of the replay-rollout kernel against the oracle (same poses as stage_fk)."""
    import numpy as np
    from helpers import Oracle, OracleEnv, random_states, relerr
    qpos, qvel = random_states(model_nograv, 40, 37, vel_scale=1.0, joint_frac=0.9)
    qpos[0, 3:7] *= 0.99
    q32 = qpos.astype(np.float32)
    xpos, xquat = emu_nograv.replay_item_fk(q32)
    _, _, _, xpos_r, xquat_r = OracleEnv(Oracle(model_nograv, "f64"), 40).replay(q32, qvel.astype(np.float32), want_poses=True)
    assert relerr(xpos, xpos_r) <= 2e-6 and relerr(xquat, xquat_r) <= 2e-6


def test_contact_rich_statistics_small(model_grav):
    table, mdesc, keep = pc.mimic_setup(model_grav)
    b = EmuBackend(model_grav)
    b.set_mimic(mdesc)
    pc.check_contact_rich_statistics(b, model_grav, table, mdesc, N=70, steps=10)


def test_forward_stages_general_solimp_power(model_grav):
    """A solimp power other than MuJoCo's default 2 takes the out-of-line powf path of the impedance."""
    import copy
    import numpy as np
    m = copy.copy(model_grav)
    m.solimp = np.array([0.9, 0.95, 0.001, 0.4, 3.0])
    pc.check_forward_stages(EmuBackend(m), m, gravity=True, N=12, seed=9)


def test_hovering_foot_boxes(emu_grav, model_grav):
    pc.check_hovering_foot_boxes(emu_grav, model_grav)


def test_sphere_box_contacts(emu_grav, model_grav):
    pc.check_sphere_box_contacts(emu_grav, model_grav)


def test_contact_list_overflow_keeps_the_deepest(model_grav):
    """More than 8 body-vs-body contacts at once (absurdly folded random poses): the 8 deepest are kept and the status says so."""
    import numpy as np
    from helpers import DBG, Oracle, random_states
    orc = Oracle(model_grav, "f64")
    be = EmuBackend(model_grav)
    found = 0
    for seed in range(300, 340):
        qpos, qvel = random_states(model_grav, 32, seed, gravity_heights=True, vel_scale=0.0, joint_frac=1.0)
        for e in range(32):
            d = orc.new_data()
            d.qpos[:35] = qpos[e].astype(np.float32)
            orc.forward(d)
            nc, ns = int(d.ncon[0]), int(d.nself[0])
            if ns <= 8:
                continue
            out = be.debug_forward(qpos[e:e + 1].astype(np.float32), np.zeros((1, 34), np.float32), np.zeros((1, 28), np.float32), np.zeros((1, 34), np.float32))
            kept = np.sort(out[0, DBG["self_dist"][0]:DBG["self_dist"][0] + 8])
            want = np.sort(d.con_dist[nc - ns:nc])[:8]
            assert int(out[0, DBG["nself"][0]]) == 8 and np.abs(kept - want).max() <= 2e-6, (seed, e, kept, want)
            found += 1
        if found >= 3:
            break
    assert found >= 2


def test_config3_punch_tracking_small(model_nograv):
    """BASELINE configs[2] by name (no-gravity model, punch clip, clip-tracking PD targets), small batch on the host emulation."""
    table, mdesc, keep, ids = pc.named_config_setup(model_nograv, ["punch"])
    b = EmuBackend(model_nograv)
    b.set_mimic(mdesc)
    pc.check_config3_punch_tracking(b, model_nograv, table, mdesc, ids, N=12)


def test_config4_backflip_jump_small(model_grav):
    """BASELINE configs[3] by name (backflip on even env ids, jump on odd ones), small batch on the host emulation."""
    table, mdesc, keep, ids = pc.named_config_setup(model_grav, ["backflip", "jump"])
    b = EmuBackend(model_grav)
    b.set_mimic(mdesc)
    pc.check_config4_backflip_jump(b, model_grav, table, mdesc, ids, N=12, N_stat=48, steps=8)


def test_alt_rate_1200hz_40_frames(model_nograv, model_grav):
    """agents_env/agent_env_template2.py:49-58: timestep 1/1200, n_frames 40."""
    pc.check_alt_rate(lambda m: EmuBackend(m), model_nograv, N=6)
    pc.check_alt_rate(lambda m: EmuBackend(m), model_grav, N=6)


def test_solver_rule_mjx_and_options(model_grav):
    """PbrlModelDesc.iterations / tolerance / ls_iterations / ls_tolerance are honoured under solver_rule = 1 (MJX's stop rule)."""
    pc.check_solver_rule(lambda m: EmuBackend(m), model_grav, N=32)


def test_capsule_box_and_box_box_contacts(emu_grav, model_grav):
    """The 15 candidates against the foot boxes: same contacts and forces as the oracle."""
    pc.check_box_pair_contacts(emu_grav, model_grav, want=60)


def test_quiet_stance_carries_the_body_weight(emu_grav, model_grav):
    """Contact self-consistency (stand-in for the missing MJX contact goldens): weight support within 5 %, pyramid edge forces
    >= 0 (friction cone), ~zero normal velocity, sub-millimetre penetration -- oracle and kernel logic."""
    pc.check_quiet_stance(emu_grav, model_grav)


def test_step_in_two_parts_is_bit_identical(model_grav):
    """env_step over a substep range: PD law + r substeps, then the other substeps + epilogue (what pbrl_step launches when it
    regroups the envs inside the step) gives the bits of the whole step, in reference and in mimic mode."""
    import os
    import numpy as np
    table, mdesc, keep = pc.mimic_setup(model_grav)
    N = 24
    rng = np.random.default_rng(12)
    kp, kd = pc.util_data.generate_kp_kd_gains()
    lo, hi = model_grav.jnt_range[1:, 0], model_grav.jnt_range[1:, 1]
    mid = (np.arange(N) % 7).astype(np.int32)
    t0 = (rng.uniform(0, 1, N) * table.header[mid, 2]).astype(np.float32)
    acts = rng.uniform(lo, hi, (2, N, 28)).astype(np.float32)
    for reward_mode in (1, 0):
        res = []
        for split in (None, "1", "4", "7"):
            if split:
                os.environ["PBRL_EMU_SPLIT"] = split
            try:
                b = EmuBackend(model_grav)
                b.set_mimic(mdesc)
                s = b.new_state(N)
                s.motion_id[:], s.ref_t0[:] = mid, t0
                b.reset(s, reward_mode=reward_mode)
                for k in range(2):
                    b.step(s, acts[k], kp, kd, reward_mode=reward_mode)
            finally:
                os.environ.pop("PBRL_EMU_SPLIT", None)
            res.append(s)
        for other in res[1:]:
            for name in ("qpos", "qvel", "time", "warm", "ctrl", "qpos_lag", "qvel_lag", "obs", "reward", "done", "metrics", "status", "cost"):
                assert np.array_equal(getattr(res[0], name), getattr(other, name)), (reward_mode, name)


def test_reward_after_one_substep(model_grav):
    table, mdesc, keep = pc.mimic_setup(model_grav)
    b = EmuBackend(model_grav)
    b.set_mimic(mdesc)
    pc.check_reward_after_one_substep(b, model_grav, table, mdesc, N=128)
