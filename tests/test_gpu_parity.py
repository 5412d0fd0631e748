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


def test_reward_after_one_substep(model_grav):
    table, mdesc, keep = pc.mimic_setup(model_grav)
    print("reward after one substep", pc.check_reward_after_one_substep(GpuBackend(model_grav, mimic_desc=mdesc), model_grav, table, mdesc))


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


@pytest.mark.parametrize("bulk", ["0", "1"])
def test_replay_rollout(gpu_nograv, model_nograv, bulk, monkeypatch):
    # both store paths of the frame-table form (ordinary stores / bulk-copy engine; the library reads the switch per call)
    monkeypatch.setenv("PBRL_REPLAY_BULK", bulk)
    pc.check_replay_rollout(gpu_nograv, model_nograv)
    pc.check_replay_rollout(gpu_nograv, model_nograv, F=227, N=4096, K=3, seed=38)   # several CTAs per step, ragged tail
    pc.check_replay_rollout(gpu_nograv, model_nograv, F=5, N=1, K=1, seed=39)
    pc.check_replay_rollout(gpu_nograv, model_nograv, F=600, N=300, K=5, seed=41)    # fewer than 4 items per clip row: kinematics per item
    pc.check_replay_rollout(gpu_nograv, model_nograv, F=5, N=2, K=41, seed=43)      # frame table with fewer envs than items per warp pass
    pc.check_replay_rollout(gpu_nograv, model_nograv, F=4, N=3, K=7, seed=44)
    pc.check_replay_rollout(gpu_nograv, model_nograv, F=7, N=1, K=64, seed=45)
    pc.check_replay_rollout(gpu_nograv, model_nograv, F=227, N=4099, K=37, seed=46)  # several passes per warp, ragged last block
    pc.check_replay_rollout_paths_agree(gpu_nograv, model_nograv)
    pc.check_replay_rollout_paths_agree(gpu_nograv, model_nograv, F=227, N=1000, K=3, seed=42)


def test_contact_rich_statistics(model_grav):
    table, mdesc, keep = pc.mimic_setup(model_grav)
    pc.check_contact_rich_statistics(GpuBackend(model_grav, mimic_desc=mdesc), model_grav, table, mdesc)


def test_hovering_foot_boxes(gpu_grav, model_grav):
    pc.check_hovering_foot_boxes(gpu_grav, model_grav)


def test_sphere_box_contacts(gpu_grav, model_grav):
    pc.check_sphere_box_contacts(gpu_grav, model_grav, want=30)


def test_config3_punch_tracking(model_nograv):
    """BASELINE configs[2] as named: final_humanoid_no_gravity.xml, punch clip reference-state init, PD targets that track the
    clip (bench.py --workload contact_free --actions reference), 32 substeps <= 1e-3 per env vs the float64 oracle."""
    table, mdesc, keep, ids = pc.named_config_setup(model_nograv, ["punch"])
    out = pc.check_config3_punch_tracking(GpuBackend(model_nograv, mimic_desc=mdesc), model_nograv, table, mdesc, ids, N=256)
    print("config3", out)


def test_config4_backflip_jump(model_grav):
    """BASELINE configs[3] as named: backflip (even env ids) + jump (odd): first env step per env, then 24-step statistics."""
    table, mdesc, keep, ids = pc.named_config_setup(model_grav, ["backflip", "jump"])
    out = pc.check_config4_backflip_jump(GpuBackend(model_grav, mimic_desc=mdesc), model_grav, table, mdesc, ids, N=256, N_stat=1024, steps=24)
    print("config4", out)


def test_alt_rate_1200hz_40_frames(model_nograv, model_grav):
    """The reference's second rate configuration (agent_env_template2.py:49-58): timestep 1/1200, n_frames 40."""
    pc.check_alt_rate(lambda m: GpuBackend(m), model_nograv, N=64)
    pc.check_alt_rate(lambda m: GpuBackend(m), model_grav, N=64)


def test_balance_order_is_a_permutation(gpu_grav):
    """pbrl_balance_order: every env id exactly once (a duplicate would silently skip an env in k_step), sorted by
    decreasing clamped cost, for costs beyond the 256 bins and negative ones, N in {1, 17, 65536}."""
    import numpy as np
    import torch
    be = gpu_grav
    L = be.L
    rng = np.random.default_rng(61)
    for N in (1, 17, 2368, 65536):
        for kind in ("small", "wide", "const"):
            cost = {"small": rng.integers(0, 12, N), "wide": rng.integers(-5, 2000, N), "const": np.full(N, 3)}[kind].astype(np.int32)
            c = torch.from_numpy(cost).to(be.dev)
            order = torch.full((N,), -1, dtype=torch.int32, device=be.dev)
            L.check(L.lib().pbrl_balance_order(be.model.handle, N, L._ptr(c), L._ptr(order), be._stream()))
            torch.cuda.synchronize()
            o = order.cpu().numpy()
            assert np.array_equal(np.sort(o), np.arange(N)), (N, kind)
            key = np.clip(cost, 0, 255)[o]
            assert (np.diff(key) <= 0).all(), (N, kind)


@pytest.mark.parametrize("pinned", [False, True])
def test_host_buffer_layer_is_bit_identical_to_device_buffers(model_grav, pinned):
    """pbrl_batch_reset_host / pbrl_batch_step_host / pbrl_batch_get_state_host (the API bench.py's e2e number flows through)
    against pbrl_reset / pbrl_step (+ pbrl_balance_order) on device buffers: identical obs / reward / done / stats and
    state, bit for bit, over several env steps in mimic mode and in reference mode.  `pinned`: the caller's buffers are
    page-locked, device-mapped host memory -- k_step then reads the action rows and writes obs / reward / done in place (no
    staging copies); pageable buffers take the copies."""
    import ctypes
    import numpy as np
    import torch
    host = (lambda *s: torch.zeros(s, dtype=torch.float32).pin_memory().numpy()) if pinned else (lambda *s: np.zeros(s, np.float32))
    table, mdesc, keep = pc.mimic_setup(model_grav)
    be = GpuBackend(model_grav, mimic_desc=mdesc)
    L, lib = be.L, be.L.lib()
    N = 2500                                     # more than one wave of 148 x 16 slots, ragged tail
    rng = np.random.default_rng(67)
    kp, kd = pc.util_data.generate_kp_kd_gains()
    lo, hi = model_grav.jnt_range[1:, 0], model_grav.jnt_range[1:, 1]
    mid = (np.arange(N) % 7).astype(np.int32)
    t0 = (rng.uniform(0, 1, N) * table.header[mid, 2]).astype(np.float32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for reward_mode in (1, 0):
        batch = ctypes.c_void_p()
        L.check(lib.pbrl_batch_create(be.model.handle, N, ctypes.byref(batch)))
        h_obs, h_rew, h_done, h_stats, h_act = host(N, 69), host(N, 3), host(N), np.zeros(8, np.float32), host(N, 28)
        L.check(lib.pbrl_batch_reset_host(batch, reward_mode, vp(mid) if reward_mode else None, vp(t0) if reward_mode else None, vp(h_obs)))
        # the same on device buffers
        z = lambda *s: torch.zeros(s, device=be.dev)
        d = dict(qpos=z(N, 35), qvel=z(N, 34), time=z(N), warm=z(N, 34), ctrl=z(N, 28), qpos_lag=z(N, 35), qvel_lag=z(N, 34), obs=z(N, 69),
                 reward=z(N, 3), done=z(N), metrics=z(N, 3), status=torch.zeros(N, dtype=torch.int32, device=be.dev), stats=z(8),
                 cost=torch.zeros(N, dtype=torch.int32, device=be.dev), env_order=torch.arange(N, dtype=torch.int32, device=be.dev))
        if reward_mode:
            d["motion_id"], d["ref_t0"] = torch.from_numpy(mid).to(be.dev), torch.from_numpy(t0).to(be.dev)
        b = L.make_buffers(**d)
        L.check(lib.pbrl_reset(be.model.handle, N, reward_mode, ctypes.byref(b), be._stream()))
        torch.cuda.synchronize()
        assert np.array_equal(h_obs, d["obs"].cpu().numpy())
        kpd, kdd = be._dp(kp), be._dp(kd)
        for k in range(3):
            act = rng.uniform(lo, hi, (N, 28)).astype(np.float32)
            h_act[:] = act
            h_obs[:], h_rew[:], h_done[:] = -7, -7, -7
            L.check(lib.pbrl_batch_step_host(batch, 8, 3, 1, vp(h_act), vp(kp), vp(kd), ctypes.c_float(0.002), reward_mode, vp(h_obs), vp(h_rew),
                                             vp(h_done), vp(h_stats)))
            a = torch.from_numpy(act).to(be.dev)
            d["stats"].zero_()
            L.check(lib.pbrl_step(be.model.handle, N, 8, 3, 1, L._ptr(a), L._ptr(kpd), L._ptr(kdd), ctypes.c_float(0.002), None,
                                  ctypes.c_uint(0x0FFFFFFF), reward_mode, ctypes.byref(b), be._stream()))
            L.check(lib.pbrl_balance_order(be.model.handle, N, L._ptr(d["cost"]), L._ptr(d["env_order"]), be._stream()))
            torch.cuda.synchronize()
            assert np.array_equal(h_obs, d["obs"].cpu().numpy()), (reward_mode, k)
            assert np.array_equal(h_rew, d["reward"].cpu().numpy()) and np.array_equal(h_done, d["done"].cpu().numpy()), (reward_mode, k)
            if reward_mode:   # the statistics are float atomics over warps: same addends, order may differ
                assert np.allclose(h_stats[:6], d["stats"].cpu().numpy()[:6], rtol=1e-5, atol=1e-3) and h_stats[5] == N
        q, v, t = np.zeros((N, 35), np.float32), np.zeros((N, 34), np.float32), np.zeros(N, np.float32)
        L.check(lib.pbrl_batch_get_state_host(batch, vp(q), vp(v), vp(t)))
        assert np.array_equal(q, d["qpos"].cpu().numpy()) and np.array_equal(v, d["qvel"].cpu().numpy()) and np.array_equal(t, d["time"].cpu().numpy())
        assert np.array_equal(h_obs, np.concatenate([q, v], 1))
        lib.pbrl_batch_destroy(batch)


def test_solver_rule_mjx_and_options(model_grav):
    """solver_rule = 1: MuJoCo / MJX's tolerance rule with the model's iterations / tolerance / ls_* options, vs the oracle under
    the same options; and the distance between the two rules' answers on contact-rich states."""
    out = pc.check_solver_rule(lambda m: GpuBackend(m), model_grav, N=128, seed=72)
    print("solver_rule", out)


def test_capsule_box_and_box_box_contacts(gpu_grav, model_grav):
    """The 15 candidates against the foot boxes (capsule-box x 14, box-box x 1) on the CUDA path: census 89 / 89."""
    out = pc.check_box_pair_contacts(gpu_grav, model_grav, want=300, seeds=range(400, 480))
    print("box pairs", out)


def test_quiet_stance_carries_the_body_weight(gpu_grav, model_grav):
    out = pc.check_quiet_stance(gpu_grav, model_grav)
    print("quiet stance", out)


def test_small_batches_split_over_the_sms_bit_identically(model_grav):
    """k_step hands a batch with fewer groups of 16 envs than SMs out in parts of 8 / 4 / 2 / 1 envs per CTA (the other warps of
    the CTA mirror the barrier sequence).  Which warp of which CTA steps an env must not change a bit of its result: the same
    rollout with the split on (default) and off (PBRL_TAIL_SPLIT=0 at model creation), for batch sizes on either side of every
    split factor and a ragged last group."""
    import os
    import numpy as np
    kp, kd = pc.util_data.generate_kp_kd_gains()
    lo, hi = model_grav.jnt_range[1:, 0], model_grav.jnt_range[1:, 1]
    os.environ["PBRL_TAIL_SPLIT"] = "0"
    try:
        be_off = GpuBackend(model_grav)
    finally:
        os.environ.pop("PBRL_TAIL_SPLIT")
    be_on = GpuBackend(model_grav)
    for N in (1, 9, 17, 100, 300, 600, 1190, 2368 + 5):
        rng = np.random.default_rng(N)
        qpos, qvel = pc.random_states(model_grav, N, N, gravity_heights=True)
        acts = rng.uniform(lo, hi, (2, N, 28)).astype(np.float32)
        res = []
        for be in (be_on, be_off):
            s = be.new_state(N)
            s.set_state(qpos.astype(np.float32), qvel.astype(np.float32))
            for k in range(2):
                be.step(s, acts[k], kp, kd)
            res.append(s)
        for name in ("qpos", "qvel", "warm", "ctrl", "obs", "time", "qpos_lag", "qvel_lag"):
            assert np.array_equal(getattr(res[0], name), getattr(res[1], name)), (N, name)
        assert np.isfinite(res[0].qpos).all() and (res[0].status & 7 == 0).all()


def test_step_in_two_parts_with_regrouping_is_bit_identical(model_grav):
    """Large batches that pass env_order / cost / status are stepped in two launches with the envs re-sorted in between
    (include/pbrl.h: pbrl_step_launches).  State, observations, rewards, status and the work estimate must be bit-identical to
    the single launch (PBRL_REGROUP_AFTER=0 at model creation), in mimic mode and in reference mode, over several steps."""
    import ctypes
    import os
    import numpy as np
    import torch
    table, mdesc, keep = pc.mimic_setup(model_grav)
    os.environ["PBRL_REGROUP_AFTER"] = "0"
    try:
        be1 = GpuBackend(model_grav, mimic_desc=mdesc)
    finally:
        os.environ.pop("PBRL_REGROUP_AFTER")
    os.environ["PBRL_REGROUP_FUSED"] = "0"   # two launches with the sort in between (the default for batches of >= 16 groups per SM)
    try:
        be2 = GpuBackend(model_grav, mimic_desc=mdesc)
    finally:
        os.environ.pop("PBRL_REGROUP_FUSED")
    os.environ["PBRL_REGROUP_FUSED"] = "1"   # the two parts inside one launch (work-class queues between them; the default below that)
    try:
        be3 = GpuBackend(model_grav, mimic_desc=mdesc)
    finally:
        os.environ.pop("PBRL_REGROUP_FUSED")
    os.environ["PBRL_REGROUP_FUSED"] = "1"
    os.environ["PBRL_REGROUP_PARTS"] = "4"   # ... in four parts (three regroupings)
    try:
        be4 = GpuBackend(model_grav, mimic_desc=mdesc)
    finally:
        os.environ.pop("PBRL_REGROUP_FUSED")
        os.environ.pop("PBRL_REGROUP_PARTS")
    N = 5000
    rng = np.random.default_rng(91)
    kp, kd = pc.util_data.generate_kp_kd_gains()
    lo, hi = model_grav.jnt_range[1:, 0], model_grav.jnt_range[1:, 1]
    mid = (np.arange(N) % 7).astype(np.int32)
    t0 = (rng.uniform(0, 1, N) * table.header[mid, 2]).astype(np.float32)
    acts = rng.uniform(lo, hi, (3, N, 28)).astype(np.float32)
    for reward_mode in (1, 0):
        outs = []
        for be in (be1, be2, be3, be4):
            L, lib = be.L, be.L.lib()
            z = lambda *s: torch.zeros(s, device=be.dev)
            d = dict(qpos=z(N, 35), qvel=z(N, 34), time=z(N), warm=z(N, 34), ctrl=z(N, 28), qpos_lag=z(N, 35), qvel_lag=z(N, 34), obs=z(N, 69),
                     reward=z(N, 3), done=z(N), metrics=z(N, 3), status=torch.zeros(N, dtype=torch.int32, device=be.dev), stats=z(8),
                     cost=torch.zeros(N, dtype=torch.int32, device=be.dev), env_order=torch.arange(N, dtype=torch.int32, device=be.dev),
                     motion_id=torch.from_numpy(mid).to(be.dev), ref_t0=torch.from_numpy(t0).to(be.dev))
            b = L.make_buffers(**d)
            assert lib.pbrl_step_launches(be.model.handle, N, 8, ctypes.byref(b)) == (3 if be is be2 else 1)
            assert lib.pbrl_step_launches(be.model.handle, 64, 8, ctypes.byref(b)) == 1
            L.check(lib.pbrl_reset(be.model.handle, N, reward_mode, ctypes.byref(b), be._stream()))
            kpd, kdd = be._dp(kp), be._dp(kd)
            for k in range(3):
                a = torch.from_numpy(acts[k]).to(be.dev)
                L.check(lib.pbrl_step(be.model.handle, N, 8, 3, 1, L._ptr(a), L._ptr(kpd), L._ptr(kdd), ctypes.c_float(0.002), None,
                                      ctypes.c_uint(0x0FFFFFFF), reward_mode, ctypes.byref(b), be._stream()))
                L.check(lib.pbrl_balance_order(be.model.handle, N, L._ptr(d["cost"]), L._ptr(d["env_order"]), be._stream()))
            torch.cuda.synchronize()
            outs.append({k: v.cpu().numpy() for k, v in d.items()})
        for other in outs[1:]:
            for name in ("qpos", "qvel", "time", "warm", "ctrl", "qpos_lag", "qvel_lag", "obs", "reward", "done", "metrics", "status", "cost"):
                assert np.array_equal(outs[0][name], other[name]), (reward_mode, name)
            assert np.allclose(outs[0]["stats"], other["stats"], rtol=1e-5, atol=1e-3)
        for o in outs:   # (the counting sort places equal costs in arrival order: the orders are permutations sorted alike, not equal)
            assert np.array_equal(np.sort(o["env_order"]), np.arange(N)) and (np.diff(np.clip(o["cost"], 0, 255)[o["env_order"]]) <= 0).all()
