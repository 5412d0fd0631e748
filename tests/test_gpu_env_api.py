"""The reference-facing Python facade (envs.py / pds_controllers.py) on the GPU: same flows as the reference's notebooks."""
import importlib
import sys

import numpy as np
import pytest

from helpers import ASSETS, GOLDEN, PKG, ROOT, Oracle, OracleEnv, mocap, relerr, start_trajectory, util_data

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    return torch, importlib.import_module(PKG + ".envs"), importlib.import_module(PKG + ".pds_controllers")


def test_eager_pd_then_step_outside_reproduces_golden_T1(mods, model_nograv):
    """testing_pds_env3.ipynb cell 21 through the facade: eager stable-PD (dt = env.dt), step_custom_target_and_outside."""
    torch, envs, pdc = mods
    kp, kd = util_data.generate_kp_kd_gains()
    envs.register_environment("humanoidMimic", envs.HumanoidTrain)
    env = envs.get_environment("humanoidMimic", reference_trajectory_qpos=None, reference_trajectory_qvel=None,
                               duration_trajectory=0.0, dict_duration={}, model_path=model_nograv, kp_gains=kp, kd_gains=kd,
                               num_envs=2)
    assert abs(env.dt - 0.016) < 1e-12 and env.sys.dt == 0.002 and env.n_frames == 8   # testing_pds_env3.ipynb cells 23-24
    assert env.action_size == 28 and env.observation_size == 69
    g = np.load(GOLDEN / "mass_traces.npz")["T1"]
    coef = torch.from_numpy(start_trajectory()).cuda()
    state = env.reset(0)
    idx = [0, 1, 2, 31, 32, 33]
    for k in range(8):
        ps = state.pipeline_state
        M = ps.qM[0].cpu().numpy()
        assert np.abs(M[np.ix_(idx, idx)] - g[k]).max() <= 2e-5, k
        t = torch.clamp(ps.time, 1, 3)
        tau = pdc.stable_pd_controller_custom_trajectory(coef, env.sys, state, ps.q, ps.qd, kp, kd, t, env.dt)
        state = env.step_custom_target_and_outside(state, tau)
    assert state.obs.shape == (2, 69) and torch.equal(state.obs[:, :35], state.pipeline_state.qpos)
    assert (state.reward == 0).all() and (state.done == 1).all()
    assert set(state.metrics) == {"step_index", "pose_error", "fall"}


def test_step_variants_and_callbacks(mods, model_nograv):
    torch, envs, pdc = mods
    kp, kd = util_data.generate_kp_kd_gains()
    N = 32
    mk = lambda: envs.HumanoidTrain(None, None, 0.0, {}, model_nograv, kp, kd, num_envs=N)
    rng = np.random.default_rng(0)
    lo, hi = model_nograv.jnt_range[1:, 0], model_nograv.jnt_range[1:, 1]
    tgt = torch.from_numpy(rng.uniform(lo, hi, (N, 28)).astype(np.float32)).cuda()
    ref = OracleEnv(Oracle(model_nograv, "f64", tolerance=1e-14), N)
    # (1) fused built-in law == same law evaluated outside + torque step == oracle
    for fn, mode in ((pdc.standard_pd_controller, 1), (pdc.feedback_pd_controller, 2), (pdc.stable_pd_controller_custom_trajectory, 3)):
        a, b = mk(), mk()
        a.set_pd_callback(fn)
        sa, sb = a.reset(0), b.reset(0)
        sa = a.step(sa, tgt)
        ps = sb.pipeline_state
        tau = fn(tgt, b.sys, sb, ps.q, ps.qd, kp, kd, ps.time, b.sys.dt)
        sb = b.step_custom_target_and_outside(sb, tau)
        assert torch.allclose(sa.obs, sb.obs, rtol=0, atol=1e-6)
        ref.reset()
        ref.step(tgt.cpu().numpy(), kp, kd, pd_mode=mode, pd_dt=0.002)
        assert relerr(sa.obs.cpu().numpy(), np.concatenate([ref.qpos, ref.qvel], 1)) <= 1e-5
    # (2) an arbitrary Python callable as the PD law (the reference passes plain functions)
    c = mk()
    c.set_pd_callback(lambda target, sys, state, q, qd, kp_, kd_, time, dt: -kp_ * (q[:, 7:] - target) - kd_ * qd[:, 6:])
    sc = c.step(c.reset(0), tgt)
    d = mk()
    d.set_pd_callback(pdc.standard_pd_controller)
    sd = d.step(d.reset(0), tgt)
    assert torch.allclose(sc.obs, sd.obs, rtol=0, atol=1e-5)
    # (3) selected joints: the others keep the previous ctrl (agent_mimic.py:122-125)
    e = mk()
    se = e.reset(0)
    coef = torch.from_numpy(start_trajectory()).cuda()
    se = e.step_selected_joints_custom_target_and_joints(se, None, coef, 7, 10, 13, 9, 17, 24, torch.full((N,), 2.0).cuda())
    ctrl = se.pipeline_state.ctrl.cpu().numpy()
    others = [i for i in range(28) if i not in (7, 10, 13, 9)]
    assert (ctrl[:, others] == 0).all() and (np.abs(ctrl[:, [7, 10, 13, 9]]) > 0).all()


def test_replay_env(mods, model_grav):
    torch, envs, _ = mods
    z = np.load(ASSETS / "mocap_punch.npz")
    F = z["data_pos"].shape[0]
    env = envs.HumanoidReplay(z["data_pos"], z["data_vel"], float(z["total_time"]), {}, model_grav, num_envs=F)
    state = env.reset(0)
    state.pipeline_state.qpos.copy_(torch.from_numpy(z["data_pos"].astype(np.float32)))   # testing_forward_env.ipynb cell 10
    state.pipeline_state.qvel.copy_(torch.from_numpy(z["data_vel"].astype(np.float32)))
    state = env.step(state, None, want_poses=True)
    assert torch.equal(state.obs[:, :35], state.pipeline_state.qpos) and (state.done == 1).all() and (state.pipeline_state.time == 0).all()
    ref = OracleEnv(Oracle(model_grav, "f64"), F)
    _, _, _, xpos, xquat = ref.replay(z["data_pos"].astype(np.float32), z["data_vel"].astype(np.float32), want_poses=True)
    assert relerr(state.info["xpos"].cpu().numpy(), xpos) <= 2e-6


def test_mimic_env_rollout(mods, model_grav):
    torch, envs, _ = mods
    kp, kd = util_data.generate_kp_kd_gains()
    table = mocap.MotionTable.load(ASSETS / "motions.npz")
    env = envs.HumanoidTrain(None, None, 0.0, {}, model_grav, kp, kd, num_envs=256, reward_mode="mimic", motion_table=table)
    state = env.reset(3)
    assert (state.pipeline_state.qpos[:, 2] > 0.05).all()
    for k in range(4):
        state.info["stats"].zero_()
        state = env.step(state, state.pipeline_state.qpos[:, 7:].clone())
    torch.cuda.synchronize()
    r = state.reward[:, 0]
    assert (r > 0).all() and (r <= 1.0 + 1e-6).all() and torch.isfinite(state.obs).all()
    assert state.info["stats"][5].item() == 256 and (state.metrics["step_index"] == 4).all()
    assert abs(state.info["stats"][0].item() - r.sum().item()) < 1e-2


def test_training_wrappers(mods, model_grav):
    """wrappers.wrap = brax's EpisodeWrapper + AutoResetWrapper."""
    torch, envs, _ = mods
    wrappers = importlib.import_module(PKG + ".wrappers")
    kp, kd = util_data.generate_kp_kd_gains()
    N = 64
    # reference mode: done = 1 every step (agent_mimic.py:72) -> every env restarts from qpos0 every step
    env = wrappers.wrap(envs.HumanoidTrain(None, None, 0.0, {}, model_grav, kp, kd, num_envs=N), episode_length=5, action_repeat=2)
    state = env.reset(0)
    assert (state.info["steps"] == 0).all() and (state.info["truncation"] == 0).all()
    first = state.info["first_obs"].clone()
    tgt = torch.zeros(N, 28, device="cuda")
    state = env.step(state, tgt)
    assert (state.done == 1).all() and torch.equal(state.obs, first) and (state.info["steps"] == 2).all()
    assert (state.pipeline_state.time == 0).all() and (state.reward == 0).all()
    state = env.step(state, tgt)
    assert (state.info["steps"] == 2).all()          # the count restarted with the episode
    # mimic mode: episodes end by termination or at episode_length; restarted envs sit on their reference pose
    table = mocap.MotionTable.load(ASSETS / "motions.npz")
    base = envs.HumanoidTrain(None, None, 0.0, {}, model_grav, kp, kd, num_envs=N, reward_mode="mimic", motion_table=table)
    env = wrappers.wrap(base, episode_length=6)
    state = env.reset(1)
    rng = np.random.default_rng(0)
    lo, hi = model_grav.jnt_range[1:, 0], model_grav.jnt_range[1:, 1]
    ended = torch.zeros(N, dtype=torch.bool, device="cuda")
    for k in range(6):
        t0_before = state.info["ref_t0"].clone()
        state = env.step(state, torch.from_numpy(rng.uniform(lo, hi, (N, 28)).astype(np.float32)).cuda())
        d = state.done != 0
        ended |= d
        assert (state.pipeline_state.time[d] == 0).all() and (state.pipeline_state.qvel_lag[d] == state.pipeline_state.qvel[d]).all()
        assert torch.equal(state.obs[d][:, :35], state.pipeline_state.qpos[d])
        assert (state.info["ref_t0"][~d] == t0_before[~d]).all()
        if d.any():
            assert (state.info["ref_t0"][d] != t0_before[d]).any()      # a fresh clip phase per restart
        assert torch.isfinite(state.obs).all() and (state.reward[:, 0] >= 0).all()
    assert ended.all()                                                   # episode_length = 6 ends the survivors
    assert (state.info["truncation"][state.info["steps"] >= 6] >= 0).all()


def test_long_wrapped_rollout_stays_finite(mods, model_grav):
    """2 048 envs x 150 env steps (1 200 substeps) of random actions through the training wrappers: every state stays finite,
    no factorisation fails, episodes end and restart, and the reward statistics stay in range."""
    torch, envs, _ = mods
    wrappers = importlib.import_module(PKG + ".wrappers")
    kp, kd = util_data.generate_kp_kd_gains()
    table = mocap.MotionTable.load(ASSETS / "motions.npz")
    N = 2048
    env = wrappers.wrap(envs.HumanoidTrain(None, None, 0.0, {}, model_grav, kp, kd, num_envs=N, reward_mode="mimic", motion_table=table),
                        episode_length=60)
    state = env.reset(5)
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    lo = torch.from_numpy(model_grav.jnt_range[1:, 0].astype(np.float32)).cuda()
    hi = torch.from_numpy(model_grav.jnt_range[1:, 1].astype(np.float32)).cuda()
    restarts = torch.zeros(N, device="cuda")
    flags = torch.zeros(N, dtype=torch.int32, device="cuda")
    rsum = 0.0
    for k in range(150):
        act = lo + (hi - lo) * torch.rand(N, 28, generator=g, device="cuda")
        state = env.step(state, act)
        restarts += state.done
        flags |= state.info["status"]
        rsum += float(state.reward[:, 0].mean())
    torch.cuda.synchronize()
    assert torch.isfinite(state.obs).all() and torch.isfinite(state.pipeline_state.qacc_warmstart).all()
    assert int((flags & 3).max()) == 0                      # never a non-finite state or a non-positive pivot
    assert (restarts >= 2).all()                            # episode_length 60 alone ends every env at least twice
    assert 0.0 < rsum / 150 < 1.0
    assert (state.info["steps"] <= 60).all()


def test_state_clone_is_independent(mods, model_grav):
    """State.clone() keeps a copy: stepping the original must not touch the clone's state, metrics, status or clip phase
    (and the clone's metrics stay bound to its own buffer)."""
    torch, envs, _ = mods
    kp, kd = util_data.generate_kp_kd_gains()
    table = mocap.MotionTable.load(ASSETS / "motions.npz")
    env = envs.HumanoidTrain(None, None, 0.0, {}, model_grav, kp, kd, num_envs=40, reward_mode="mimic", motion_table=table)
    state = env.reset(2)
    state = env.step(state, state.pipeline_state.qpos[:, 7:].clone())
    torch.cuda.synchronize()
    keep = state.clone()
    snap = {k: v.clone() for k, v in keep.info.items() if torch.is_tensor(v)}
    snap_q, snap_m = keep.pipeline_state.qpos.clone(), {k: v.clone() for k, v in keep.metrics.items()}
    for k in range(3):
        state = env.step(state, state.pipeline_state.qpos[:, 7:].clone())
    torch.cuda.synchronize()
    assert torch.equal(keep.pipeline_state.qpos, snap_q) and not torch.equal(state.pipeline_state.qpos, snap_q)
    for k, v in snap.items():
        assert torch.equal(keep.info[k], v), k
        assert keep.info[k].data_ptr() != state.info[k].data_ptr(), k
    for k, v in snap_m.items():
        assert torch.equal(keep.metrics[k], v), k
    assert (state.metrics["step_index"] == 4).all() and (keep.metrics["step_index"] == 1).all()
    # the clone steps on its own buffers: its metrics are views of its own _metrics
    keep = env.step(keep, keep.pipeline_state.qpos[:, 7:].clone())
    torch.cuda.synchronize()
    assert (keep.metrics["step_index"] == 2).all() and (state.metrics["step_index"] == 4).all()


def test_step_rejects_malformed_inputs(mods, model_grav):
    """Raw pointers go to the kernel: wrong shapes must raise before the launch instead of reading out of bounds."""
    torch, envs, pdc = mods
    kp, kd = util_data.generate_kp_kd_gains()
    env = envs.HumanoidTrain(None, None, 0.0, {}, model_grav, kp, kd, num_envs=8)
    state = env.reset(0)
    good = torch.zeros(8, 28, device="cuda")
    for bad in (torch.zeros(7, 28, device="cuda"), torch.zeros(8, 27, device="cuda"), torch.zeros(8, 28, 3, device="cuda"), torch.zeros(8, 28, 4, 1, device="cuda")):
        with pytest.raises(ValueError):
            env.step(state, bad)
    with pytest.raises(ValueError):
        env.step_custom_target(state, None, good, torch.zeros(5, device="cuda"))
    other = envs.HumanoidTrain(None, None, 0.0, {}, model_grav, kp, kd, num_envs=4).reset(0)
    with pytest.raises(ValueError):
        env.step(other, good)
    with pytest.raises(ValueError):
        pdc.stable_pd_controller_custom_trajectory(torch.zeros(3, 28, device="cuda"), env.sys, state, state.pipeline_state.q, state.pipeline_state.qd,
                                                   kp, kd, state.pipeline_state.time, 0.002)
    state = env.step(state, good)                       # the well-formed call still works: (28,) and (28, 4) broadcast
    state = env.step(state, torch.zeros(28, device="cuda"))
    state = env.step(state, torch.zeros(28, 4, device="cuda"))
    torch.cuda.synchronize()
    assert torch.isfinite(state.obs).all()


def test_ppo_loop_consumes_the_env(mods):
    """SURVEY.md section 8f N4 (consumer side): a PPO loop -- rollouts through wrappers.wrap(HumanoidTrain(mimic)), GAE, the
    clipped surrogate -- runs on the env's device tensors without a host round trip, stays finite and updates the policy."""
    torch, _, _ = mods
    sys.path.insert(0, str(ROOT / "examples"))
    ppo = importlib.import_module("ppo_mimic")
    torch.manual_seed(0)
    hist, net = ppo.train(num_envs=256, iters=3, horizon=8, log=lambda r: None)
    assert len(hist) == 3 and all(np.isfinite([h["mean_reward"], h["loss"], h["done_frac"]]).all() for h in hist)
    assert all(0.0 <= h["mean_reward"] <= 1.0 + 1e-6 for h in hist)          # the mimic reward is a convex mix of exp(-k e) terms
    assert float(net.pi[-1].bias.detach().abs().max()) > 0.0                              # the optimiser moved the policy off its initial zero offset
    assert all(torch.isfinite(p).all() for p in net.parameters())
