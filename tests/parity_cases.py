"""Parity cases shared by the host-emulation tests (no GPU) and the `-m gpu` tests (libpbrl on a B200).

Each case drives a back end (helpers.EmuBackend / helpers.GpuBackend) and the oracle on identical seeded
inputs and compares at the tolerances BASELINE.json's north_star states:
  fp32 relative error <= 1e-5 on qpos/qvel after one substep, <= 1e-3 after 32 contact-free substeps;
  contact-rich rollouts: matching termination / contact statistics.
The oracle runs in float64 with a tight solver tolerance; the back end computes in fp32.
"""
import numpy as np

from helpers import ASSETS, DBG, GOLDEN, Oracle, OracleEnv, mimic, mocap, random_states, relerr, row_relerr, start_trajectory, util_data

TOL_1 = 1e-5    # after one substep (north_star)
TOL_32 = 1e-3   # after 32 contact-free substeps (north_star)


def _oracle(model):
    return Oracle(model, "f64", tolerance=1e-14)


def check_forward_stages(backend, model, gravity, N=24, seed=3):
    """Every stage of one mjx.forward against the oracle's intermediates."""
    orc = _oracle(model)
    rng = np.random.default_rng(seed)
    qpos, qvel = random_states(model, N, seed, gravity_heights=gravity, vel_scale=1.0, joint_frac=0.6)
    if not gravity:  # drive a few joints past their limits so that limit rows are active
        qpos[:, 7 + 9] = rng.uniform(-0.3, 0.0, N)   # right elbow below its lower limit 0
        qpos[:, 7 + 17] = rng.uniform(0.0, 0.3, N)   # right knee above its upper limit 0
    ctrl = rng.uniform(-1, 1, (N, 28))
    warm = rng.normal(size=(N, 34))
    out = backend.debug_forward(qpos, qvel, ctrl, warm)
    q32, v32, c32, w32 = [np.asarray(x, np.float32) for x in (qpos, qvel, ctrl, warm)]
    worst = {}
    active_total = self_total = 0
    for e in range(N):
        d = orc.new_data()
        d.qpos[:35], d.qvel[:34], d.ctrl[:28], d.qacc_warmstart[:34] = q32[e], v32[e], c32[e], w32[e]
        orc.forward(d)
        get = lambda k: out[e, DBG[k][0]:DBG[k][0] + DBG[k][1]]
        ref = dict(xpos=d.xpos[:42], xquat=d.xquat[:56], com=d.com[:3], qM=d.qM_dense().ravel(), bias=d.qfrc_bias[:34],
                   qfrc_smooth=d.qfrc_smooth[:34], qacc_smooth=d.qacc_smooth[:34], qacc=d.qacc[:34],
                   qfrc_constraint=d.qfrc_constraint[:34])
        nc, nself = int(d.ncon[0]), int(d.nself[0])
        for k, r in ref.items():
            # near-parallel capsules (thigh against thigh) make the closest-point problem ill-conditioned: the fp32 contact
            # normal moves in the 5th digit, and so do the accelerations (the fp32 oracle agrees with the kernel to 3e-7)
            kk = k + "_selfcontact" if nself and k in ("qacc", "qfrc_constraint") else k
            worst[kk] = max(worst.get(kk, 0.0), relerr(get(k), r))
        nfloor = nc - nself
        n_act = int((d.con_dist[:nfloor] - d.con_includemargin[:nfloor] < 0).sum())
        assert int(get("ncon_active")[0]) == n_act, (e, get("ncon_active"), n_act)
        # body-vs-body contacts inside the margin: same set (the two sides list them in different pair orders)
        assert int(get("nself")[0]) == min(nself, 8), (e, get("nself"), nself)
        if 0 < nself <= 8:
            assert np.abs(np.sort(get("self_dist")[:nself]) - np.sort(d.con_dist[nfloor:nc])).max() <= 2e-6
            f_ref = np.sort(d.efc_force[28 + 4 * nfloor:28 + 4 * nc].reshape(nself, 4).sum(1))
            f_got = np.sort(get("self_force")[:4 * nself].reshape(nself, 4).sum(1))
            assert np.abs(f_got - f_ref).max() <= 1e-3 * max(1.0, np.abs(f_ref).max()), (e, f_got, f_ref)
        self_total += nself
        active_total += n_act + int((d.efc_force[:28] > 0).sum())
    tol = dict(xpos=2e-6, xquat=2e-6, com=2e-6, qM=2e-6, bias=2e-5, qfrc_smooth=2e-5, qacc_smooth=2e-5, qacc=5e-5,
               qfrc_constraint=1e-4, qacc_selfcontact=2e-4, qfrc_constraint_selfcontact=2e-4)
    for k, v in worst.items():
        assert v <= tol[k], (k, v, worst)
    assert active_total > 0, "the case must exercise active constraint rows"
    worst["self_contacts"] = self_total
    return worst


def _paired_envs(backend, model, N, seed, gravity):
    orc = _oracle(model)
    env = OracleEnv(orc, N)
    qpos, qvel = random_states(model, N, seed, gravity_heights=gravity)
    q32, v32 = qpos.astype(np.float32), qvel.astype(np.float32)
    env.set_state(q32, v32)
    s = backend.new_state(N)
    s.set_state(q32, v32)
    return env, s


def _self_penetration(model, qpos):
    """Most negative body-vs-body contact distance of each state (0 if nothing overlaps), from the oracle's collision pass."""
    orc = _oracle(model)
    out = np.zeros(len(qpos))
    for e, q in enumerate(qpos):
        d = orc.new_data()
        d.qpos[:35] = q
        orc.forward(d)
        nc, ns = int(d.ncon[0]), int(d.nself[0])
        if ns:
            out[e] = min(0.0, d.con_dist[nc - ns:nc].min())
    return out


def assert_qacc_errors_by_depth(errs, pens, what=""):
    """Bound on the relative error of qacc for states with body-vs-body contacts, stated by the deepest overlap `pens` (metres).
    The Hessian M + D u u^T of such a row is the worse conditioned the deeper the overlap; kernel, host emulation and the
    oracle's own fp32 build share one error distribution per bin (tools/dev/dense_accuracy.py: p95 3e-5 below 10 mm, 6e-5 below
    30 mm, 1e-4 below 60 mm, 5e-4 beyond, isolated outliers up to 1e-1 beyond 30 mm on any of the three).  <= 3e-4 up to 10 mm
    (all a rollout reaches), <= 2e-3 up to 30 mm; deeper (unphysical) overlaps through their 90th percentile, <= 1e-3."""
    errs, pens = np.asarray(errs), np.asarray(pens)
    bound = np.where(pens < 0.010, 3e-4, np.where(pens < 0.030, 2e-3, np.inf))
    assert (errs <= bound).all(), (what, float((errs / bound).max()))
    deep = pens >= 0.030
    assert deep.sum() == 0 or np.quantile(errs[deep], 0.9) <= 1e-3, (what, float(np.quantile(errs[deep], 0.9)))


def check_one_substep(backend, model, gravity, N=32, seed=5):
    """fp32 relative error <= 1e-5 on qpos/qvel after ONE substep (torque-driven, random ctrl)."""
    env, s = _paired_envs(backend, model, N, seed, gravity)
    q0, s0v = env.qpos.copy(), env.qvel.copy()
    rng = np.random.default_rng(seed)
    tau = rng.uniform(-0.6, 0.6, (N, 28)).astype(np.float32)
    kp, kd = util_data.generate_kp_kd_gains()
    env.step(tau, kp, kd, pd_mode=0, n_frames=1)
    backend.step(s, tau, kp, kd, pd_mode=0, n_frames=1)
    eq, ev = row_relerr(s.qpos, env.qpos), row_relerr(s.qvel, env.qvel)
    # Random poses interpenetrate by centimetres (limbs inside the torso): forces of 1e3 N on an ill-conditioned
    # closest-point problem.  The north-star tolerance is asserted on the envs whose bodies overlap by at most 2 mm
    # (every state a rollout can reach); the bound on the others grows with the overlap (512 random states per model, host
    # emulation and the oracle's own fp32 build alike: max 8e-6 up to 10 mm, 3e-5 up to 30 mm, 5e-5 up to 60 mm, 2e-4 beyond).
    pen = -_self_penetration(model, q0)
    deep = pen > 2e-3
    assert (~deep).sum() >= N // 3
    assert eq[~deep].max() <= TOL_1 and ev[~deep].max() <= TOL_1, (eq[~deep].max(), ev[~deep].max())
    bound = np.where(pen <= 2e-3, TOL_1, np.where(pen <= 0.010, 2e-5, np.where(pen <= 0.030, 5e-5, np.where(pen <= 0.060, 1e-4, 3e-4))))
    assert eq.max() <= TOL_1 and (ev <= bound).all(), (eq.max(), (ev / bound).max())
    # (the envelope is the conditioning of the problem in fp32, not the kernel's arithmetic)
    env32 = OracleEnv(Oracle(model, "f32"), N)
    env32.set_state(q0.astype(np.float32), np.asarray(s0v, np.float32))
    env32.step(tau, kp, kd, pd_mode=0, n_frames=1)
    assert row_relerr(env32.qvel, env.qvel)[~deep].max() <= TOL_1   # the fp32 oracle meets the tolerance on the same set
    assert np.allclose(s.time, env.time, atol=1e-7)
    assert relerr(s.qpos_lag, env.qpos_lag) <= 1e-6 and relerr(s.warm, env.warm) <= 1e-4
    assert (s.status == 0).all()
    return float(eq.max()), float(ev.max())


def check_32_substeps(backend, model, pd_mode=3, N=32, seed=7):
    """<= 1e-3 after 32 contact-free substeps (4 env steps x 8), PD targets inside the joint ranges.  Random poses put limbs
    against each other in a quarter of the envs (body-vs-body rows; the oracle says which): the north-star bound is asserted
    on every env whose rollout stays contact-free, the others -- contact rows switching on and off amplify fp32 rounding --
    are held to 5e-3."""
    env, s = _paired_envs(backend, model, N, seed, gravity=False)
    rng = np.random.default_rng(seed)
    kp, kd = util_data.generate_kp_kd_gains()
    lo, hi = model.jnt_range[1:, 0], model.jnt_range[1:, 1]
    contact = np.array([_any_contact_row(model, q) for q in env.qpos])
    for k in range(4):
        action = (rng.uniform(-0.6, 0.6, (N, 28)) if pd_mode == 0 else rng.uniform(lo, hi, (N, 28))).astype(np.float32)
        env.step(action, kp, kd, pd_mode=pd_mode, pd_dt=0.002, n_frames=8)
        backend.step(s, action, kp, kd, pd_mode=pd_mode, pd_dt=0.002, n_frames=8)
        contact |= np.array([_any_contact_row(model, q) for q in env.qpos])
    eq, ev = row_relerr(s.qpos, env.qpos), row_relerr(s.qvel, env.qvel)
    free = ~contact
    assert free.mean() >= 0.5, free.mean()
    assert eq[free].max() <= TOL_32 and ev[free].max() <= TOL_32, (eq[free].max(), ev[free].max())
    assert eq.max() <= 5 * TOL_32 and ev.max() <= 5 * TOL_32, (eq.max(), ev.max())
    assert relerr(s.obs[free], np.concatenate([env.qpos, env.qvel], 1)[free]) <= TOL_32
    assert (s.reward == 0).all() and (s.done == 1).all()  # agent_mimic.py:72,84
    return float(eq[free].max()), float(ev[free].max())


def check_pd_torque(backend, model, N=16, seed=11):
    """The three PD laws + dense qM / qfrc_bias against the oracle (pds_controllers_agents.py)."""
    env, s = _paired_envs(backend, model, N, seed, gravity=False)
    rng = np.random.default_rng(seed)
    kp, kd = util_data.generate_kp_kd_gains()
    lo, hi = model.jnt_range[1:, 0], model.jnt_range[1:, 1]
    tgt = rng.uniform(lo, hi, (N, 28)).astype(np.float32)
    env.step(tgt, kp, kd, pd_mode=1, n_frames=3)          # make the lag state differ from the state
    backend.step(s, tgt, kp, kd, pd_mode=1, n_frames=3)
    s.qpos[:], s.qvel[:], s.qpos_lag[:], s.qvel_lag[:] = env.qpos, env.qvel, env.qpos_lag, env.qvel_lag
    out = {}
    for mode, dt in ((1, 0.002), (2, 0.016), (3, 0.016), (3, 0.002)):
        t_ref, qM_ref, b_ref = env.pd_torque(tgt, kp, kd, mode, dt, want_qM=True)
        t, qM, b = backend.pd_torque(s, tgt, kp, kd, mode, dt, want_qM=True)
        out[(mode, dt)] = relerr(t, t_ref)
        assert relerr(t, t_ref) <= 2e-5, (mode, dt, relerr(t, t_ref))
        assert relerr(qM, qM_ref) <= 2e-6 and relerr(b, b_ref) <= 2e-5
    coef = np.tile(start_trajectory(), (N, 1, 1)).astype(np.float32)
    tt = np.full(N, 1.7, np.float32)
    t_ref = env.pd_torque(coef, kp, kd, 3, 0.016, time=tt)
    t = backend.pd_torque(s, coef, kp, kd, 3, 0.016, time=tt)
    assert relerr(t, t_ref) <= 2e-5
    return out


def check_golden_T1(backend, model_nograv, nprint=11):
    """The reference's own known answers through the back end: testing_pds_env3.ipynb cell 21
    (eager stable-PD with dt = env.dt = 0.016, then step_custom_target_and_outside)."""
    g = np.load(GOLDEN / "mass_traces.npz")["T1"]
    kp, kd = util_data.generate_kp_kd_gains()
    coef = start_trajectory()[None].astype(np.float32)
    s = backend.new_state(1)
    backend.reset(s)
    idx = [0, 1, 2, 31, 32, 33]
    errs = []
    for k in range(nprint):
        t = np.clip(s.time, 1, 3).astype(np.float32)
        tau, qM, _ = backend.pd_torque(s, coef, kp, kd, 3, 0.016, time=t, want_qM=True)
        errs.append(float(np.abs(qM[0][np.ix_(idx, idx)] - g[k]).max()))
        backend.step(s, tau, kp, kd, pd_mode=0, n_frames=8)
    assert max(errs[:4]) <= 5e-7, errs       # T0 and the first prints: print precision of fp32
    assert max(errs) <= 2e-5, errs           # prints 4..10 (tight window)
    return errs


def check_golden_T2_T4(backend, model_nograv, nprint=6):
    """Jitted stable-PD inside the step (pd_dt = sys.dt) and the selected-joint variants
    (testing_pds_env.ipynb cells 24/30/33; agent_mimic.py:94-190,200-236)."""
    g = np.load(GOLDEN / "mass_traces.npz")
    kp, kd = util_data.generate_kp_kd_gains()
    coef = start_trajectory()[None].astype(np.float32)
    idx = [0, 1, 2, 31, 32, 33]
    out = {}
    for name, joints in (("T2", None), ("T3", [7, 10, 13, 9]), ("T4", [7, 10, 13, 9, 17, 24])):
        mask = 0x0FFFFFFF if joints is None else sum(1 << j for j in joints)
        s = backend.new_state(1)
        backend.reset(s)
        errs = []
        for k in range(nprint):
            t = np.clip(s.time, 1, 3).astype(np.float32)
            _, qM, _ = backend.pd_torque(s, coef, kp, kd, 3, 0.002, time=t, want_qM=True)
            errs.append(float(np.abs(qM[0][np.ix_(idx, idx)] - g[name][k]).max()))
            backend.step(s, coef, kp, kd, pd_mode=3, pd_dt=0.002, n_frames=8, pd_time=t, ctrl_mask=mask)
        assert max(errs[:3]) <= 1e-6 and max(errs) <= 5e-5, (name, errs)
        out[name] = errs
    return out


def check_replay(backend, model, N=64, seed=13):
    """HumanoidReplay.step: obs = [qpos, qvel], reward zeros(3), done 1; body poses vs the oracle FK."""
    orc = _oracle(model)
    env = OracleEnv(orc, N)
    qpos, qvel = random_states(model, N, seed, vel_scale=1.0, joint_frac=0.9)
    qpos[0, 3:7] *= 0.995  # un-normalised root quaternion: FK normalises, obs echoes
    q32, v32 = qpos.astype(np.float32), qvel.astype(np.float32)
    obs_r, rew_r, done_r, xpos_r, xquat_r = env.replay(q32, v32, want_poses=True)
    obs, rew, done, xpos, xquat = backend.replay(q32, v32, want_poses=True)
    assert np.array_equal(obs, np.concatenate([q32, v32], 1))  # bit-exact echo
    assert (rew == 0).all() and (done == 1).all()
    assert relerr(xpos, xpos_r) <= 2e-6 and relerr(xquat, xquat_r) <= 2e-6
    return relerr(xpos, xpos_r)


def check_reset(backend, model):
    s = backend.new_state(5)
    s.qpos[:] = 7.0
    s.warm[:] = 3.0
    backend.reset(s)
    assert np.array_equal(s.qpos, np.tile(model.qpos0.astype(np.float32), (5, 1)))
    assert np.array_equal(s.qpos_lag, s.qpos) and (s.qvel == 0).all() and (s.qvel_lag == 0).all()
    assert (s.warm == 0).all() and (s.ctrl == 0).all() and (s.time == 0).all()
    assert np.array_equal(s.obs, np.concatenate([s.qpos, s.qvel], 1))
    assert (s.reward == 0).all() and (s.done == 0).all()


def check_contact_rollout(backend, model_grav, N=32, steps=40, seed=17):
    """Contact-rich drop test: bodies fall from reset-like poses onto the floor.  Trajectories are compared
    statistically (the contact problem amplifies fp32 noise): touchdown step, final root height, and the
    early (pre-chaos) window at the 32-substep tolerance."""
    env, s = _paired_envs(backend, model_grav, N, seed, gravity=True)
    kp, kd = util_data.generate_kp_kd_gains()
    tau = np.zeros((N, 28), np.float32)
    first_ref, first = np.full(N, -1), np.full(N, -1)
    early = None
    for k in range(steps):
        env.step(tau, kp, kd, pd_mode=0, n_frames=8)
        backend.step(s, tau, kp, kd, pd_mode=0, n_frames=8)
        if k == 3:
            early = max(row_relerr(s.qpos, env.qpos).max(), row_relerr(s.qvel, env.qvel).max())
        # touchdown detector: vertical root velocity stops decreasing (impact)
        for arr, fst in ((env.qvel, first_ref), (s.qvel, first)):
            hit = (fst < 0) & (arr[:, 2] > -9.81 * 0.016 * (k + 1) * 0.5)
            fst[hit] = k
    assert early <= TOL_32 * 5, early
    assert (np.abs(first - first_ref) <= 1).mean() >= 0.9, (first, first_ref)
    zr, z = env.qpos[:, 2], s.qpos[:, 2]
    assert np.isfinite(s.qpos).all() and (s.status & 3 == 0).all()
    assert abs(np.median(z) - np.median(zr)) < 0.02 and np.abs(z - zr).mean() < 0.05, (np.median(z), np.median(zr))
    assert z.min() > 0.02  # nothing fell through the floor
    return float(early), float(np.abs(z - zr).mean())


def mimic_setup(model):
    table = mocap.MotionTable.load(ASSETS / "motions.npz")
    mdesc, keep = mimic.make_mimic_desc(model, table)
    return table, mdesc, keep


def check_mimic(backend, model_grav, table, mdesc, N=28, steps=3, seed=23):
    """Mimic mode (build-defined, parity unpinned): reference-state initialisation, DeepMimic reward, termination,
    metrics and the block-reduced statistics against the oracle's restatement of the same spec."""
    orc = _oracle(model_grav)
    env = OracleEnv(orc, N)
    rng = np.random.default_rng(seed)
    mid = (np.arange(N) % 7).astype(np.int32)
    t0 = (rng.uniform(0, 1, N) * table.header[mid, 2]).astype(np.float32)
    t0[:7] = 0.0                                        # exactly on the first frame
    t0[7:14] = table.header[mid[7:14], 2] * 0.999       # just before the end (wrap / clamp handling)
    env.reset_mimic(mdesc, mid, t0)
    s = backend.new_state(N)
    s.motion_id[:], s.ref_t0[:] = mid, t0
    backend.reset(s, reward_mode=1)
    # fp32 blend factor (t/dt - frame) and near-gimbal Euler angles cost a few 1e-6 relative
    assert relerr(s.qpos, env.qpos) <= 2e-5 and relerr(s.qvel, env.qvel) <= 5e-5
    assert np.array_equal(s.qpos, s.qpos_lag) and (s.time == 0).all() and (s.metrics == 0).all()
    # the first-frame envs start exactly on the table's qpos (root pos, unit root quat, Euler angles)
    fr = table.frames[table.header[mid[:7], 0].astype(int)]
    assert np.abs(s.qpos[:7, :7] - fr[:, :7]).max() <= 2e-6 and np.abs(s.qpos[:7, 7:] - fr[:, 7:35]).max() <= 2e-5
    kp, kd = util_data.generate_kp_kd_gains()
    for k in range(steps):
        tgt = env.qpos[:, 7:].astype(np.float32)
        st_ref = env.step_mimic(tgt, kp, kd)
        s.stats[:] = 0
        backend.step(s, tgt, kp, kd, reward_mode=1)
        assert s.stats[5] == N and (s.reward[:, 0] > 0).all() and (s.reward[:, 0] <= 1.0 + 1e-6).all()
        if k == 0:  # first env step: per-env parity (8 contact-rich substeps)
            assert row_relerr(s.qpos, env.qpos).max() <= 1e-4
            assert np.abs(s.reward - env.reward).max() <= 1e-3 and (s.done == env.done).all()
            assert np.abs(s.metrics - env.metrics).max() <= 2e-3
            assert np.abs(s.stats[:6] - st_ref[:6]).max() <= 1e-2
        else:       # later: contact-rich rollouts amplify fp32 noise -> statistics + the well-conditioned majority
            good = row_relerr(s.qpos, env.qpos) <= 1e-3
            assert good.mean() >= 0.9, good.mean()
            assert np.abs(s.reward[good] - env.reward[good]).max() <= 2e-2
            assert abs(s.stats[0] - st_ref[0]) / N <= 5e-3              # mean reward
            assert abs(s.stats[2] - st_ref[2]) <= max(2, 0.03 * N)       # terminations
            assert abs(s.stats[3] - st_ref[3]) <= max(2, 0.03 * N)       # falls
    # a pose held exactly on the reference with zero velocity error right after reset scores close to the maximum
    return float(env.reward[:, 0].mean())


def check_hands_against_thighs(backend, model_nograv, N=16, steps=12, seed=29):
    """Body-vs-body contacts in the regime a rollout reaches: the arms are driven inwards (shoulder_x targets) so that
    the wrist spheres press on the thigh capsules.  Per-env parity with the oracle while the contact is shallow, and the
    contact does its job: the wrists do not sink into the thighs (they do with self_collision = 0)."""
    orc = _oracle(model_nograv)
    env = OracleEnv(orc, N)
    env.reset()
    s = backend.new_state(N)
    backend.reset(s)
    rng = np.random.default_rng(seed)
    kp, kd = util_data.generate_kp_kd_gains()
    tgt = np.zeros((N, 28), np.float32)
    names = model_nograv.jnt_names[1:]
    tgt[:, names.index("right_shoulder_x")] = rng.uniform(0.15, 0.45, N)    # right arm towards the body
    tgt[:, names.index("left_shoulder_x")] = -rng.uniform(0.15, 0.45, N)
    tgt[:, names.index("right_elbow")] = rng.uniform(0.0, 0.3, N)
    worst, touched, per_env = 0.0, np.zeros(N, bool), np.zeros(N)
    for k in range(steps):
        env.step(tgt, kp, kd, pd_mode=1, n_frames=8)
        backend.step(s, tgt, kp, kd, pd_mode=1, n_frames=8)
        pen = _self_penetration(model_nograv, env.qpos)
        touched |= _self_contacts(model_nograv, env.qpos) > 0
        per_env = np.maximum(per_env, row_relerr(s.qpos, env.qpos))
        worst = float(per_env.max())
        assert pen.min() > -5e-3, (k, pen.min())           # soft contact: millimetres of penetration at the impact, then it settles
    assert touched.mean() >= 0.75, touched.mean()           # the case exercises the rows
    # ~100 contact substeps in fp32 vs the float64 oracle: the 32-substep tolerance of the north star for every env,
    # a tenth of it for the typical one (the host emulation and the GPU land on the same numbers)
    assert worst <= TOL_32 and np.median(per_env) <= 1e-4, (worst, np.median(per_env))
    assert (s.status & 7 == 0).all()
    return worst


def _self_contacts(model, qpos):
    orc = _oracle(model)
    out = np.zeros(len(qpos), int)
    for e, q in enumerate(qpos):
        d = orc.new_data()
        d.qpos[:35] = q
        orc.forward(d)
        out[e] = int(d.nself[0])
    return out


def check_autoreset(backend, model_grav, table, mdesc, N=21, seed=31):
    """pbrl_autoreset (brax AutoResetWrapper semantics): finished envs restart -- qpos0 in reference mode, the reference
    pose at the given clip phase in mimic mode -- with the first observation of the new episode in `obs`; the other envs
    and the terminal reward / done / metrics are untouched; the step count of the restarted envs is zeroed."""
    rng = np.random.default_rng(seed)
    kp, kd = util_data.generate_kp_kd_gains()
    # reference mode
    s = backend.new_state(N)
    backend.reset(s)
    tau = rng.uniform(-0.5, 0.5, (N, 28)).astype(np.float32)
    backend.step(s, tau, kp, kd, pd_mode=0)
    before = {k: getattr(s, k).copy() for k in ("qpos", "qvel", "time", "warm", "ctrl", "obs", "reward", "done", "metrics")}
    done = (np.arange(N) % 3 == 0).astype(np.float32)
    steps = np.full(N, 7, np.int32)
    backend.autoreset(s, done, steps=steps)
    r, k = done != 0, done == 0
    q0 = model_grav.qpos0.astype(np.float32)
    assert np.array_equal(s.qpos[r], np.tile(q0, (r.sum(), 1))) and (s.qvel[r] == 0).all() and (s.time[r] == 0).all()
    assert (s.warm[r] == 0).all() and (s.ctrl[r] == 0).all() and np.array_equal(s.qpos_lag[r], s.qpos[r])
    assert np.array_equal(s.obs[r], np.concatenate([s.qpos[r], s.qvel[r]], 1))
    for name in ("qpos", "qvel", "time", "warm", "ctrl", "obs"):
        assert np.array_equal(getattr(s, name)[k], before[name][k]), name
    for name in ("reward", "done", "metrics"):
        assert np.array_equal(getattr(s, name), before[name]), name
    assert (steps[r] == 0).all() and (steps[k] == 7).all()
    # mimic mode: restart at a fresh phase == a reset at that phase
    s = backend.new_state(N)
    mid = (np.arange(N) % 7).astype(np.int32)
    s.motion_id[:] = mid
    s.ref_t0[:] = (rng.uniform(0, 1, N) * table.header[mid, 2]).astype(np.float32)
    backend.reset(s, reward_mode=1)
    backend.step(s, s.qpos[:, 7:].copy(), kp, kd, reward_mode=1)
    rew, met = s.reward.copy(), s.metrics.copy()
    nxt = (rng.uniform(0, 1, N) * table.header[mid, 2]).astype(np.float32)
    keep_q = s.qpos.copy()
    backend.autoreset(s, done, next_ref_t0=nxt, reward_mode=1)
    fresh = backend.new_state(N)
    fresh.motion_id[:], fresh.ref_t0[:] = mid, nxt
    backend.reset(fresh, reward_mode=1)
    assert np.array_equal(s.qpos[r], fresh.qpos[r]) and np.array_equal(s.qvel[r], fresh.qvel[r])
    assert np.array_equal(s.ref_t0[r], nxt[r]) and np.array_equal(s.obs[r], fresh.obs[r]) and (s.time[r] == 0).all()
    assert np.array_equal(s.qpos[k], keep_q[k]) and np.array_equal(s.reward, rew) and np.array_equal(s.metrics, met)


def check_replay_rollout(backend, model, F=39, N=50, K=7, seed=37):
    """pbrl_replay_rollout: env e shows frame (frame0[e] + k) mod F at step k; every (step, env) item equals
    HumanoidReplay.step at that frame (obs bit-exact echo, zeros / ones, body poses vs the oracle's FK)."""
    rng = np.random.default_rng(seed)
    qpos, qvel = random_states(model, F, seed, vel_scale=1.0, joint_frac=0.9)
    q32, v32 = qpos.astype(np.float32), qvel.astype(np.float32)
    frame0 = rng.integers(0, F, N).astype(np.int32)
    out = backend.replay_rollout(q32, v32, frame0, N, K)
    env = OracleEnv(_oracle(model), F)
    _, _, _, xpos_r, xquat_r = env.replay(q32, v32, want_poses=True)
    for k in range(K):
        f = (frame0 + k) % F
        assert np.array_equal(out["obs"][k], np.concatenate([q32[f], v32[f]], 1))
        assert relerr(out["xpos"][k], xpos_r[f]) <= 2e-6 and relerr(out["xquat"][k], xquat_r[f]) <= 2e-6
    assert (out["reward"] == 0).all() and (out["done"] == 1).all()
    out2 = backend.replay_rollout(q32, v32, None, N, K)   # frame0 = env id
    k2 = min(2, K - 1)
    assert np.array_equal(out2["obs"][k2][:, :35], q32[(np.arange(N) + k2) % F])


def check_reward_after_one_substep(backend, model_grav, table, mdesc, N=256, seed=3):
    """North star: 'fp32 relative error <= 1e-5 on qpos/qvel and reward after one substep'.  The reward half of it, in mimic
    mode: reference-state initialisation over the seven clips (feet on the floor), one substep of stable PD towards random
    targets, reward (total, pose term, end-effector term) against the float64 oracle <= 1e-5, same termination flags.  (The
    pose term needs the geodesic angle in its well-conditioned form: through the arc cosine it was 2.6e-5 off.)"""
    rng = np.random.default_rng(seed)
    mid = (np.arange(N) % 7).astype(np.int32)
    t0 = (rng.uniform(0, 1, N) * table.header[mid, 2]).astype(np.float32)
    kp, kd = util_data.generate_kp_kd_gains()
    lo, hi = model_grav.jnt_range[1:, 0], model_grav.jnt_range[1:, 1]
    s = backend.new_state(N)
    s.motion_id[:], s.ref_t0[:] = mid, t0
    backend.reset(s, reward_mode=1)
    env = OracleEnv(_oracle(model_grav), N)
    env.reset_mimic(mdesc, mid, t0)
    for n in ("qpos", "qvel", "qpos_lag", "qvel_lag"):   # identical initial states: the oracle starts from the back end's fp32 rows
        getattr(env, n)[:] = getattr(s, n)
    tgt = rng.uniform(lo, hi, (N, 28)).astype(np.float32)
    env.step_mimic(tgt, kp, kd, pd_mode=3, pd_dt=0.002, n_frames=1)
    backend.step(s, tgt, kp, kd, pd_mode=3, pd_dt=0.002, n_frames=1, reward_mode=1)
    err = np.abs(np.asarray(s.reward, np.float64) - env.reward)
    assert err.max() <= TOL_1 and (s.done == env.done).all(), (err.max(0), int((s.done != env.done).sum()))
    assert np.abs(s.metrics - env.metrics).max() <= 1e-5
    return float(err.max())


def check_replay_rollout_paths_agree(backend, model, F=39, N=256, K=8, seed=40):
    """pbrl_replay_rollout has two forms: kinematics per item (few items per clip row), or once per clip row into a frame table
    that the rollout then gathers from (many items per row).  Same replay_item_fk arithmetic: the poses must be bit-identical."""
    rng = np.random.default_rng(seed)
    qpos, qvel = random_states(model, F, seed, vel_scale=1.0, joint_frac=0.9)
    q32, v32 = qpos.astype(np.float32), qvel.astype(np.float32)
    per_item = backend.replay_rollout(q32, v32, np.arange(F, dtype=np.int32), F, 1)       # F items < 4 F: one item per row
    frame0 = rng.integers(0, F, N).astype(np.int32)
    assert N * K >= 4 * F
    table = backend.replay_rollout(q32, v32, frame0, N, K)                                 # the frame-table form
    for k in range(K):
        f = (frame0 + k) % F
        for name in ("obs", "xpos", "xquat", "reward", "done"):
            assert np.array_equal(table[name][k], per_item[name][0][f]), (name, k)


def check_contact_rich_statistics(backend, model_grav, table, mdesc, N=1024, steps=24, seed=43):
    """North star: 'contact-rich rollouts are additionally checked by matching termination steps and reward statistics'.
    The bench workload (reference-state init over all clips, random target angles, stable PD, mimic reward) on the back end
    and on the oracle (fp32 build, all host threads): per-step mean reward, termination-step histogram, fall fraction."""
    orc = Oracle(model_grav, "f32")
    env = OracleEnv(orc, N)
    rng = np.random.default_rng(seed)
    mid = (np.arange(N) % 7).astype(np.int32)
    t0 = (rng.uniform(0, 1, N) * table.header[mid, 2]).astype(np.float32)
    env.reset_mimic(mdesc, mid, t0)
    s = backend.new_state(N)
    s.motion_id[:], s.ref_t0[:] = mid, t0
    backend.reset(s, reward_mode=1)
    kp, kd = util_data.generate_kp_kd_gains()
    lo, hi = model_grav.jnt_range[1:, 0], model_grav.jnt_range[1:, 1]
    first_ref, first = np.full(N, steps), np.full(N, steps)
    rew_ref, rew = [], []
    for k in range(steps):
        tgt = rng.uniform(lo, hi, (N, 28)).astype(np.float32)
        env.step_mimic(tgt, kp, kd)
        s.stats[:] = 0
        backend.step(s, tgt, kp, kd, reward_mode=1)
        rew_ref.append(env.reward[:, 0].mean())
        rew.append(s.reward[:, 0].mean())
        first_ref = np.where((first_ref == steps) & (env.done != 0), k, first_ref)
        first = np.where((first == steps) & (s.done != 0), k, first)
    rew_ref, rew = np.array(rew_ref), np.array(rew)
    assert np.isfinite(s.qpos).all() and (s.status & 1 == 0).all()
    assert np.abs(rew - rew_ref).max() <= 0.01 + 0.05 * rew_ref.max(), (rew, rew_ref)
    h_ref = np.bincount(first_ref, minlength=steps + 1) / N
    h = np.bincount(first, minlength=steps + 1) / N
    assert np.abs(np.cumsum(h) - np.cumsum(h_ref)).max() <= 0.04, (np.cumsum(h), np.cumsum(h_ref))   # termination-step CDF
    assert abs((first < steps).mean() - (first_ref < steps).mean()) <= 0.03
    assert (first == first_ref).mean() >= 0.8, (first == first_ref).mean()                               # per env: same termination step for most
    return float(np.abs(rew - rew_ref).max()), float((first == first_ref).mean())


def check_hovering_foot_boxes(backend, model_grav):
    """Foot boxes that hover inside the 1 mm margin without penetrating the floor: mjx plane_convex masks every corner out
    (`support > max(0, max(support) - 1e-3)` is false everywhere) and `_manifold_points` then adds -1e6 to every score in
    float32.  The squared distances of the corners across the foot's long side (4 x 0.0885^2 = 0.0313 > half a float32 step at
    1e6) survive that rounding, the others collapse: the picks are corner 0, corner 4, 0, 0, i.e. TWO contacts per foot (exact
    arithmetic would give four, a masked arg-max one).  Oracle and back end must both list exactly those, at several tilts."""
    orc = _oracle(model_grav)
    rng = np.random.default_rng(77)
    q, tot = [], 0
    for k in range(16):
        qpos = np.array(model_grav.qpos0, np.float64)
        ang = rng.uniform(-0.004, 0.004, 3) if k else np.zeros(3)   # small tilts keep all eight corners inside the margin band
        qq = np.array([1.0, *(0.5 * ang)])
        qpos[3:7] = qq / np.linalg.norm(qq)
        d = orc.new_data()
        d.qpos[:35] = qpos.astype(np.float32)
        orc.forward(d)
        nf = int(d.ncon[0]) - int(d.nself[0])
        box = [c for c in range(nf) if int(model_grav.geom_type[int(d.con_geom[c])]) == 6]
        # lowest corner 0.4 mm above the floor: use the exact corner heights (xpos of the feet), not the manifold's picks
        lo = min(_box_corner_heights(model_grav, d))
        qpos[2] += 0.0004 - lo
        q.append(qpos)
    q32 = np.array(q, np.float32)
    N = len(q32)
    z = np.zeros
    out = backend.debug_forward(q32, z((N, 34), np.float32), z((N, 28), np.float32), z((N, 34), np.float32))
    for e in range(N):
        d = orc.new_data()
        d.qpos[:35] = q32[e]
        orc.forward(d)
        nf = int(d.ncon[0]) - int(d.nself[0])
        act = d.con_dist[:nf] - d.con_includemargin[:nf] < 0
        box_act = [c for c in range(nf) if act[c] and int(model_grav.geom_type[int(d.con_geom[c])]) == 6]
        get = lambda k: out[e, DBG[k][0]:DBG[k][0] + DBG[k][1]]
        assert int(get("ncon_active")[0]) == int(act.sum()), (e, get("ncon_active"), int(act.sum()))
        # the rows sit 0.4 - 1 mm from the floor: a float32 ulp of the corner height (1e-7 m) is 2e-4 of their violation
        assert relerr(get("qacc"), d.qacc[:34]) <= 2e-4 and relerr(get("qfrc_constraint"), d.qfrc_constraint[:34]) <= 2e-4
        if e == 0:  # level feet: corners 0 and 4 of each foot
            assert len(box_act) == 4, box_act
        tot += len(box_act)
    assert tot >= 2 * N
    return tot


def _box_corner_heights(model, d):
    hs = []
    for g in range(len(model.geom_type)):
        if int(model.geom_type[g]) != 6:
            continue
        b = int(model.geom_bodyid[g])
        R = _quat_mat(np.array(d.xquat[4 * b:4 * b + 4]))
        gp = np.array(d.xpos[3 * b:3 * b + 3]) + R @ model.geom_pos[g]
        Rg = R @ _quat_mat(model.geom_quat[g])
        for v in range(8):
            c = np.array([(1 if (v >> (2 - k)) & 1 else -1) * model.geom_size[g][k] for k in range(3)])
            hs.append(float((gp + Rg @ c)[2]))
    return hs


def _quat_mat(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def check_sphere_box_contacts(backend, model_grav, want=12):
    """Sphere-box candidates (torso / wrist spheres against the foot boxes, mjx _sphere_convex): random poses that contain
    such a contact, found with the oracle; the back end must list the same contacts (distances) and resolve them alike."""
    orc = _oracle(model_grav)
    states = []
    for seed in range(100, 220):
        qpos, qvel = random_states(model_grav, 64, seed, gravity_heights=True, vel_scale=1.0, joint_frac=1.0)
        for e in range(64):
            d = orc.new_data()
            d.qpos[:35], d.qvel[:34] = qpos[e].astype(np.float32), qvel[e].astype(np.float32)
            orc.forward(d)
            nc, ns = int(d.ncon[0]), int(d.nself[0])
            kinds = [int(model_grav.geom_type[int(d.con_geom[c])]) for c in range(nc - ns, nc)]
            if 6 in kinds and ns <= 8:
                states.append((qpos[e], qvel[e]))
        if len(states) >= want:
            break
    assert len(states) >= want
    q32 = np.array([s[0] for s in states], np.float32)
    v32 = np.array([s[1] for s in states], np.float32)
    N = len(states)
    out = backend.debug_forward(q32, v32, np.zeros((N, 28), np.float32), np.zeros((N, 34), np.float32))
    nbox, errs, pens = 0, [], []
    for e in range(N):
        d = orc.new_data()
        d.qpos[:35], d.qvel[:34] = q32[e], v32[e]
        orc.forward(d)
        get = lambda k: out[e, DBG[k][0]:DBG[k][0] + DBG[k][1]]
        nc, ns = int(d.ncon[0]), int(d.nself[0])
        nf = nc - ns
        assert int(get("nself")[0]) == ns, (e, get("nself"), ns)
        assert np.abs(np.sort(get("self_dist")[:ns]) - np.sort(d.con_dist[nf:nc])).max() <= 2e-6
        f_ref = np.sort(d.efc_force[28 + 4 * nf:28 + 4 * nc].reshape(ns, 4).sum(1))
        f_got = np.sort(get("self_force")[:4 * ns].reshape(ns, 4).sum(1))
        assert np.abs(f_got - f_ref).max() <= 2e-3 * max(1.0, np.abs(f_ref).max()), (e, f_got, f_ref)
        errs.append(relerr(get("qacc"), d.qacc[:34]))
        pens.append(-min(d.con_dist[nf:nc]))
        nbox += sum(int(model_grav.geom_type[int(d.con_geom[c])]) == 6 for c in range(nf, nc))
    assert nbox >= want
    # random poses overlap by centimetres in several places at once: the typical one is at rounding level, the bound on the
    # others grows with the overlap
    assert np.median(errs) <= 1e-5, np.median(errs)
    assert_qacc_errors_by_depth(errs, pens, "sphere-box")
    return nbox


# ---------------------------------------------------------------------------------------------------------------------
# The BASELINE.json configurations by name: reference-state initialisation from the named clips at
# a uniform phase, stable PD whose targets are the clip's own joint angles at the end of the env step ("--actions reference").
def named_config_setup(model, clips):
    table = mocap.MotionTable.load(ASSETS / "motions.npz")
    mdesc, keep = mimic.make_mimic_desc(model, table, clips)
    table_ids = np.array([table.names.index(c) for c in clips])
    return table, mdesc, keep, table_ids


def _tracking_rollout(backend, orc_env, model, table, mdesc, table_ids, N, steps, seed, n_frames=8, on_step=None, lift=0.0):
    """Reference-state init + `steps` env steps of clip-tracking stable PD on a back end (or None) and on an oracle env."""
    rng = np.random.default_rng(seed)
    nclip = len(table_ids)
    mid = (np.arange(N) % nclip).astype(np.int32)
    dur = table.header[table_ids[mid], 2]
    t0 = (rng.uniform(0, 1, N) * dur).astype(np.float32)
    orc_env.reset_mimic(mdesc, mid, t0)
    s = None
    if backend is not None:
        s = backend.new_state(N)
        s.motion_id[:], s.ref_t0[:] = mid, t0
        backend.reset(s, reward_mode=1)
    if lift:  # the same clip pose, `lift` metres above the floor (no floor contact)
        orc_env.qpos[:, 2] += lift
        orc_env.qpos_lag[:, 2] += lift
        if s is not None:
            s.qpos[:, 2] += np.float32(lift)
            s.qpos_lag[:, 2] += np.float32(lift)
    kp, kd = util_data.generate_kp_kd_gains()
    h = float(model.timestep)
    for k in range(steps):
        tgt = table.reference_joint_angles(table_ids[mid], t0.astype(np.float64) + (k + 1) * n_frames * h).astype(np.float32)
        st_ref = orc_env.step_mimic(tgt, kp, kd, pd_mode=3, pd_dt=h, n_frames=n_frames)
        if s is not None:
            s.stats[:] = 0
            backend.step(s, tgt, kp, kd, pd_mode=3, pd_dt=h, n_frames=n_frames, reward_mode=1)
        if on_step is not None:
            on_step(k, s, orc_env, st_ref)
    return s, mid, t0


def check_config3_punch_tracking(backend, model_nograv, table, mdesc, table_ids, N=32, seed=51):
    """BASELINE configs[2]: final_humanoid_no_gravity.xml, PD tracking of humanoid3d_punch.txt from reference-state
    initialisation, 32 substeps (4 env steps x 8) against the float64 oracle.

    (a) The clip pose lifted 1 m off the floor -- the contact-free dynamics path the config names.  Clip-tracking targets keep
        a few actuators inside the linear range of the ctrl clamp, where the loop gain gear x kp amplifies rounding (the
        reference's quirk): fp32 relative error <= 1e-3 on qpos / qvel for EVERY env after 16 substeps; after
        32 substeps for at least 98 % of the envs (median <= 1e-4, none beyond 5e-2), and not for fewer envs than the
        oracle's own fp32 build manages on the same inputs.  The oracle confirms that no contact row was active.
    (b) The clip pose as it is: the feet rest on the floor plane, so floor rows switch on and off in every env although
        gravity is off, and a foot that hovers inside the margin goes through the float32 rounding of mjx's manifold mask
        (check_hovering_foot_boxes).  The oracle's OWN fp32 build leaves its float64 build by more than 1e-3 in 1 - 2 % of these
        envs; the bound is therefore: median <= 1e-4, at least 94 % of the envs <= 1e-3, no env beyond 5e-2, and a count of envs
        beyond 1e-3 comparable with the fp32 oracle's (<= 2 x + 4: the envelope is the conditioning of the rollout in fp32)."""
    # (a) contact-free
    def both(steps):
        env = OracleEnv(_oracle(model_nograv), N)
        contact = np.zeros(N, bool)

        def on_step(k, s, e, st):
            for i in range(N):
                contact[i] |= _any_contact_row(model_nograv, e.qpos[i])
        s, _, _ = _tracking_rollout(backend, env, model_nograv, table, mdesc, table_ids, N, steps, seed, on_step=on_step, lift=1.0)
        assert np.isfinite(s.qpos).all() and (s.status & 7 == 0).all()
        return np.maximum(row_relerr(s.qpos, env.qpos), row_relerr(s.qvel, env.qvel)), ~contact, env
    e16, free, _ = both(2)
    assert free.mean() >= 0.9, free.mean()
    assert e16[free].max() <= TOL_32, e16[free].max()
    e32, free, env = both(4)
    env32 = OracleEnv(Oracle(model_nograv, "f32"), N)
    _tracking_rollout(None, env32, model_nograv, table, mdesc, table_ids, N, 4, seed, lift=1.0)
    o32 = np.maximum(row_relerr(env32.qpos, env.qpos), row_relerr(env32.qvel, env.qvel))
    n_bad, n_bad32 = int((e32[free] > TOL_32).sum()), int((o32[free] > TOL_32).sum())
    assert np.median(e32[free]) <= 1e-4 and e32[free].max() <= 5e-2, (np.median(e32[free]), e32[free].max())
    assert n_bad <= max(n_bad32 + 1, int(0.02 * N)), (n_bad, n_bad32)
    out = dict(free_envs=int(free.sum()), free_16_max=float(e16[free].max()), free_32_median=float(np.median(e32[free])),
               free_32_beyond_1e3=n_bad, free_32_oracle_f32_beyond_1e3=n_bad32, free_32_max=float(e32[free].max()))
    # (b) as named: feet on the floor
    env = OracleEnv(_oracle(model_nograv), N)
    s, mid, t0 = _tracking_rollout(backend, env, model_nograv, table, mdesc, table_ids, N, 4, seed)
    err = np.maximum(row_relerr(s.qpos, env.qpos), row_relerr(s.qvel, env.qvel))
    env32 = OracleEnv(Oracle(model_nograv, "f32"), N)
    _tracking_rollout(None, env32, model_nograv, table, mdesc, table_ids, N, 4, seed)
    err32 = np.maximum(row_relerr(env32.qpos, env.qpos), row_relerr(env32.qvel, env.qvel))
    n_bad, n_bad32 = int((err > TOL_32).sum()), int((err32 > TOL_32).sum())
    assert np.isfinite(s.qpos).all() and (s.status & 7 == 0).all()
    assert np.median(err) <= 1e-4 and n_bad <= max(1, int(0.06 * N)) and err.max() <= 5e-2, (np.median(err), n_bad, err.max())
    # both are float32 roundings of the same ill-conditioned rollouts: the counts are samples of one distribution, not ordered
    assert n_bad <= 2 * n_bad32 + 4, (n_bad, n_bad32)
    good = err <= TOL_32
    assert np.abs(s.reward[good, 0] - env.reward[good, 0]).max() <= 2e-3 and (s.done[good] == env.done[good]).all()
    out.update(named_median=float(np.median(err)), named_beyond_1e3=n_bad, named_max=float(err.max()), oracle_f32_beyond_1e3=n_bad32)
    return out


def _any_contact_row(model, qpos):
    """True if the float64 oracle sees an active floor or body-vs-body row at this pose."""
    orc = _oracle(model)
    d = orc.new_data()
    d.qpos[:35] = qpos
    orc.forward(d)
    nc = int(d.ncon[0])
    return bool((d.con_dist[:nc] - d.con_includemargin[:nc] < 0).any())


def _floor_active(model, qpos):
    orc = _oracle(model)
    d = orc.new_data()
    d.qpos[:35] = qpos
    orc.forward(d)
    nf = int(d.ncon[0]) - int(d.nself[0])
    return bool((d.con_dist[:nf] - d.con_includemargin[:nf] < 0).any())


def check_config4_backflip_jump(backend, model_grav, table, mdesc, table_ids, N=64, N_stat=512, steps=24, seed=53):
    """BASELINE configs[3]: final_humanoid.xml, contact-rich mimic of humanoid3d_backflip.txt (even env ids) and
    humanoid3d_jump.txt (odd).  First env step (8 contact substeps): per-env parity with the float64 oracle; then the
    statistics the north star names over `steps` env steps: termination-step histogram and per-step reward mean."""
    env = OracleEnv(_oracle(model_grav), N)
    first = {}

    def on_first(k, s, e, st):
        if k == 0:
            first.update(eq=row_relerr(s.qpos, e.qpos), ev=row_relerr(s.qvel, e.qvel), rew=np.abs(s.reward[:, 0] - e.reward[:, 0]).max(),
                         done=(s.done == e.done).all(), stats=np.abs(s.stats[:6] - st[:6]).max())
    _tracking_rollout(backend, env, model_grav, table, mdesc, table_ids, N, 1, seed, on_step=on_first)
    # 8 substeps with feet on the floor: the 32-contact-free-substep bound of the north star is met per env, the typical env
    # sits two orders below it
    assert first["eq"].max() <= TOL_32 and first["ev"].max() <= TOL_32, (first["eq"].max(), first["ev"].max())
    assert np.median(first["eq"]) <= 1e-5, np.median(first["eq"])
    assert first["rew"] <= 1e-3 and first["done"] and first["stats"] <= 1e-2 * max(1, N / 64)
    # statistics (fp32 oracle, all host threads): the rollout the bench times, with the clip-tracking targets
    env32 = OracleEnv(Oracle(model_grav, "f32"), N_stat)
    fr, f, rr, r = np.full(N_stat, steps), np.full(N_stat, steps), [], []

    def on_stat(k, s, e, st):
        rr.append(e.reward[:, 0].mean()); r.append(s.reward[:, 0].mean())
        fr[:] = np.where((fr == steps) & (e.done != 0), k, fr)
        f[:] = np.where((f == steps) & (s.done != 0), k, f)
    s, _, _ = _tracking_rollout(backend, env32, model_grav, table, mdesc, table_ids, N_stat, steps, seed + 1, on_step=on_stat)
    rr, r = np.array(rr), np.array(r)
    assert np.isfinite(s.qpos).all() and (s.status & 1 == 0).all()
    assert np.abs(r - rr).max() <= 0.01 + 0.05 * rr.max(), (r, rr)
    h_ref, h = np.bincount(fr, minlength=steps + 1) / N_stat, np.bincount(f, minlength=steps + 1) / N_stat
    assert np.abs(np.cumsum(h) - np.cumsum(h_ref)).max() <= 0.04, (np.cumsum(h), np.cumsum(h_ref))
    assert (f == fr).mean() >= 0.8, (f == fr).mean()
    return dict(first_qpos=float(first["eq"].max()), first_qvel=float(first["ev"].max()), reward=float(np.abs(r - rr).max()),
                same_termination_step=float((f == fr).mean()))


def check_alt_rate(backend_factory, model_nograv_or_grav, N=16, seed=57):
    """The reference's second rate configuration (agents_env/agent_env_template2.py:49-58: timestep = 1/1200, n_frames = 40,
    env.dt = 1/30): one env step (40 substeps, stable PD with pd_dt = sys.dt) against the float64 oracle."""
    import copy
    m = copy.copy(model_nograv_or_grav)
    m.timestep = 1.0 / 1200.0
    backend = backend_factory(m)
    env = OracleEnv(_oracle(m), N)
    qpos, qvel = random_states(m, N, seed, gravity_heights=False)
    q32, v32 = qpos.astype(np.float32), qvel.astype(np.float32)
    env.set_state(q32, v32)
    s = backend.new_state(N)
    s.set_state(q32, v32)
    rng = np.random.default_rng(seed)
    kp, kd = util_data.generate_kp_kd_gains()
    lo, hi = m.jnt_range[1:, 0], m.jnt_range[1:, 1]
    tgt = rng.uniform(lo, hi, (N, 28)).astype(np.float32)
    h = np.float32(1.0 / 1200.0)
    env.step(tgt, kp, kd, pd_mode=3, pd_dt=float(h), n_frames=40)
    backend.step(s, tgt, kp, kd, pd_mode=3, pd_dt=float(h), n_frames=40)
    eq, ev = row_relerr(s.qpos, env.qpos), row_relerr(s.qvel, env.qvel)
    assert np.allclose(s.time, 40 * h, rtol=1e-5) and np.allclose(s.time, env.time, rtol=1e-5)
    # 40 substeps, <= 1e-3 per env; an env on which the oracle's OWN float32 build leaves its float64 build by more than 3e-4
    # (a joint driven onto its limit inside the 40 substeps: the row switches on one substep apart, and by how much the two
    # float32 computations then differ depends on the rounding, 1e-3 .. 7e-3 across hosts) is held to 2e-2 instead
    env32 = OracleEnv(Oracle(m, "f32"), N)
    env32.set_state(q32, v32)
    env32.step(tgt, kp, kd, pd_mode=3, pd_dt=float(h), n_frames=40)
    e32 = np.maximum(row_relerr(env32.qpos, env.qpos), row_relerr(env32.qvel, env.qvel))
    bound = np.where(e32 > 3e-4, 2e-2, TOL_32)
    err = np.maximum(eq, ev)
    assert (err <= bound).all() and (e32 > 3e-4).sum() <= max(1, N // 32) and np.median(err) <= 2e-5, (err.max(), e32.max(), np.median(err))
    assert (s.status & 7 == 0).all()
    return float(eq.max()), float(ev.max())


def check_solver_rule(backend_factory, model_grav, N=48, seed=71):
    """Solver options of the model (iterations / tolerance / ls_iterations / ls_tolerance) under solver_rule = 1, MuJoCo / MJX's
    stop rule, against the oracle run with the same options -- and how far the default rule (exact
    active set) sits from it on contact-rich states."""
    import copy
    m1 = copy.copy(model_grav)
    m1.solver_rule = 1
    be1, be0 = backend_factory(m1), backend_factory(model_grav)
    rng = np.random.default_rng(seed)
    qpos, qvel = random_states(model_grav, N, seed, gravity_heights=True, vel_scale=1.0, joint_frac=0.6)
    ctrl, warm = rng.uniform(-1, 1, (N, 28)), rng.normal(size=(N, 34))
    q32, v32, c32, w32 = [np.asarray(x, np.float32) for x in (qpos, qvel, ctrl, warm)]
    out1, out0 = be1.debug_forward(q32, v32, c32, w32), be0.debug_forward(q32, v32, c32, w32)
    orc_mjx, orc_exact, orc_mjx32 = Oracle(model_grav, "f64"), _oracle(model_grav), Oracle(model_grav, "f32")   # tolerance 1e-8 / 1e-14 / 1e-8
    get = lambda out, e, k: out[e, DBG[k][0]:DBG[k][0] + DBG[k][1]]
    e_rule1, e_rule0, e_o32, gap, same_iter, nit1, nit0, selfc = [], [], [], [], 0, [], [], []
    for e in range(N):
        res = []
        for orc in (orc_mjx, orc_exact, orc_mjx32):
            d = orc.new_data()
            d.qpos[:35], d.qvel[:34], d.ctrl[:28], d.qacc_warmstart[:34] = q32[e], v32[e], c32[e], w32[e]
            orc.forward(d)
            res.append((np.array(d.qacc[:34], np.float64), int(d.solver_niter[0])))
        selfc.append(int(d.nself[0]) > 0)
        (a_mjx, n_mjx), (a_exact, _), (a_mjx32, n_mjx32) = res
        e_o32.append(relerr(a_mjx32, a_mjx))
        e_rule1.append(relerr(get(out1, e, "qacc"), a_mjx))
        e_rule0.append(relerr(get(out0, e, "qacc"), a_exact))
        gap.append(relerr(a_mjx, a_exact))
        k1 = int(get(out1, e, "niter")[0])
        nit1.append(k1); nit0.append(int(get(out0, e, "niter")[0]))
        same_iter += int(abs(k1 - n_mjx32) <= 1 or abs(k1 - n_mjx) <= 1)
    # the two rules reach the same minimiser: the MJX-rule answer of the float64 oracle sits within 1e-6 of the converged one
    assert max(gap) <= 1e-6, max(gap)
    # per env: rounding level without body-vs-body contacts.  Random poses put limbs inside each other; the Hessian M + D u u^T
    # of such a row is the worse conditioned the deeper the overlap, and a single fp32 Newton step inherits it.  The bound is
    # therefore stated by penetration depth (tools/dev/dense_accuracy.py: kernel, host emulation and the oracle's own fp32 build
    # have the same error distribution per bin -- p95 3e-5 below 10 mm, 6e-5 below 30 mm, 1e-4 below 60 mm, 5e-4 beyond, with
    # isolated outliers up to 1e-1 on any of the three beyond 30 mm): <= 5e-5 without overlap, <= 3e-4 up to 10 mm (all a
    # rollout reaches), <= 2e-3 up to 30 mm; deeper (unphysical) states only through their 90th percentile (<= 1e-3).
    e_rule1, e_rule0, e_o32, selfc = np.array(e_rule1), np.array(e_rule0), np.array(e_o32), np.array(selfc)
    pen = -_self_penetration(model_grav, q32)
    bound = np.where(~selfc, 5e-5, np.where(pen < 0.010, 3e-4, np.where(pen < 0.030, 2e-3, np.inf)))
    deep = selfc & (pen >= 0.030)
    for err in (e_rule1, e_rule0, e_o32):
        assert np.median(err) <= 1e-5 and (err <= bound).all(), (np.median(err), (err / bound).max())
        assert deep.sum() == 0 or np.quantile(err[deep], 0.9) <= 1e-3, np.quantile(err[deep], 0.9)
    # same stopping point as the oracle under the same rule for most envs (+-1): with tolerance 1e-8 the rule compares fp32
    # rounding noise of the cost against 2e-6, so the last iterations are a coin toss on both sides
    assert same_iter >= 0.8 * N, same_iter
    assert max(nit1) <= model_grav.iterations
    # iterations cap honoured: with iterations = 1 the solver stops after one iteration and says so
    mc = copy.copy(m1)
    mc.iterations = 1
    outc = backend_factory(mc).debug_forward(q32, v32, c32, w32)
    assert max(int(get(outc, e, "niter")[0]) for e in range(N)) <= 1
    return dict(rule1_vs_oracle=float(max(e_rule1)), rule0_vs_exact=float(max(e_rule0)), mjx_vs_exact_f64=float(max(gap)),
                mean_iter_rule1=float(np.mean(nit1)), mean_iter_rule0=float(np.mean(nit0)))


def check_box_pair_contacts(backend, model_grav, want=40, seeds=range(400, 440)):
    """The 15 candidates against the foot boxes (capsule-box x 14: mjx capsule_convex, two contact slots each; box-box x 1:
    mjx convex_convex, four slots): random poses that contain such contacts, found with the oracle; the back end must list
    the same contacts (distances, set-wise) and resolve them alike (contact forces, accelerations)."""
    orc = _oracle(model_grav)
    states = []
    for seed in seeds:
        qpos, qvel = random_states(model_grav, 64, seed, gravity_heights=True, vel_scale=1.0, joint_frac=1.0)
        for e in range(64):
            d = orc.new_data()
            d.qpos[:35], d.qvel[:34] = qpos[e].astype(np.float32), qvel[e].astype(np.float32)
            orc.forward(d)
            nc, ns = int(d.ncon[0]), int(d.nself[0])
            kinds = [(int(model_grav.geom_type[int(d.con_geom1[c])]), int(model_grav.geom_type[int(d.con_geom[c])])) for c in range(nc - ns, nc)]
            if any(k in ((3, 6), (6, 6)) for k in kinds) and ns <= 8:
                states.append((qpos[e], qvel[e]))
        if len(states) >= want:
            break
    assert len(states) >= want, len(states)
    q32 = np.array([s[0] for s in states], np.float32)
    v32 = np.array([s[1] for s in states], np.float32)
    N = len(states)
    out = backend.debug_forward(q32, v32, np.zeros((N, 28), np.float32), np.zeros((N, 34), np.float32))
    ncb = nbb = n32 = 0
    errs, derr, pens = [], [], []
    for e in range(N):
        d = orc.new_data()
        d.qpos[:35], d.qvel[:34] = q32[e], v32[e]
        orc.forward(d)
        get = lambda k: out[e, DBG[k][0]:DBG[k][0] + DBG[k][1]]
        nc, ns = int(d.ncon[0]), int(d.nself[0])
        nf = nc - ns
        k_ns = int(get("nself")[0])
        same = k_ns == ns and np.abs(np.sort(get("self_dist")[:ns]) - np.sort(d.con_dist[nf:nc])).max() <= 5e-6
        if not same:
            # Feet pushed centimetres into each other (random poses): the last pick of the four-point manifold is a near tie that
            # rounding decides, so float32 and float64 evaluations of the SAME algorithm keep different candidates (4 % of random
            # deep overlaps, none of the overlaps below 5 mm: tests/test_self_collision.py).  Only such states may differ.
            bb = [c for c in range(nf, nc) if model_grav.geom_type[int(d.con_geom1[c])] == 6]
            assert bb and min(d.con_dist[c] for c in bb) < -5e-3, (e, k_ns, ns, np.sort(get("self_dist")[:k_ns]), np.sort(d.con_dist[nf:nc]))
            n32 += 1
            continue
        # fp32 closest-point / clipping arithmetic on centimetre-deep random overlaps: distances agree to a few microns
        derr.append(np.abs(np.sort(get("self_dist")[:ns]) - np.sort(d.con_dist[nf:nc])).max())
        f_ref = np.sort(d.efc_force[28 + 4 * nf:28 + 4 * nc].reshape(ns, 4).sum(1))
        f_got = np.sort(get("self_force")[:4 * ns].reshape(ns, 4).sum(1))
        assert np.abs(f_got - f_ref).max() <= 5e-3 * max(1.0, np.abs(f_ref).max()), (e, f_got, f_ref)
        errs.append(relerr(get("qacc"), d.qacc[:34]))
        pens.append(-min(d.con_dist[nf:nc]))
        for c in range(nf, nc):
            k = (int(model_grav.geom_type[int(d.con_geom1[c])]), int(model_grav.geom_type[int(d.con_geom[c])]))
            ncb += k == (3, 6)
            nbb += k == (6, 6)
    assert ncb >= want // 2 and nbb >= 3, (ncb, nbb)
    assert max(derr) <= 5e-6 and n32 <= max(1, N // 20), (max(derr), n32)
    assert np.median(errs) <= 2e-5, np.median(errs)
    assert_qacc_errors_by_depth(errs, pens, "box pairs")
    return dict(states=N, capsule_box_contacts=int(ncb), box_box_contacts=int(nbb), deep_foot_overlaps_with_tie_sensitive_manifold=n32, dist_err=float(max(derr)), qacc_median=float(np.median(errs)), qacc_max=float(max(errs)))


def check_quiet_stance(backend, model_grav, settle=900, window=100):
    """Contact self-consistency in the one regime with a closed-form answer (no MJX golden exists for contact rows): the
    humanoid dropped from reset onto its feet and held at qpos0 by a moderate joint PD (400 N m / rad through the actuators,
    well inside the ctrl range, so the loop does not chatter) stands still.  Over the last `window` substeps:
      * the summed normal force carries the body weight within 5 % (45 kg x 9.81);
      * every pyramid edge force is >= 0, i.e. the tangential force of a contact stays inside mu x its normal force,
        and the net horizontal force is a small fraction of the weight (mean < 5 %, never above 10 %);
      * the normal velocity of every active contact is ~0 and penetration stays far below a millimetre.
    Runs on the float64 oracle and, if given, on a back end (stepping its own state; forces from its own stage dump)."""
    W = float(model_grav.total_mass) * 9.81
    assert abs(W - 45 * 9.81) < 1e-6
    gear = model_grav.act_gear
    kp, kd = (400.0 / gear).astype(np.float32), (40.0 / gear).astype(np.float32)
    tgt = np.zeros((1, 28), np.float32)
    orc = Oracle(model_grav, "f64", tolerance=1e-12)
    env = OracleEnv(orc, 1)
    env.reset()
    s = None
    if backend is not None:
        s = backend.new_state(1)
        backend.reset(s)
    fn_o, fn_k, fxy, vn, pen, edge_min = [], [], [], [], [], []
    for k in range(settle + window):
        if k >= settle:   # forces of the forward pass that the next substep is about to take
            d = orc.new_data()
            d.qpos[:35], d.qvel[:34], d.ctrl[:28], d.qacc_warmstart[:34] = env.qpos[0], env.qvel[0], env.ctrl[0], env.warm[0]
            # the PD law runs every substep here (n_frames = 1): the ctrl of the coming substep
            d.ctrl[:28] = np.clip(-kp * env.qpos[0, 7:] - kd * env.qvel[0, 6:], -0.5, 0.5)
            orc.forward(d)
            nc, nef = int(d.ncon[0]), int(d.nefc[0])
            f = np.array(d.efc_force[28:nef]).reshape(nc, 4)
            edge_min.append(f.min() if nc else 0.0)
            fn_o.append(f.sum())
            fxy.append(np.hypot(d.qfrc_constraint[0], d.qfrc_constraint[1]))
            J = np.array(d.efc_J[:nef * 36]).reshape(nef, 36)[28:, :34].reshape(nc, 4, 34)
            act = f.sum(1) > 1e-9
            if act.any():   # normal row = mean of the four pyramid rows (n +- mu t)
                vn.append(np.abs(J[act].mean(1) @ np.array(d.qvel[:34])).max())
            pen.append(float(np.min(d.con_dist[:nc])) if nc else 0.0)
            if s is not None:
                ctrl_k = np.clip(-kp * s.qpos[0, 7:] - kd * s.qvel[0, 6:], -0.5, 0.5)[None].astype(np.float32)
                out = backend.debug_forward(s.qpos, s.qvel, ctrl_k, s.warm)
                fn_k.append(out[0, DBG["con_force"][0]:DBG["con_force"][0] + DBG["con_force"][1]].sum() +
                            out[0, DBG["self_force"][0]:DBG["self_force"][0] + DBG["self_force"][1]].sum())
        env.step(tgt, kp, kd, pd_mode=1, n_frames=1)
        if s is not None:
            backend.step(s, tgt, kp, kd, pd_mode=1, n_frames=1)
    assert np.abs(env.qvel[0]).max() < 0.3 and abs(env.qpos[0, 2] - 0.8815) < 2e-3          # standing, not collapsing
    assert abs(np.mean(fn_o) / W - 1.0) <= 0.05, np.mean(fn_o) / W
    assert min(edge_min) >= 0.0 and max(fxy) <= 0.10 * W and np.mean(fxy) <= 0.05 * W, (min(edge_min), max(fxy) / W, np.mean(fxy) / W)
    assert max(vn) <= 0.02 and min(pen) >= -2e-4, (max(vn), min(pen))
    out = dict(oracle_weight_ratio=float(np.mean(fn_o) / W), net_horizontal_over_weight=float(max(fxy) / W), max_normal_velocity=float(max(vn)),
               deepest_penetration=float(min(pen)))
    if s is not None:
        assert abs(np.mean(fn_k) / W - 1.0) <= 0.05, np.mean(fn_k) / W
        assert abs(s.qpos[0, 2] - env.qpos[0, 2]) < 1e-3 and np.abs(s.qvel[0]).max() < 0.3
        out["backend_weight_ratio"] = float(np.mean(fn_k) / W)
    return out
