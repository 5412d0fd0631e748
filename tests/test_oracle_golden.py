"""Pin the oracle: the reference's own known answers (notebook stdout) and analytic invariants."""
import numpy as np
import pytest

from helpers import GOLDEN, Oracle, OracleEnv, start_trajectory, util_data

IDX = [0, 1, 2, 31, 32, 33]


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN / "mass_traces.npz")


def _trace(orc, pd_dt, golden, nprint, selected=None, inside=False):
    kp, kd = util_data.generate_kp_kd_gains()
    coef = start_trajectory()[None]
    env = OracleEnv(orc, 1)
    env.reset()
    mask = 0x0FFFFFFF if selected is None else sum(1 << j for j in selected)
    errs = []
    for k in range(nprint):
        t = np.clip(env.time.astype(np.float32), 1, 3)
        tau, qM, _ = env.pd_torque(coef, kp, kd, 3, pd_dt, time=t, want_qM=True)
        errs.append(float(np.abs(qM[0][np.ix_(IDX, IDX)] - golden[k]).max()))
        if inside:
            env.step(coef, kp, kd, pd_mode=3, pd_dt=pd_dt, n_frames=8, pd_time=t, ctrl_mask=mask)
        else:
            env.step(tau, kp, kd, pd_mode=0, n_frames=8)
    return errs


def test_T0_mass_matrix_at_qpos0(model_nograv, golden):
    orc = Oracle(model_nograv, "f64")
    d = orc.new_data()
    orc.forward(d)
    M = d.qM_dense()
    assert np.abs(M[np.ix_(IDX, IDX)] - golden["T1"][0]).max() <= 1.5e-7   # fp32 print precision
    assert np.allclose(M, M.T, atol=1e-12)
    assert np.allclose(np.diag(M)[:3], 45.0, atol=1e-9)                       # total mass
    assert abs(M[33, 33] - 1.00531075) < 1e-8 and abs(M[32, 32] - 1.00539408) < 1e-8 and abs(M[31, 31] - 1.00143333) < 1e-8
    assert abs(M[31, 33] - 1.0125e-3) < 1e-9
    assert np.linalg.eigvalsh(M).min() > 0


def test_T1_eager_stable_pd(model_nograv, golden):
    """testing_pds_env3.ipynb cell 21: eager stable PD (dt = env.dt = 0.016) + 8-substep torque step."""
    errs = _trace(Oracle(model_nograv, "f64", tolerance=1e-14), 0.016, golden["T1"], 11)
    assert max(errs[:4]) <= 5e-7, errs      # SURVEY.md App. C recommended gates
    assert max(errs[4:]) <= 1e-5, errs


def test_T1_f32_oracle(model_nograv, golden):
    errs = _trace(Oracle(model_nograv, "f32"), 0.016, golden["T1"], 11)
    assert max(errs) <= 1e-5, errs


@pytest.mark.parametrize("name,selected", [("T2", None), ("T3", [7, 10, 13, 9]), ("T4", [7, 10, 13, 9, 17, 24])])
def test_T2_T4_jitted_stable_pd(model_nograv, golden, name, selected):
    """testing_pds_env.ipynb cells 24/30/33: stable PD inside the step with dt = sys.dt = 0.002."""
    errs = _trace(Oracle(model_nograv, "f64", tolerance=1e-14), 0.002, golden[name], 6, selected, inside=True)
    assert max(errs[:3]) <= 1e-6 and max(errs) <= 3e-5, errs


def test_ablation_limits_matter(model_nograv, golden):
    """Dropping the joint-limit rows must break the trace: the goldens pin them."""
    orc = Oracle(model_nograv, "f64", tolerance=1e-14)
    orc.set_model_field("jnt_limited", np.zeros(29, np.int32))
    errs = _trace(orc, 0.016, golden["T1"], 5)
    assert max(errs[3:]) > 2e-5, errs


def test_T1_statistics_long_run(model_nograv, golden):
    """Beyond ~20 env steps the closed loop is chaotic; only statistics are comparable."""
    kp, kd = util_data.generate_kp_kd_gains()
    coef = start_trajectory()[None]
    env = OracleEnv(Oracle(model_nograv, "f32"), 1)
    env.reset()
    vals = []
    for k in range(200):
        t = np.clip(env.time.astype(np.float32), 1, 3)
        tau, qM, _ = env.pd_torque(coef, kp, kd, 3, 0.016, time=t, want_qM=True)
        vals.append(qM[0][0, 32])
        env.step(tau, kp, kd, pd_mode=0, n_frames=8)
    g = golden["T1_m0_32"][:200]
    vals = np.array(vals)
    assert np.isfinite(vals).all()
    assert abs(vals.min() - g.min()) < 0.02 and abs(vals.max() - g.max()) < 0.02
    assert np.allclose(qM[0][0, 0], 45.0, atol=1e-4)


def test_momentum_conserved_without_gravity(model_nograv):
    """No gravity, zero ctrl, no floor contact: linear momentum M[0:3,:] qvel is conserved by the step -- also while limbs
    push against each other (body-vs-body contact forces act in equal and opposite pairs)."""
    orc = Oracle(model_nograv, "f64", tolerance=1e-14)
    rng = np.random.default_rng(0)
    d = orc.new_data()
    d.qpos[7:35] = rng.uniform(-0.2, 0.2, 28) * np.sign(model_nograv.jnt_range[1:, 1] + model_nograv.jnt_range[1:, 0] + 1e-9)
    d.qpos[7 + 9] = 0.5; d.qpos[7 + 13] = 0.5; d.qpos[7 + 17] = -0.5; d.qpos[7 + 24] = -0.5   # elbows/knees inside their ranges
    d.qvel[:34] = 0.3 * rng.normal(size=34)
    d.qpos[2] = 2.0                                  # well above the floor: no contact
    orc.forward(d)
    p0 = d.qM_dense()[:3] @ d.qvel[:34]
    for _ in range(50):
        orc.step(d)
    orc.forward(d)
    p1 = d.qM_dense()[:3] @ d.qvel[:34]
    # semi-implicit Euler conserves momentum to O(h) per unit time: 0.1 s of tumbling drifts by ~5e-5 relative
    # (2.3e-4 with the limbs in contact)
    assert np.abs(p1 - p0).max() < 4e-4 * np.abs(p0).max()


def test_static_stance_supports_weight(model_grav):
    """Contact self-consistency (no reference golden exists for contacts): after touchdown the summed normal
    force is about the body weight and penetration stays sub-millimetre (SURVEY.md App. A preamble)."""
    orc = Oracle(model_grav, "f64", tolerance=1e-12)
    d = orc.new_data()
    ratios = []
    for s in range(300):
        orc.step(d)
        if 120 <= s < 300:
            nef = int(d.nefc[0])
            ratios.append(d.efc_force[28:nef].sum() / (45 * 9.81))
    assert 0.85 < np.mean(ratios) < 1.2, np.mean(ratios)
    nc = int(d.ncon[0])
    assert d.con_dist[:nc].min() > -1e-3


def test_solver_optimum_against_an_independent_minimiser(model_grav):
    """The constraint solve is `argmin 1/2 (a - a_s)^T M (a - a_s) + sum_r 1/2 D_r min(0, J_r a - aref_r)^2` (SURVEY.md App. A.9,
    MuJoCo's primal problem with pyramidal cones).  Whatever iteration the oracle runs, its qacc must be THE minimiser of that
    strictly convex cost: checked here without any of the oracle's solver code -- (i) stationarity in numpy, (ii) the same point
    reached by scipy's trust-region Newton from a different start, (iii) the dual reading: row forces f = -D min(0, J a - aref)
    are non-negative, complementary to the row residuals, and qfrc_constraint = J^T f.  States: random contact-rich poses (floor
    contacts, joint limits, limbs touching), the oracle at tolerance 1e-14."""
    from scipy.optimize import minimize
    from helpers import random_states
    orc = Oracle(model_grav, "f64", tolerance=1e-14)
    qpos, qvel = random_states(model_grav, 16, 77, gravity_heights=True, vel_scale=1.0, joint_frac=0.9)
    qpos[:, 2] = np.linspace(0.05, 0.9, 16)   # lying on / sunk into the floor ... standing height
    rng = np.random.default_rng(78)
    nv, rows_seen, self_rows = 34, 0, 0
    for e in range(16):
        d = orc.new_data()
        d.qpos[:35], d.qvel[:34], d.ctrl[:28], d.qacc_warmstart[:34] = qpos[e], qvel[e], rng.uniform(-1, 1, 28), rng.normal(size=34)
        orc.forward(d)
        ne = int(d.nefc[0])
        on = np.array(d.efc_enabled[:ne]) != 0   # rows that exist in MuJoCo's list: limits / contacts inside their margin
        M = np.array(d.qM_dense(), np.float64)
        J = np.array(d.efc_J, np.float64).reshape(-1, 36)[:ne, :nv][on]
        D, aref = np.array(d.efc_D[:ne], np.float64)[on], np.array(d.efc_aref[:ne], np.float64)[on]
        ne = int(on.sum())
        a, a_s = np.array(d.qacc[:nv], np.float64), np.array(d.qacc_smooth[:nv], np.float64)
        assert np.allclose(M, M.T) and (D > 0).all()
        rows_seen += ne
        self_rows += 4 * int(d.nself[0])
        cost = lambda x: 0.5 * (x - a_s) @ M @ (x - a_s) + 0.5 * np.sum(D * np.minimum(J @ x - aref, 0.0) ** 2)
        grad = lambda x: M @ (x - a_s) + J.T @ (D * np.minimum(J @ x - aref, 0.0))
        hess = lambda x: M + (J * (D * (J @ x - aref < 0))[:, None]).T @ J
        scale = np.abs(M @ (a - a_s)).max() + np.abs(M @ a_s).max()
        # (i) stationary point of the documented cost
        assert np.abs(grad(a)).max() <= 1e-9 * scale, (e, np.abs(grad(a)).max(), scale)
        # (ii) an unrelated optimiser, started from zero, ends at the same point and no lower
        res = minimize(cost, np.zeros(nv), jac=grad, hess=hess, method="trust-exact", options={"gtol": 1e-10 * scale, "maxiter": 500})
        assert np.abs(res.x - a).max() <= 1e-6 * max(1.0, np.abs(a).max()), (e, np.abs(res.x - a).max())
        assert cost(a) <= res.fun + 1e-9 * max(1.0, abs(res.fun))
        # (iii) forces
        jar = J @ a - aref
        f = -D * np.minimum(jar, 0.0)
        assert (f >= 0).all() and (ne == 0 or np.abs(f * np.maximum(jar, 0.0)).max() == 0.0)
        assert np.allclose(np.array(d.efc_force[:len(on)], np.float64)[on], f, rtol=1e-10, atol=1e-10 * scale)
        assert np.allclose(np.array(d.qfrc_constraint[:nv], np.float64), J.T @ f, rtol=1e-9, atol=1e-9 * scale)
    assert rows_seen >= 200 and self_rows > 0, (rows_seen, self_rows)   # the set really contains contact rows of both kinds


def test_contact_rows_are_the_kinematic_jacobian_of_the_contact_points(model_grav):
    """A pyramidal contact row is `(n +- mu t_a) . (v_p(body2) - v_p(body1))` as a linear function of qvel, where v_p(b) is the
    velocity of the material point of body b that sits at the contact position (MuJoCo's definition of the contact Jacobian;
    SURVEY.md App. A.8).  Checked without the oracle's Jacobian code: every dof is nudged (world translation, body-frame
    rotation of the free joint, hinge angles), the oracle's forward KINEMATICS moves the two material points, and central
    differences give the rows -- floor contacts and body-vs-body contacts (where the row is a difference of two bodies).
    A limit row is +-1 on its own dof."""
    orc = Oracle(model_grav, "f64", tolerance=1e-14)
    from helpers import random_states
    qpos, qvel = random_states(model_grav, 6, 91, gravity_heights=True, vel_scale=1.0, joint_frac=0.9)
    qpos[:, 2] = [0.1, 0.25, 0.4, 0.55, 0.7, 0.85]
    nv, eps = 34, 1e-6
    gb = np.asarray(model_grav.geom_bodyid)

    def kin(q):
        d = orc.new_data()
        d.qpos[:35] = q
        orc.forward(d)
        return np.array(d.xpos[:42], np.float64).reshape(14, 3), np.array(d.xmat[:126], np.float64).reshape(14, 3, 3)

    def nudge(q, i, h):
        q = q.copy()
        if i < 3:
            q[i] += h
        elif i < 6:   # angular dofs of the free joint are in the body frame: q <- q * exp(h e / 2)
            w, v = q[3], q[4:7]
            e = np.zeros(3); e[i - 3] = 1.0
            dw, dv = np.cos(h / 2), np.sin(h / 2) * e
            q[3], q[4:7] = w * dw - v @ dv, w * dv + dw * v + np.cross(v, dv)
        else:
            q[i + 1] += h
        return q

    rows_floor = rows_self = rows_limit = 0
    for e in range(6):
        d = orc.new_data()
        d.qpos[:35], d.qvel[:34] = qpos[e], qvel[e]
        orc.forward(d)
        q0 = np.array(d.qpos[:35], np.float64)
        xpos, xmat = kin(q0)
        ne, nc = int(d.nefc[0]), int(d.ncon[0])
        J = np.array(d.efc_J, np.float64).reshape(-1, 36)[:ne, :nv]
        nlim = ne - 4 * nc
        for r in range(nlim):   # limit rows: one +-1
            assert np.count_nonzero(J[r]) == 1 and abs(J[r]).max() == 1.0 and np.flatnonzero(J[r])[0] >= 6
            rows_limit += 1
        plus = [kin(nudge(q0, i, eps)) for i in range(nv)]
        minus = [kin(nudge(q0, i, -eps)) for i in range(nv)]
        for c in range(nc):
            if not np.array(d.efc_enabled[nlim + 4 * c:nlim + 4 * c + 4]).any():
                continue
            p = np.array(d.con_pos[3 * c:3 * c + 3], np.float64)
            fr = np.array(d.con_frame[9 * c:9 * c + 9], np.float64).reshape(3, 3)
            mu = float(d.con_mu[c])
            b2, b1 = int(gb[int(d.con_geom[c])]), int(gb[int(d.con_geom1[c])])
            def point_jac(b):
                if b == 0:
                    return np.zeros((3, nv))
                local = xmat[b].T @ (p - xpos[b])
                return np.stack([((plus[i][0][b] + plus[i][1][b] @ local) - (minus[i][0][b] + minus[i][1][b] @ local)) / (2 * eps)
                                 for i in range(nv)], 1)
            Jp = point_jac(b2) - point_jac(b1)
            for k in range(4):
                want = (fr[0] + (1.0 if k % 2 == 0 else -1.0) * mu * fr[1 + k // 2]) @ Jp
                got = J[nlim + 4 * c + k]
                assert np.abs(got - want).max() <= 1e-6 * max(1.0, np.abs(want).max()), (e, c, k, np.abs(got - want).max())
            if b1 == 0:
                rows_floor += 4
            else:
                rows_self += 4
    assert rows_floor >= 40 and rows_self >= 8 and rows_limit >= 6 * 28, (rows_floor, rows_self, rows_limit)


PENDULUM = """<mujoco model="pend">
  <compiler angle="radian" inertiafromgeom="true"/>
  <option timestep="0.002" gravity="0 0 -9.81"/>
  <worldbody>
    <body name="arm" pos="0 0 1">
      <joint name="hinge" type="hinge" axis="0 1 0" range="-0.5 0.5" limited="true"/>
      <geom name="rod" type="capsule" fromto="0 0 0 0 0 -0.5" size="0.05" density="1000"/>
    </body>
  </worldbody>
  <actuator><motor name="m" joint="hinge" gear="1" ctrlrange="-1 1" ctrllimited="true"/></actuator>
</mujoco>
"""


def test_limit_row_known_answers_on_a_one_dof_model(tmp_path):
    """MuJoCo's documented soft-constraint law on the one system where it has a closed form: a single hinge, where the
    diagonal approximation of A = J M^-1 J^T is exact (dof_invweight0 = 1 / M), so the constrained acceleration in row space is
    `(1 - d) a0 + d aref` with `aref = -b v - k d r`, `b = 2 / (dmax timeconst)`, `k = 1 / (dmax^2 timeconst^2 dampratio^2)` and
    the impedance d(r) of solimp = (0.9, 0.95, 0.001, 0.5, 2), solref = (0.02, 1) (the defaults the humanoid uses).  Hand
    numbers: |r| >= width -> d = 0.95, k d = 2631.5789, b = 105.26316; |r| = width / 2 -> d = 0.925 (the sigmoid's midpoint).
    finish_row (oracle) is shared by limit and contact rows, so this pins the solref / solimp arithmetic of both."""
    from helpers import mjcf
    xml = tmp_path / "pend.xml"
    xml.write_text(PENDULUM)
    m = mjcf.compile_mjcf(xml)
    assert (m.nq, m.nv, m.nu) == (1, 1, 1)
    orc = Oracle(m, "f64", tolerance=1e-14)
    hi = 0.5
    for r, v, d_want in ((-0.01, 0.0, 0.95), (-0.01, 0.7, 0.95), (-0.01, -1.3, 0.95), (-0.0005, 0.2, 0.925), (-0.002, 0.0, 0.95)):
        for side in (+1, -1):   # upper / lower limit
            d = orc.new_data()
            d.qpos[0], d.qvel[0], d.ctrl[0] = side * (hi - r), side * v, 0.3
            orc.forward(d)
            M, a0 = float(d.qM[0]), float(d.qacc_smooth[0])
            assert abs(float(m.dof_invweight0[0]) * M - 1.0) <= 1e-9
            J = -side                                  # the row measures the distance to the limit
            assert int(d.nefc[0]) == 1 and float(d.efc_J[0]) == J and float(d.efc_enabled[0]) != 0
            assert abs(float(d.efc_pos[0]) - r) <= 1e-12
            k, b = 1.0 / (0.95 ** 2 * 0.02 ** 2), 2.0 / (0.95 * 0.02)
            aref = -b * (J * side * v) - k * d_want * r
            assert abs(float(d.efc_aref[0]) - aref) <= 1e-9 * abs(aref), (r, v, side, float(d.efc_aref[0]), aref)
            assert abs(float(d.efc_D[0]) - d_want / (1.0 - d_want) * M) <= 1e-9 * M
            want = (1.0 - d_want) * (J * a0) + d_want * aref if J * a0 - aref < 0 else J * a0
            assert abs(J * float(d.qacc[0]) - want) <= 1e-9 * max(1.0, abs(want)), (r, v, side, J * float(d.qacc[0]), want)
    assert abs(k * 0.95 - 2631.5789) < 1e-3 and abs(b - 105.26316) < 1e-4
    # inside the range: the row exists but is switched off, qacc = qacc_smooth
    d = orc.new_data()
    d.qpos[0], d.qvel[0] = 0.2, 0.5
    orc.forward(d)
    assert float(d.efc_enabled[0]) == 0 and float(d.qacc[0]) == float(d.qacc_smooth[0])
