"""Body-vs-body contacts (sphere / capsule / sphere-box candidates): the candidate census of the model
compiler, the distance functions of the oracle against closed forms, Newton's third law, the kernel logic (host emulation)
against the oracle, and what the reference's own golden traces say about these contacts."""
import numpy as np
import pytest

import parity_cases as pc
from helpers import GOLDEN, EmuBackend, Oracle, OracleEnv, PKG, start_trajectory, util_data


def test_candidate_census(model_grav):
    """SC 32, CC 24, CB 14, SB 10, PC 8, SS 8, PS 5, PB 2, BB 1 -> 104 candidates."""
    import importlib
    mjcf = importlib.import_module(PKG + ".mjcf")
    assert mjcf.pair_census(model_grav) == dict(PS=5, PC=8, PB=2, SS=8, SC=32, SB=10, CC=24, CB=14, BB=1)
    assert len(model_grav.pair_geom1) == 104
    t1, t2 = model_grav.geom_type[model_grav.pair_geom1], model_grav.geom_type[model_grav.pair_geom2]
    assert (t1 <= t2).all()                                            # MJX dispatch order: lower geom type first
    b1, b2 = model_grav.geom_bodyid[model_grav.pair_geom1], model_grav.geom_bodyid[model_grav.pair_geom2]
    assert (b1 != b2).all()
    par = model_grav.body_parent
    assert not any((x != 0 and par[y] == x) or (y != 0 and par[x] == y) for x, y in zip(b1, b2))  # no parent-child pair


def _pose(orc, model, **angles):
    d = orc.new_data()
    names = model.jnt_names
    for n, v in angles.items():
        d.qpos[model.jnt_qposadr[names.index(n)]] = v
    orc.forward(d)
    return d


def test_distance_functions_against_closed_forms(model_nograv):
    """At qpos0 every geom frame is axis-aligned: wrist sphere vs thigh capsule (sphere-capsule), thigh vs thigh
    (parallel capsules) and chest vs neck are hand-computable."""
    m = model_nograv
    orc = Oracle(m, "f64")
    d = orc.new_data()
    orc.forward(d)
    gx = np.array(d.geom_xpos[:m.ngeom * 3]).reshape(-1, 3)
    g = {n: i for i, n in enumerate(m.geom_names)}
    # sphere - capsule: the wrist centre sits 6 mm above the upper end of the thigh segment (vertical capsule, half
    # length 0.15) -> the closest point is that end
    w, h = gx[g["right_wrist"]], gx[g["right_hip"]]
    assert 0 < (w[2] - h[2]) - m.geom_size[g["right_hip"], 1] < 0.01
    want = np.linalg.norm(w - (h + [0, 0, 0.15])) - 0.04 - 0.055
    assert abs(want - 6.2974e-3) < 1e-6                              # arms hang 6.3 mm from the thighs at reset
    assert int(d.nself[0]) == 0                                       # outside the 1 mm margin: no row
    # press the arm inwards until the pair is inside the margin and compare dist / normal / position
    d = _pose(orc, m, right_shoulder_x=0.03)
    assert int(d.nself[0]) == 1
    c = int(d.ncon[0]) - 1
    assert m.geom_names[int(d.con_geom1[c])] == "right_wrist" and m.geom_names[int(d.con_geom[c])] == "right_hip"
    gx = np.array(d.geom_xpos[:m.ngeom * 3]).reshape(-1, 3)
    gm = np.array(d.geom_xmat[:m.ngeom * 9]).reshape(-1, 3, 3)
    w, h, ax = gx[g["right_wrist"]], gx[g["right_hip"]], gm[g["right_hip"]][:, 2]
    t = np.clip(np.dot(w - h, ax), -0.15, 0.15)
    cp = h + t * ax
    n = (cp - w) / np.linalg.norm(cp - w)
    dist = np.linalg.norm(cp - w) - 0.095
    assert abs(d.con_dist[c] - dist) < 1e-7 and dist < 1e-3
    assert np.abs(d.con_frame[9 * c:9 * c + 3] - n).max() < 1e-6     # normal: sphere -> capsule (geom1 -> geom2)
    assert np.abs(d.con_pos[3 * c:3 * c + 3] - (w + n * (0.04 + 0.5 * dist))).max() < 1e-7
    fr = d.con_frame[9 * c:9 * c + 9].reshape(3, 3)
    assert np.abs(fr @ fr.T - np.eye(3)).max() < 1e-12 and np.linalg.det(fr) > 0.999


def test_contact_forces_are_internal(model_nograv):
    """Rows are jac(body2) - jac(body1): the pair of forces cancels on the free joint's translation dofs, so a floating
    humanoid whose arm pushes on its thigh keeps its linear momentum (checked through qfrc_constraint and a rollout)."""
    m = model_nograv
    orc = Oracle(m, "f64", tolerance=1e-14)
    d = _pose(orc, m, right_shoulder_x=0.08, left_shoulder_x=-0.06)
    assert int(d.nself[0]) >= 2 and d.efc_force[28:].max() > 1.0
    assert np.abs(d.qfrc_constraint[:3]).max() < 1e-9 * np.abs(d.qfrc_constraint[:34]).max()
    # angular momentum about the com: the net wrench of the contact forces on the whole tree is zero as well
    assert np.abs(d.qfrc_constraint[3:6]).max() < 1e-9 * np.abs(d.qfrc_constraint[:34]).max()


@pytest.fixture(scope="module")
def emu_nograv(model_nograv):
    return EmuBackend(model_nograv)


def test_kernel_logic_hands_against_thighs(emu_nograv, model_nograv):
    pc.check_hands_against_thighs(emu_nograv, model_nograv, N=6, steps=10)


def test_switch_off_lets_the_hands_through(model_nograv):
    """self_collision = 0 restores the floor-only behaviour (the arms sink into the thighs)."""
    import copy
    m0 = copy.copy(model_nograv)
    m0.self_collision = 0
    kp, kd = util_data.generate_kp_kd_gains()
    tgt = np.zeros((1, 28), np.float32)
    tgt[0, model_nograv.jnt_names[1:].index("right_shoulder_x")] = 0.4
    depth = {}
    for name, mm in (("on", model_nograv), ("off", m0)):
        be = EmuBackend(mm)
        s = be.new_state(1)
        be.reset(s)
        for _ in range(12):
            be.step(s, tgt, kp, kd, pd_mode=1, n_frames=8)
        depth[name] = float(pc._self_penetration(model_nograv, s.qpos.astype(np.float64))[0])
    assert depth["on"] > -3e-3 and depth["off"] < -1e-2, depth


def test_golden_trace_cannot_pin_the_contact_rows(model_nograv):
    """What the reference's own artefacts can say about body-vs-body contacts (no known answer pins them).
    In the T1 run the right wrist first comes within the 1 mm margin of the right thigh in env step 13, and the golden
    trace leaves the contact-free restatement at print 15 -- but not because of that contact: the notebook's PD loop runs
    its motors in saturation and is chaotic in the low-inertia joints, so the fp32 and fp64 builds of the SAME contact-free
    restatement already differ by > 0.05 rad/s (right ankle) at env step 12, and that difference reaches the printed block
    at the 1e-5 level from print 14 on.  Up to print 13 the trace is a known answer (with or without the contact rows);
    beyond, it is not."""
    golden = np.load(GOLDEN / "mass_traces.npz")["T1"]
    kp, kd = util_data.generate_kp_kd_gains()
    coef = start_trajectory()[None]
    idx = [0, 1, 2, 31, 32, 33]
    errs, first_touch, qvel12 = {}, None, {}
    for name, prec, sc in (("off", "f64", False), ("on", "f64", True), ("off32", "f32", False)):
        orc = Oracle(model_nograv, prec, tolerance=1e-14 if prec == "f64" else None, self_collision=sc)
        env = OracleEnv(orc, 1)
        env.reset()
        e = []
        for k in range(17):
            t = np.clip(env.time.astype(np.float32), 1, 3)
            tau, qM, _ = env.pd_torque(coef, kp, kd, 3, 0.016, time=t, want_qM=True)
            e.append(float(np.abs(qM[0][np.ix_(idx, idx)] - golden[k]).max()))
            if k == 12:
                assert np.abs(tau).min() > 0.5 or (np.abs(tau) > 0.5).mean() > 0.9   # (nearly) every motor is saturated
            for f in range(8):
                if name == "on" and first_touch is None and pc._self_contacts(model_nograv, env.qpos)[0] > 0:
                    first_touch = k
                env.step(tau, kp, kd, pd_mode=0, n_frames=1)
            if k == 12:
                qvel12[name] = env.qvel[0].astype(np.float64).copy()
        errs[name] = e
    assert first_touch == 13                                          # the reference's run does contain the contact
    assert max(errs["off"][:14]) <= 1e-5 and max(errs["on"][:14]) <= 1e-5   # known answer up to print 13, contact rows or not
    assert errs["off"][15] > 5e-5                                     # ... and the trace leaves the restatement at print 15
    assert np.abs(qvel12["off32"] - qvel12["off"]).max() > 0.05       # chaos: two precisions of one model, before any contact
    assert max(errs["off32"][:14]) <= 1e-5                            # the printed block is blind to it until print 13


def test_sphere_box_distance_closed_forms():
    """The oracle's sphere-box function on hand-computable placements (box half sizes .0885 x .045 x .0275, axis aligned):
    above a face, beyond an edge, beyond a corner, centre inside."""
    import ctypes
    from oracle.oracle import HERE
    lib = ctypes.CDLL(str(HERE / "liboracle_f64.so"))
    if not hasattr(lib, "orc_test_sphere_box"):
        pytest.skip("oracle built without the test hook")
    half = np.array([0.0885, 0.045, 0.0275])
    eye = np.eye(3).ravel().copy()
    bp = np.zeros(3)

    def run(c, r):
        c = np.asarray(c, float)
        dist, pos, fr = ctypes.c_double(), np.zeros(3), np.zeros(9)
        lib.orc_test_sphere_box(c.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(r), bp.ctypes.data_as(ctypes.c_void_p),
                                eye.ctypes.data_as(ctypes.c_void_p), half.ctypes.data_as(ctypes.c_void_p), ctypes.byref(dist),
                                pos.ctypes.data_as(ctypes.c_void_p), fr.ctypes.data_as(ctypes.c_void_p))
        return dist.value, pos, fr[:3]
    r = 0.04
    # touching the top face from above (penetrating 1 mm): normal points from the sphere down into the box
    d, pos, n = run([0.02, 0.01, 0.0275 + r - 0.001], r)
    assert abs(d + 0.001) < 1e-12 and np.allclose(n, [0, 0, -1]) and np.allclose(pos, [0.02, 0.01, 0.0275 - 0.0005])
    # beyond the +x edge of the top face: closest point on the edge
    c = np.array([0.0885 + 0.02, 0.0, 0.0275 + 0.02])
    d, pos, n = run(c, r)
    want = np.hypot(0.02, 0.02) - r
    assert abs(d - want) < 1e-12 and np.allclose(n, np.array([-0.02, 0, -0.02]) / np.hypot(0.02, 0.02))
    # beyond a corner
    c = half + 0.02
    d, pos, n = run(c, r)
    assert abs(d - (np.sqrt(3) * 0.02 - r)) < 1e-12
    # hovering 0.5 mm above the middle of the top face: the face itself has support >= 0 and is skipped (MJX rule), so the
    # reported distance comes from a side face's edge and is larger than the gap -> no row inside the margin
    d, pos, n = run([0.0, 0.0, 0.0275 + r + 0.0005], r)
    assert d > 0.001


def _hook(name):
    import ctypes
    from oracle.oracle import HERE
    lib = ctypes.CDLL(str(HERE / "liboracle_f64.so"))
    if not hasattr(lib, name):
        pytest.skip("oracle built without the test hooks")
    return lib, ctypes


def test_segment_distances_against_brute_force():
    """The restated closest-point functions (mjx math.closest_segment_point / closest_segment_to_segment_points) against a
    brute-force minimisation over both segments, on random capsule pairs including near-parallel and crossing ones."""
    lib, ctypes = _hook("orc_test_capsule_capsule")
    rng = np.random.default_rng(7)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    s = np.linspace(0, 1, 401)
    worst = 0.0
    for case in range(200):
        a0, a1, b0 = rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.3, 0.3, 3)
        if case % 4 == 0:      # near-parallel
            b1 = b0 + (a1 - a0) * rng.uniform(0.5, 1.5) + 1e-3 * rng.normal(size=3)
        elif case % 4 == 1:    # crossing lines
            mid = a0 + 0.5 * (a1 - a0)
            d = rng.normal(size=3)
            b0, b1 = mid - 0.2 * d + 1e-3 * rng.normal(size=3), mid + 0.2 * d
        else:
            b1 = rng.uniform(-0.3, 0.3, 3)
        r1, r2 = 0.04, 0.055
        dist, pos, fr = ctypes.c_double(), np.zeros(3), np.zeros(9)
        lib.orc_test_capsule_capsule(P(a0), P(a1), ctypes.c_double(r1), P(b0), P(b1), ctypes.c_double(r2), ctypes.byref(dist), P(pos), P(fr))
        pa = a0[None, :] + s[:, None] * (a1 - a0)
        pb = b0[None, :] + s[:, None] * (b1 - b0)
        dmin = np.sqrt(((pa[:, None, :] - pb[None, :, :]) ** 2).sum(-1)).min() - r1 - r2
        # the restatement is exact up to its 1e-6 regularisers; the grid over-estimates by at most its spacing
        assert dist.value <= dmin + 1e-6 and dist.value >= dmin - 2e-3, (case, dist.value, dmin)
        worst = max(worst, abs(dist.value - dmin))
        n = fr[:3]
        assert abs(np.linalg.norm(n) - 1) < 1e-9
    # sphere against segment
    for case in range(100):
        c, b0, b1 = rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.3, 0.3, 3)
        dist, pos, fr = ctypes.c_double(), np.zeros(3), np.zeros(9)
        lib.orc_test_sphere_capsule(P(c), ctypes.c_double(0.04), P(b0), P(b1), ctypes.c_double(0.055), ctypes.byref(dist), P(pos), P(fr))
        pb = b0[None, :] + np.linspace(0, 1, 20001)[:, None] * (b1 - b0)
        dmin = np.sqrt(((pb - c) ** 2).sum(-1)).min() - 0.095
        assert abs(dist.value - dmin) < 2e-5, (case, dist.value, dmin)
