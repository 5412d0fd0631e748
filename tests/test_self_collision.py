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


# ---------------------------------------------------------------------------------------------------------------------
# capsule-box and box-box candidates (the 15 pairs against the foot boxes, models/final_humanoid.xml:87,106)
def _rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def _capsule_box(cpos, cmat, r, hl, bpos, bmat, half):
    lib, ctypes = _hook("orc_test_capsule_box")
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    a = [np.ascontiguousarray(x, float).ravel().copy() for x in (cpos, cmat, bpos, bmat, half)]
    dist, pos, nrm = np.zeros(2), np.zeros(6), np.zeros(6)
    lib.orc_test_capsule_box(P(a[0]), P(a[1]), ctypes.c_double(r), ctypes.c_double(hl), P(a[2]), P(a[3]), P(a[4]), P(dist), P(pos), P(nrm))
    return dist, pos.reshape(2, 3), nrm.reshape(2, 3)


def _box_box(c1, m1, h1, c2, m2, h2):
    lib, ctypes = _hook("orc_test_box_box")
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    a = [np.ascontiguousarray(x, float).ravel().copy() for x in (c1, m1, h1, c2, m2, h2)]
    dist, pos, nrm = np.zeros(4), np.zeros(12), np.zeros(12)
    lib.orc_test_box_box(*[P(x) for x in a], P(dist), P(pos), P(nrm))
    return dist, pos.reshape(4, 3), nrm.reshape(4, 3)


FOOT = np.array([0.0885, 0.045, 0.0275])


def test_capsule_box_closed_forms():
    """The oracle's capsule-box function (mjx capsule_convex) on hand-computable placements of a shin-sized capsule
    (radius .05, half length .155) against a foot-sized box."""
    r, hl = 0.05, 0.155
    I = np.eye(3)
    big = np.array([0.5, 0.5, 0.0275])                 # a wide box: the face case without an edge nearby
    # lying flat (axis along x) on the top face, 2 mm deep: two contacts under the segment's end points, normal capsule -> box
    lay = _rot([0, 1, 0], np.pi / 2)                    # capsule z axis -> world x
    d, pos, n = _capsule_box([0.0, 0.0, 0.0275 + r - 0.002], lay, r, hl, np.zeros(3), I, big)
    assert np.allclose(d, [-0.002, -0.002], atol=1e-12) and np.allclose(n, [[0, 0, -1]] * 2)
    assert np.allclose(sorted(pos[:, 0]), [-hl, hl], atol=1e-12) and np.allclose(pos[:, 2], 0.0275 - 0.001, atol=1e-12)
    # the same capsule overhanging the +x side: its axis passes 48 mm (< radius) above the face's +x edge, and MJX then reports
    # ONE contact against that edge (and switches the second slot off) instead of the two face contacts
    d, pos, n = _capsule_box([0.45, 0.0, 0.0275 + r - 0.002], lay, r, hl, np.zeros(3), I, big)
    assert abs(d[0] + 0.002) < 1e-9 and d[1] == 1.0 and np.allclose(n[0], [0, 0, -1], atol=1e-4) and np.allclose(pos[0], [0.5, 0, 0.0275 - 0.001], atol=1e-5)  # (the 1e-6 regulariser of closest_segment_point shifts the edge point by microns)
    # tilted by 20 degrees so that the overhanging end rises clear of the edge: the low end is a face contact, the segment is
    # clipped at the face's side plane (x = 0.5) where the capsule hovers far above the face (slot on, distance positive)
    th = np.deg2rad(20.0)
    tilt = _rot([0, 1, 0], np.pi / 2 - th)              # capsule axis = (cos th, 0, sin th)
    c = np.array([0.40, 0.0, 0.0275 + r - 0.002 + hl * np.sin(th)])
    d, pos, n = _capsule_box(c, tilt, r, hl, np.zeros(3), I, big)
    lo_x = 0.40 - hl * np.cos(th)
    h_clip = (r - 0.002) + (0.5 - lo_x) * np.tan(th)    # height of the axis above the face where it leaves the footprint
    assert np.allclose(sorted(d), [-0.002, h_clip - r], atol=1e-9) and np.allclose(n, [[0, 0, -1]] * 2)
    assert np.allclose(sorted(pos[:, 0]), [lo_x, 0.5], atol=1e-9)
    # hovering 3 mm above: separated along the face normal -> nothing inside the 1 mm margin
    d, pos, n = _capsule_box([0.0, 0.0, 0.0275 + r + 0.003], lay, r, hl, np.zeros(3), I, big)
    assert (d > 0.001).all()
    # standing upright on the middle of the small (foot) box, lower cap 1 mm into the top face: the closest edge is 66 mm from
    # the segment (> radius), so these are face contacts -- the lower end 1 mm deep, the upper end far above the face
    d, pos, n = _capsule_box([0.0, 0.0, 0.0275 + hl + r - 0.001], I, r, hl, np.zeros(3), I, FOOT)
    assert np.allclose(sorted(d), [-0.001, 2 * hl - 0.001], atol=1e-12) and np.allclose(n, [[0, 0, -1]] * 2) and np.allclose(pos[:, :2], 0, atol=1e-12)
    # the same capsule moved to y = 40 mm: now 5 mm from the face's +y edge (49.3 mm from the segment end) -> ONE contact against that edge, second slot off
    d, pos, n = _capsule_box([0.0, 0.04, 0.0275 + hl + r - 0.001], I, r, hl, np.zeros(3), I, FOOT)
    lower, edge_pt = np.array([0.0, 0.04, 0.0275 + r - 0.001]), np.array([0.0, 0.045, 0.0275])
    gap = np.linalg.norm(lower - edge_pt)
    assert abs(-d[0] - (r - gap)) < 1e-9 and d[1] == 1.0
    assert np.allclose(n[0], (edge_pt - lower) / gap, atol=1e-6)                       # from the capsule towards the edge
    assert np.allclose(pos[0], 0.5 * (edge_pt + lower + (edge_pt - lower) / gap * r), atol=1e-6)
    # far away: both slots off
    d, pos, n = _capsule_box([0.5, 0.5, 0.5], I, r, hl, np.zeros(3), I, FOOT)
    assert (d > 0.001).all()
    # frame invariance: rotating and translating both geoms together moves the contacts with them
    R, t = _rot([1, 2, 3], 0.7), np.array([0.3, -0.2, 0.5])
    d0, p0, n0 = _capsule_box([0.02, 0.01, 0.0275 + r - 0.002], lay, r, hl, np.zeros(3), I, big)
    d1, p1, n1 = _capsule_box(R @ np.array([0.02, 0.01, 0.0275 + r - 0.002]) + t, R @ lay, r, hl, t, R, big)
    assert np.allclose(d0, d1, atol=1e-12) and np.allclose(p1, p0 @ R.T + t, atol=1e-12) and np.allclose(n1, n0 @ R.T, atol=1e-12)


def test_box_box_closed_forms():
    """The oracle's box-box function (mjx convex_convex: separating axes + face manifold / edge contact) on hand-computable
    placements of two foot-sized boxes."""
    I = np.eye(3)
    # one foot on top of the other, 2 mm deep, shifted by (2 cm, 1 cm): contacts at corners of the overlap rectangle.  With
    # exactly parallel edges the last pick of mjx _manifold_points is a tie between a new corner and one already taken (both
    # are equally far from the triangle's edge), so three or four distinct corners come back; the yawed case below gives four.
    c2 = np.array([0.02, 0.01, 2 * 0.0275 - 0.002])
    d, pos, n = _box_box(np.zeros(3), I, FOOT, c2, I, FOOT)
    on = d < 0
    assert on.sum() >= 3 and np.allclose(d[on], -0.002, atol=1e-12) and (d[~on] == 1.0).all() and np.allclose(n, [[0, 0, 1]] * 4)  # normal: box 1 -> box 2
    corners = {(round(x, 9), round(y, 9)) for x in (0.02 - 0.0885, 0.0885) for y in (0.01 - 0.045, 0.045)}
    got = {(round(x, 9), round(y, 9)) for x, y in pos[on, :2]}
    assert got <= corners and len(got) == on.sum()
    assert np.allclose(pos[on, 2], 0.0275 - 0.001, atol=1e-12)                              # midway between the two faces
    # side by side along y with a 5 mm gap: separated, nothing inside the margin
    d, pos, n = _box_box(np.zeros(3), I, FOOT, [0, 2 * 0.045 + 0.005, 0], I, FOOT)
    assert (d > 0.001).all()
    # side by side, overlapping 1 mm along y: the y faces are the reference / incident pair
    d, pos, n = _box_box(np.zeros(3), I, FOOT, [0, 2 * 0.045 - 0.001, 0], I, FOOT)
    on = d < 0
    assert on.sum() >= 3 and np.allclose(d[on], -0.001, atol=1e-12) and np.allclose(n, [[0, 1, 0]] * 4)
    assert set(np.round(np.abs(pos[on, 0]), 9)) == {0.0885} and set(np.round(np.abs(pos[on, 2]), 9)) == {0.0275}
    # edge against edge: box 2 rolled 45 degrees about x and yawed 90 degrees, lowered until its bottom edge dips 1 mm into the
    # top +x... the two closest edges cross at right angles -> one contact, the other three slots off
    R2 = _rot([0, 0, 1], np.pi / 2) @ _rot([1, 0, 0], np.pi / 4)
    R1 = _rot([1, 0, 0], np.pi / 4)
    reach1 = (0.045 + 0.0275) / np.sqrt(2)              # height of box 1's top edge (along x) above its centre
    reach2 = (0.045 + 0.0275) / np.sqrt(2)
    c2 = np.array([0.0, 0.0, reach1 + reach2 - 0.001])
    d, pos, n = _box_box(np.zeros(3), R1, FOOT, c2, R2, FOOT)
    assert abs(d[0] + 0.001) < 1e-9 and (d[1:] == 1.0).all()
    off = (0.045 - 0.0275) / np.sqrt(2)                 # the rolled boxes' extreme edges sit this far off their centre planes
    assert np.allclose(n[0], [0, 0, 1], atol=1e-9) and np.allclose(np.abs(pos[0]), [off, off, reach1 - 0.0005], atol=1e-5)
    # frame invariance
    R, t = _rot([3, 1, 2], 1.1), np.array([-0.4, 0.2, 0.9])
    c2 = np.array([0.02, 0.01, 2 * 0.0275 - 0.002])
    Rz = _rot([0, 0, 1], 0.3)
    d0, p0, n0 = _box_box(np.zeros(3), I, FOOT, c2, Rz, FOOT)
    d1, p1, n1 = _box_box(t, R, FOOT, R @ c2 + t, R @ Rz, FOOT)
    assert (d0 < 0).sum() == 4 and np.allclose(d0, d1, atol=1e-10) and np.allclose(p1, p0 @ R.T + t, atol=1e-10) and np.allclose(n1, n0 @ R.T, atol=1e-10)
    # the rotated overlap polygon: all four contact points lie inside both footprints and 2 mm deep
    q = p0 - c2
    assert (np.abs(p0[:, :2]) <= FOOT[:2] + 1e-9).all() and (np.abs(q @ Rz)[:, :2] <= FOOT[:2] + 1e-9).all() and np.allclose(d0, -0.002, atol=1e-12)


def test_kernel_convex_functions_against_the_oracle():
    """The kernel's capsule-box / box-box functions (compiled for the host) against the oracle's on random placements of a
    shin-sized capsule / a second foot around a foot box.  Capsule-box agrees contact by contact.  Box-box: the last of the four
    manifold picks can be a near tie between two valid candidates that fp32 / fp64 rounding decides differently -- in a few
    percent of centimetre-deep random overlaps (the oracle's own fp32 build shows the same rate), well under one percent of
    the overlaps up to 5 mm that a rollout reaches; everything else agrees contact by contact (distance, position, normal)."""
    import ctypes
    from helpers import build_emu
    from oracle.oracle import HERE
    emu = ctypes.CDLL(str(build_emu()))
    o64, o32 = ctypes.CDLL(str(HERE / "liboracle_f64.so")), ctypes.CDLL(str(HERE / "liboracle_f32.so"))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rng = np.random.default_rng(3)

    def rot():
        q = rng.normal(size=4)
        w, x, y, z = q / np.linalg.norm(q)
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    margin = 1e-3
    stats = dict(bb=[0, 0], bb_shallow=[0, 0], bb_o32=[0, 0], cb=[0, 0])
    for case in range(6000):
        R1, R2, c2 = rot(), rot(), rng.normal(size=3) * 0.09
        # box against box
        a = [np.ascontiguousarray(x, float).ravel().copy() for x in (np.zeros(3), R1, FOOT, c2, R2, FOOT)]
        b = [x.astype(np.float32) for x in a]
        d64, p64, n64 = np.zeros(4), np.zeros(12), np.zeros(12)
        o64.orc_test_box_box(*[P(x) for x in a], P(d64), P(p64), P(n64))
        on64 = np.sort(d64[d64 < margin])
        if len(on64):
            out = np.zeros(32, np.float32)
            emu.emu_box_box(*[P(x) for x in b], P(out))
            dk = out.reshape(4, 8)[:, 3]
            d32, p32, n32 = np.zeros(4, np.float32), np.zeros(12, np.float32), np.zeros(12, np.float32)
            o32.orc_test_box_box(P(b[0]), P(b[1]), P(a[2]), P(b[3]), P(b[4]), P(a[5]), P(d32), P(p32), P(n32))
            for name, dd, keys in (("k", dk, ("bb", "bb_shallow")), ("o32", d32, ("bb_o32",))):
                on = np.sort(dd[dd < margin])
                bad = len(on) != len(on64) or np.abs(on - on64).max() > 5e-6
                stats[keys[0]][0] += bad
                stats[keys[0]][1] += 1
                if len(keys) > 1 and -on64.min() < 5e-3:
                    stats[keys[1]][0] += bad
                    stats[keys[1]][1] += 1
                    if not bad:   # positions and normals of the shallow contacts too (order-free: match by distance)
                        rk = out.reshape(4, 8)[np.argsort(np.where(dk < margin, dk, 9))[:len(on)]]
                        r64 = np.argsort(np.where(d64 < margin, d64, 9))[:len(on)]
                        if len(on) == 1 or np.diff(on64).min() > 1e-5:
                            assert np.abs(rk[:, :3] - p64.reshape(4, 3)[r64]).max() < 5e-6 and np.abs(rk[:, 4:7] - n64.reshape(4, 3)[r64]).max() < 5e-5
        # capsule against box
        r, hl = 0.05, 0.155
        cm, cc = rot(), rng.normal(size=3) * 0.12
        dd64, pp64, nn64 = np.zeros(2), np.zeros(6), np.zeros(6)
        ac = [np.ascontiguousarray(x, float).ravel().copy() for x in (cc, cm, np.zeros(3), R1, FOOT)]
        o64.orc_test_capsule_box(P(ac[0]), P(ac[1]), ctypes.c_double(r), ctypes.c_double(hl), P(ac[2]), P(ac[3]), P(ac[4]), P(dd64), P(pp64), P(nn64))
        onc = np.sort(dd64[dd64 < margin])
        if len(onc):
            out = np.zeros(16, np.float32)
            axis_hl = (cm[:, 2] * hl).astype(np.float32)
            emu.emu_capsule_box(P(ac[0].astype(np.float32)), P(axis_hl), ctypes.c_float(r), P(ac[2].astype(np.float32)), P(ac[3].astype(np.float32)),
                                P(ac[4].astype(np.float32)), P(out))
            dk = out.reshape(2, 8)[:, 3]
            on = np.sort(dk[dk < margin])
            bad = len(on) != len(onc) or np.abs(on - onc).max() > 5e-6
            stats["cb"][0] += bad
            stats["cb"][1] += 1
    assert stats["bb"][1] > 1500 and stats["bb_shallow"][1] > 100 and stats["cb"][1] > 1000, stats
    assert stats["bb_shallow"][0] <= max(1, 0.01 * stats["bb_shallow"][1]), stats   # shallow overlaps: a tie-decided fourth pick at most once in a few hundred
    assert stats["cb"][0] <= 0.01 * stats["cb"][1], stats             # capsule-box: no tie-sensitive selection
    assert stats["bb"][0] <= 0.06 * stats["bb"][1] and abs(stats["bb"][0] - stats["bb_o32"][0]) <= 0.02 * stats["bb"][1], stats
