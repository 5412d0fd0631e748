"""MJCF -> ModelConst compiler (host side, setup time, float64 numpy).

The reference builds its model with `mujoco.MjModel.from_xml_path(path)` followed by
`brax.io.mjcf.load_model` (/root/reference/agents_env/agent_env_template.py:44-46).
Neither package exists here, so this module restates the part of MuJoCo's model
compiler the humanoid files need:

  * one global <default> class, <compiler angle="radian" inertiafromgeom="true">,
    <option gravity/timestep>, nested <body>, <joint type=free|hinge>,
    <geom type=plane|sphere|capsule|box> (size / fromto / mass), <contact><exclude>,
    <actuator><motor gear ctrlrange ctrllimited>;
  * inertia from geoms (sphere / capsule / box closed forms, parallel axis);
  * the constants MuJoCo derives at qpos0 (`mj_setConst`): dof_invweight0,
    body_invweight0, stat.meaninertia;
  * the tree tables the CUDA kernels consume (levels, ancestor lists, the sparse
    layout of the joint-space inertia matrix).

Everything render-only (sites, cameras, lights, assets) is ignored.
"""
from __future__ import annotations

import json
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_BOX = 0, 2, 3, 6  # mjtGeom values
JNT_FREE, JNT_HINGE = 0, 3  # mjtJoint values

_GEOM_TYPES = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE, "box": GEOM_BOX}


def _vec(s, n=None, default=None):
    if s is None:
        return None if default is None else np.asarray(default, dtype=np.float64)
    v = np.asarray([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return v


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def axis_angle_quat(axis, angle):
    s = np.sin(0.5 * angle)
    return np.array([np.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def z_to_quat(vec):
    """Quaternion rotating +z onto `vec` (MuJoCo `mjuu_z2quat`)."""
    v = vec / np.linalg.norm(vec)
    z = np.array([0.0, 0.0, 1.0])
    axis = np.cross(z, v)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        return np.array([1.0, 0, 0, 0]) if v[2] > 0 else np.array([0.0, 1.0, 0, 0])
    axis /= s
    ang = np.arctan2(s, v[2])
    return axis_angle_quat(axis, ang)


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


@dataclass
class ModelConst:
    """Flat model description.  Arrays are float64 / int32 numpy."""
    name: str = ""
    nbody: int = 0
    nq: int = 0
    nv: int = 0
    nu: int = 0
    njnt: int = 0
    ngeom: int = 0
    # options
    timestep: float = 0.002
    gravity: np.ndarray = field(default_factory=lambda: np.array([0, 0, -9.81]))
    solref: np.ndarray = field(default_factory=lambda: np.array([0.02, 1.0]))
    solimp: np.ndarray = field(default_factory=lambda: np.array([0.9, 0.95, 0.001, 0.5, 2.0]))
    impratio: float = 1.0
    tolerance: float = 1e-8
    iterations: int = 100
    ls_iterations: int = 50
    ls_tolerance: float = 0.01
    solver_rule: int = 0  # how the Newton solver stops: 0 exact active set, 1 MuJoCo / MJX tolerance rule (include/pbrl.h)
    # bodies
    body_names: list = field(default_factory=list)
    body_parent: np.ndarray = None
    body_pos: np.ndarray = None
    body_mass: np.ndarray = None
    body_ipos: np.ndarray = None
    body_inertia: np.ndarray = None  # (nbody,3,3) about ipos, body axes
    body_jntadr: np.ndarray = None
    body_jntnum: np.ndarray = None
    body_dofadr: np.ndarray = None
    body_dofnum: np.ndarray = None
    body_level: np.ndarray = None
    body_invweight0: np.ndarray = None  # (nbody,2)
    # joints / dofs
    jnt_names: list = field(default_factory=list)
    jnt_type: np.ndarray = None
    jnt_bodyid: np.ndarray = None
    jnt_qposadr: np.ndarray = None
    jnt_dofadr: np.ndarray = None
    jnt_axis: np.ndarray = None
    jnt_pos: np.ndarray = None
    jnt_limited: np.ndarray = None
    jnt_range: np.ndarray = None
    jnt_margin: np.ndarray = None
    dof_bodyid: np.ndarray = None
    dof_jntid: np.ndarray = None
    dof_parentid: np.ndarray = None
    dof_armature: np.ndarray = None
    dof_damping: np.ndarray = None
    dof_invweight0: np.ndarray = None
    dof_depth: np.ndarray = None  # number of proper ancestor dofs
    qpos0: np.ndarray = None
    # sparse inertia layout: row i holds i, parent(i), parent(parent(i)), ...
    M_rowadr: np.ndarray = None
    M_rownnz: np.ndarray = None
    M_row: np.ndarray = None
    M_col: np.ndarray = None
    # geoms
    geom_names: list = field(default_factory=list)
    geom_type: np.ndarray = None
    geom_bodyid: np.ndarray = None
    geom_size: np.ndarray = None
    geom_pos: np.ndarray = None
    geom_quat: np.ndarray = None
    geom_friction: np.ndarray = None
    geom_margin: np.ndarray = None
    geom_gap: np.ndarray = None
    geom_contype: np.ndarray = None
    geom_conaffinity: np.ndarray = None
    # collision candidates (MuJoCo broad-phase filter rules, evaluated once: MJX has no dynamic broad phase)
    pair_geom1: np.ndarray = None  # geom with the lower mjtGeom type first (MJX dispatch order), then the lower id
    pair_geom2: np.ndarray = None
    self_collision: int = 1  # 0: evaluate the floor pairs only
    # actuators
    act_names: list = field(default_factory=list)
    act_dofadr: np.ndarray = None
    act_gear: np.ndarray = None
    act_ctrllimited: np.ndarray = None
    act_ctrlrange: np.ndarray = None
    # stats
    meaninertia: float = 0.0
    total_mass: float = 0.0
    exclude: list = field(default_factory=list)

    # ---- serialisation -------------------------------------------------------------------
    def to_json(self) -> str:
        d = {}
        for k, v in self.__dict__.items():
            if isinstance(v, np.ndarray):
                d[k] = {"dtype": str(v.dtype), "shape": list(v.shape), "data": v.ravel().tolist()}
            else:
                d[k] = v
        return json.dumps(d, indent=0)

    @staticmethod
    def from_json(s: str) -> "ModelConst":
        d = json.loads(s)
        m = ModelConst()
        for k, v in d.items():
            if isinstance(v, dict) and "dtype" in v:
                v = np.asarray(v["data"], dtype=v["dtype"]).reshape(v["shape"])
            setattr(m, k, v)
        return m

    def body_id(self, name):
        return self.body_names.index(name)


# --------------------------------------------------------------------------------------------
# geometry -> inertia (MuJoCo `mjCGeom::SetInertia`)
# --------------------------------------------------------------------------------------------
def geom_volume(gtype, size):
    if gtype == GEOM_SPHERE:
        return 4.0 / 3.0 * np.pi * size[0] ** 3
    if gtype == GEOM_CAPSULE:
        r, hl = size[0], size[1]
        return np.pi * r * r * (2 * hl) + 4.0 / 3.0 * np.pi * r ** 3
    if gtype == GEOM_BOX:
        return 8.0 * size[0] * size[1] * size[2]
    return 0.0


def geom_inertia_diag(gtype, size, mass):
    """Principal inertia of a primitive about its own centre, in the geom frame."""
    if gtype == GEOM_SPHERE:
        i = 0.4 * mass * size[0] ** 2
        return np.array([i, i, i])
    if gtype == GEOM_BOX:
        a, b, c = size
        return mass / 3.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
    if gtype == GEOM_CAPSULE:
        r, hl = size[0], size[1]
        L = 2.0 * hl
        m_s = mass * 4.0 * r / (4.0 * r + 6.0 * hl)  # both hemispherical caps = one sphere
        m_c = mass - m_s
        ixy = m_c * (3 * r * r + L * L) / 12.0 + 0.4 * m_s * r * r + m_s * L * (3 * r + 2 * L) / 8.0
        iz = 0.5 * m_c * r * r + 0.4 * m_s * r * r
        return np.array([ixy, ixy, iz])
    return np.zeros(3)


# --------------------------------------------------------------------------------------------
# kinematics + joint-space inertia in float64 (compile-time only)
# --------------------------------------------------------------------------------------------
def kinematics(m: ModelConst, qpos):
    """World poses of all bodies.  Returns xpos, xquat, xmat, xanchor, xaxis."""
    xpos = np.zeros((m.nbody, 3))
    xquat = np.zeros((m.nbody, 4))
    xquat[0, 0] = 1.0
    xanchor = np.zeros((m.njnt, 3))
    xaxis = np.zeros((m.njnt, 3))
    for b in range(1, m.nbody):
        p = m.body_parent[b]
        pos = xpos[p] + quat_to_mat(xquat[p]) @ m.body_pos[b]
        quat = xquat[p].copy()
        for j in range(m.body_jntadr[b], m.body_jntadr[b] + m.body_jntnum[b]):
            qa = m.jnt_qposadr[j]
            if m.jnt_type[j] == JNT_FREE:
                pos = np.asarray(qpos[qa:qa + 3], dtype=np.float64).copy()
                quat = np.asarray(qpos[qa + 3:qa + 7], dtype=np.float64)
                quat = quat / np.linalg.norm(quat)
                xanchor[j] = pos
                xaxis[j] = [0, 0, 1]
            else:
                R = quat_to_mat(quat)
                xanchor[j] = pos + R @ m.jnt_pos[j]
                xaxis[j] = R @ m.jnt_axis[j]
                quat = quat_mul(quat, axis_angle_quat(m.jnt_axis[j], qpos[qa] - m.qpos0[qa]))
                pos = xanchor[j] - quat_to_mat(quat) @ m.jnt_pos[j]
        xpos[b] = pos
        xquat[b] = quat / np.linalg.norm(quat)
    xmat = np.stack([quat_to_mat(q) for q in xquat])
    return xpos, xquat, xmat, xanchor, xaxis


def dof_jacobians(m: ModelConst, qpos, point_of_body=None):
    """Per-body 6 x nv Jacobians [rot; trans] of the body inertial-frame origin (xipos)."""
    xpos, xquat, xmat, xanchor, xaxis = kinematics(m, qpos)
    xipos = xpos + np.einsum("bij,bj->bi", xmat, m.body_ipos)
    J = np.zeros((m.nbody, 6, m.nv))
    for b in range(1, m.nbody):
        a = b
        while a > 0:
            for j in range(m.body_jntadr[a], m.body_jntadr[a] + m.body_jntnum[a]):
                d = m.jnt_dofadr[j]
                if m.jnt_type[j] == JNT_FREE:
                    J[b, 3:6, d:d + 3] = np.eye(3)
                    for k in range(3):
                        ax = xmat[a][:, k]
                        J[b, 0:3, d + 3 + k] = ax
                        J[b, 3:6, d + 3 + k] = np.cross(ax, xipos[b] - xpos[a])
                else:
                    J[b, 0:3, d] = xaxis[j]
                    J[b, 3:6, d] = np.cross(xaxis[j], xipos[b] - xanchor[j])
            a = m.body_parent[a]
    return J, xpos, xquat, xmat, xipos


def mass_matrix(m: ModelConst, qpos):
    """Dense joint-space inertia  M = diag(armature) + sum_b J_b^T I6_b J_b  (SURVEY.md App. E.2)."""
    J, xpos, xquat, xmat, xipos = dof_jacobians(m, qpos)
    M = np.diag(m.dof_armature.astype(np.float64))
    for b in range(1, m.nbody):
        Iw = xmat[b] @ m.body_inertia[b] @ xmat[b].T
        Jr, Jt = J[b, 0:3], J[b, 3:6]
        M += Jr.T @ Iw @ Jr + m.body_mass[b] * (Jt.T @ Jt)
    return M


# --------------------------------------------------------------------------------------------
# XML -> ModelConst
# --------------------------------------------------------------------------------------------
def compile_mjcf(path) -> ModelConst:
    path = Path(path)
    root = ET.parse(path).getroot()
    m = ModelConst(name=path.stem)

    comp = root.find("compiler")
    if comp is not None:
        if comp.get("angle", "degree") != "radian":
            raise NotImplementedError("only angle=radian")
    opt = root.find("option")
    if opt is not None:
        if opt.get("gravity") is not None:
            m.gravity = _vec(opt.get("gravity"), 3)
        if opt.get("timestep") is not None:
            m.timestep = float(opt.get("timestep"))
        if opt.get("integrator", "Euler") != "Euler":
            raise NotImplementedError("only the Euler integrator (MJX default) is restated")
    dflt = root.find("default")
    d_joint = dict(dflt.find("joint").attrib) if dflt is not None and dflt.find("joint") is not None else {}
    d_geom = dict(dflt.find("geom").attrib) if dflt is not None and dflt.find("geom") is not None else {}
    d_motor = dict(dflt.find("motor").attrib) if dflt is not None and dflt.find("motor") is not None else {}

    bodies = [dict(name="world", parent=0, pos=np.zeros(3))]
    joints, geoms = [], []

    def add_geom(g, bid):
        a = {**d_geom, **g.attrib}
        gt = _GEOM_TYPES[a.get("type", "sphere")]
        size = np.zeros(3)
        sz = _vec(a.get("size"))
        pos = _vec(a.get("pos"), 3, [0, 0, 0])
        quat = _vec(a.get("quat"), 4, [1, 0, 0, 0])
        if a.get("fromto") is not None:
            ft = _vec(a["fromto"], 6)
            vec = ft[0:3] - ft[3:6]
            pos = 0.5 * (ft[0:3] + ft[3:6])
            quat = z_to_quat(vec)
            size[0] = sz[0]
            size[1] = 0.5 * np.linalg.norm(vec)
        elif sz is not None:
            size[:sz.size] = sz
        mass = float(a["mass"]) if "mass" in a else float(a.get("density", 1000.0)) * geom_volume(gt, size)
        geoms.append(dict(
            name=a.get("name", f"geom{len(geoms)}"), type=gt, body=bid, size=size, pos=pos, quat=quat,
            mass=mass if gt != GEOM_PLANE else 0.0,
            friction=_vec(a.get("friction"), None, [1, 0.005, 0.0001]),
            margin=float(a.get("margin", 0.0)), gap=float(a.get("gap", 0.0)),
            contype=int(a.get("contype", 1)), conaffinity=int(a.get("conaffinity", 1)),
            condim=int(a.get("condim", 3))))

    def walk(elem, parent_id):
        for g in elem.findall("geom"):
            add_geom(g, parent_id)
        for b in elem.findall("body"):
            bid = len(bodies)
            bodies.append(dict(name=b.get("name"), parent=parent_id, pos=_vec(b.get("pos"), 3, [0, 0, 0])))
            if b.get("quat") is not None or b.get("euler") is not None:
                raise NotImplementedError("body orientation offsets are not used by the humanoid files")
            for j in b.findall("joint"):
                a = {**d_joint, **j.attrib}
                jt = a.get("type", "hinge")
                if jt == "free":
                    # MuJoCo: free joints take no defaults for armature/damping/limits unless set explicitly
                    a = {**j.attrib}
                joints.append(dict(
                    name=a.get("name"), type=JNT_FREE if jt == "free" else JNT_HINGE, body=bid,
                    axis=_vec(a.get("axis"), 3, [0, 0, 1]), pos=_vec(a.get("pos"), 3, [0, 0, 0]),
                    limited=(a.get("limited", "false") == "true") and jt != "free",
                    range=_vec(a.get("range"), 2, [0, 0]), margin=float(a.get("margin", 0.0)),
                    armature=float(a.get("armature", 0.0)), damping=float(a.get("damping", 0.0))))
                if jt not in ("free", "hinge"):
                    raise NotImplementedError(jt)
            walk(b, bid)

    walk(root.find("worldbody"), 0)

    nb = m.nbody = len(bodies)
    m.body_names = [b["name"] for b in bodies]
    m.body_parent = np.array([b["parent"] for b in bodies], dtype=np.int32)
    m.body_pos = np.stack([b["pos"] for b in bodies])
    m.body_level = np.zeros(nb, dtype=np.int32)
    for b in range(1, nb):
        m.body_level[b] = m.body_level[m.body_parent[b]] + 1

    # joints / dofs
    m.njnt = len(joints)
    m.jnt_names = [j["name"] for j in joints]
    m.jnt_type = np.array([j["type"] for j in joints], dtype=np.int32)
    m.jnt_bodyid = np.array([j["body"] for j in joints], dtype=np.int32)
    m.jnt_axis = np.stack([j["axis"] / np.linalg.norm(j["axis"]) for j in joints])
    m.jnt_pos = np.stack([j["pos"] for j in joints])
    m.jnt_limited = np.array([j["limited"] for j in joints], dtype=np.int32)
    m.jnt_range = np.stack([j["range"] for j in joints])
    m.jnt_margin = np.array([j["margin"] for j in joints])
    qadr, dadr, qa, da = [], [], 0, 0
    for j in joints:
        qadr.append(qa)
        dadr.append(da)
        qa += 7 if j["type"] == JNT_FREE else 1
        da += 6 if j["type"] == JNT_FREE else 1
    m.nq, m.nv = qa, da
    m.jnt_qposadr = np.array(qadr, dtype=np.int32)
    m.jnt_dofadr = np.array(dadr, dtype=np.int32)
    m.body_jntadr = np.full(nb, -1, dtype=np.int32)
    m.body_jntnum = np.zeros(nb, dtype=np.int32)
    m.body_dofadr = np.full(nb, -1, dtype=np.int32)
    m.body_dofnum = np.zeros(nb, dtype=np.int32)
    for ji, j in enumerate(joints):
        b = j["body"]
        if m.body_jntadr[b] < 0:
            m.body_jntadr[b] = ji
            m.body_dofadr[b] = m.jnt_dofadr[ji]
        m.body_jntnum[b] += 1
        m.body_dofnum[b] += 6 if j["type"] == JNT_FREE else 1
    m.dof_bodyid = np.zeros(m.nv, dtype=np.int32)
    m.dof_jntid = np.zeros(m.nv, dtype=np.int32)
    m.dof_armature = np.zeros(m.nv)
    m.dof_damping = np.zeros(m.nv)
    for ji, j in enumerate(joints):
        n = 6 if j["type"] == JNT_FREE else 1
        for k in range(n):
            d = m.jnt_dofadr[ji] + k
            m.dof_bodyid[d] = j["body"]
            m.dof_jntid[d] = ji
            m.dof_armature[d] = j["armature"]
            m.dof_damping[d] = j["damping"]
    # dof parent: previous dof of the same body, else last dof of the nearest ancestor body with dofs
    m.dof_parentid = np.full(m.nv, -1, dtype=np.int32)
    for d in range(m.nv):
        b = m.dof_bodyid[d]
        if d > m.body_dofadr[b]:
            m.dof_parentid[d] = d - 1
        else:
            a = m.body_parent[b]
            while a > 0 and m.body_dofnum[a] == 0:
                a = m.body_parent[a]
            if a > 0:
                m.dof_parentid[d] = m.body_dofadr[a] + m.body_dofnum[a] - 1
    m.dof_depth = np.zeros(m.nv, dtype=np.int32)
    for d in range(m.nv):
        if m.dof_parentid[d] >= 0:
            m.dof_depth[d] = m.dof_depth[m.dof_parentid[d]] + 1
    m.qpos0 = np.zeros(m.nq)
    for ji, j in enumerate(joints):
        if j["type"] == JNT_FREE:
            b = j["body"]
            m.qpos0[m.jnt_qposadr[ji]:m.jnt_qposadr[ji] + 3] = m.body_pos[b]
            m.qpos0[m.jnt_qposadr[ji] + 3] = 1.0
    # sparse inertia layout
    rowadr, rownnz, rows, cols = [], [], [], []
    for i in range(m.nv):
        rowadr.append(len(rows))
        j = i
        while j >= 0:
            rows.append(i)
            cols.append(j)
            j = m.dof_parentid[j]
        rownnz.append(len(rows) - rowadr[-1])
    m.M_rowadr = np.array(rowadr, dtype=np.int32)
    m.M_rownnz = np.array(rownnz, dtype=np.int32)
    m.M_row = np.array(rows, dtype=np.int32)
    m.M_col = np.array(cols, dtype=np.int32)

    # geoms
    m.ngeom = len(geoms)
    m.geom_names = [g["name"] for g in geoms]
    m.geom_type = np.array([g["type"] for g in geoms], dtype=np.int32)
    m.geom_bodyid = np.array([g["body"] for g in geoms], dtype=np.int32)
    m.geom_size = np.stack([g["size"] for g in geoms])
    m.geom_pos = np.stack([g["pos"] for g in geoms])
    m.geom_quat = np.stack([g["quat"] for g in geoms])
    fr = np.zeros((m.ngeom, 3))
    for i, g in enumerate(geoms):
        f = g["friction"]
        fr[i, :f.size] = f
        if f.size < 3:
            fr[i, f.size:] = [1, 0.005, 0.0001][f.size:]
    m.geom_friction = fr
    m.geom_margin = np.array([g["margin"] for g in geoms])
    m.geom_gap = np.array([g["gap"] for g in geoms])
    m.geom_contype = np.array([g["contype"] for g in geoms], dtype=np.int32)
    m.geom_conaffinity = np.array([g["conaffinity"] for g in geoms], dtype=np.int32)

    # body inertial frames from geoms (inertiafromgeom="true")
    m.body_mass = np.zeros(nb)
    m.body_ipos = np.zeros((nb, 3))
    m.body_inertia = np.zeros((nb, 3, 3))
    for b in range(1, nb):
        gs = [g for g in geoms if g["body"] == b and g["mass"] > 0]
        mass = sum(g["mass"] for g in gs)
        if mass <= 0:
            raise ValueError(f"body {bodies[b]['name']} has no mass")
        com = sum(g["mass"] * g["pos"] for g in gs) / mass
        I = np.zeros((3, 3))
        for g in gs:
            R = quat_to_mat(g["quat"])
            Ig = R @ np.diag(geom_inertia_diag(g["type"], g["size"], g["mass"])) @ R.T
            o = g["pos"] - com
            I += Ig + g["mass"] * (o @ o * np.eye(3) - np.outer(o, o))
        m.body_mass[b], m.body_ipos[b], m.body_inertia[b] = mass, com, I
    m.total_mass = float(m.body_mass.sum())

    # actuators
    acts = []
    act = root.find("actuator")
    if act is not None:
        for a in act.findall("motor"):
            at = {**d_motor, **a.attrib}
            ji = m.jnt_names.index(at["joint"])
            gear = _vec(at.get("gear"), None, [1.0])[0]
            acts.append(dict(name=at.get("name"), dof=int(m.jnt_dofadr[ji]), gear=gear,
                             limited=at.get("ctrllimited", "false") == "true",
                             range=_vec(at.get("ctrlrange"), 2, [0, 0])))
    m.nu = len(acts)
    m.act_names = [a["name"] for a in acts]
    m.act_dofadr = np.array([a["dof"] for a in acts], dtype=np.int32)
    m.act_gear = np.array([a["gear"] for a in acts])
    m.act_ctrllimited = np.array([a["limited"] for a in acts], dtype=np.int32)
    m.act_ctrlrange = np.stack([a["range"] for a in acts]) if acts else np.zeros((0, 2))

    con = root.find("contact")
    if con is not None:
        m.exclude = [[e.get("body1"), e.get("body2")] for e in con.findall("exclude")]
    pairs = collision_pairs(m)
    m.pair_geom1 = np.array([p[0] for p in pairs], dtype=np.int32)
    m.pair_geom2 = np.array([p[1] for p in pairs], dtype=np.int32)

    # constants at qpos0 (MuJoCo mj_setConst)
    M0 = mass_matrix(m, m.qpos0)
    Minv = np.linalg.inv(M0)
    m.meaninertia = float(np.mean(np.diag(M0)))
    diw = np.diag(Minv).copy()
    for ji, j in enumerate(joints):
        if j["type"] == JNT_FREE:
            d = m.jnt_dofadr[ji]
            diw[d:d + 3] = diw[d:d + 3].mean()
            diw[d + 3:d + 6] = diw[d + 3:d + 6].mean()
    m.dof_invweight0 = diw
    J, *_ = dof_jacobians(m, m.qpos0)
    m.body_invweight0 = np.zeros((nb, 2))
    for b in range(1, nb):
        A = J[b] @ Minv @ J[b].T  # rows 0:3 rotational, 3:6 translational
        m.body_invweight0[b, 0] = np.trace(A[3:6, 3:6]) / 3.0
        m.body_invweight0[b, 1] = np.trace(A[0:3, 0:3]) / 3.0
    return m


def collision_pairs(m: ModelConst):
    """Static collision candidates, SURVEY.md App. A.7 / B.4 (MuJoCo `mj_collision` filters with the MJX rule that
    every candidate is always evaluated): geoms of different bodies whose contype/conaffinity bits are compatible,
    not parent and child (a parent that is the world is exempt), not listed under <contact><exclude>.
    Returns [(g1, g2)] with type(g1) <= type(g2), sorted by (type1, type2, g1, g2)."""
    excl = {frozenset((m.body_id(a), m.body_id(b))) for a, b in m.exclude}
    out = []
    for g1 in range(m.ngeom):
        for g2 in range(g1 + 1, m.ngeom):
            b1, b2 = int(m.geom_bodyid[g1]), int(m.geom_bodyid[g2])
            if b1 == b2:
                continue
            if not ((m.geom_contype[g1] & m.geom_conaffinity[g2]) or (m.geom_contype[g2] & m.geom_conaffinity[g1])):
                continue
            if (b1 != 0 and m.body_parent[b2] == b1) or (b2 != 0 and m.body_parent[b1] == b2):
                continue
            if frozenset((b1, b2)) in excl:
                continue
            a, b = (g1, g2) if m.geom_type[g1] <= m.geom_type[g2] else (g2, g1)
            out.append((a, b))
    out.sort(key=lambda p: (int(m.geom_type[p[0]]), int(m.geom_type[p[1]]), p[0], p[1]))
    return out


def pair_census(m: ModelConst):
    """{'SC': 32, ...}: candidate count per geom-type pair (P plane, S sphere, C capsule, B box)."""
    code = {GEOM_PLANE: "P", GEOM_SPHERE: "S", GEOM_CAPSULE: "C", GEOM_BOX: "B"}
    out = {}
    for a, b in zip(m.pair_geom1, m.pair_geom2):
        k = code[int(m.geom_type[a])] + code[int(m.geom_type[b])]
        out[k] = out.get(k, 0) + 1
    return out


def load_model(path) -> ModelConst:
    """Load an MJCF file (compiled on the fly) or a compiled `*.model.json` asset."""
    path = Path(path)
    if path.suffix == ".json":
        return ModelConst.from_json(path.read_text())
    return compile_mjcf(path)
