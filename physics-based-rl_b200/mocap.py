"""DeepMimic mocap txt -> MuJoCo-layout reference trajectories (host side, setup time, float64).

Mirror of the reference's converter interface (/root/reference/utils/SimpleConverter.py:15-238,
helpers /root/reference/some_math/math_utils.py:15-65 and
/root/reference/some_math/transformations.py:1031-1097,1174-1193), restated with plain numpy:
pyquaternion and jax are not needed.  Same attribute names (`data`, `data_pos`, `data_vel`,
`durations`, `total_time`, `duration_dict`) and the same quirks:

  * Y-up left-handed -> Z-up right-handed:  pos (x,y,z)->(x,-z,y), quat (w,x,y,z)->(w,x,-z,y);
  * 3-dof joints become intrinsic X-Y-Z Euler angles ('rxyz');
  * angular "velocity" = axis*angle/dt of conj(q_cur)*q_prev (i.e. the negative of the usual
    body-frame rate), first frame velocity 0, frame k uses the duration of frame k-1;
  * quaternions are not normalised.

`MotionTable` (bottom) is the build-defined part: the packed per-clip tables the CUDA kernels
interpolate in mimic mode ("parity unpinned").
"""
from __future__ import annotations

import json
import math
from pathlib import Path

import numpy as np

from .util_data import BODY_JOINTS, BODY_JOINTS_IN_DP_ORDER, DOF_DEF

MOTION_NAMES = ["walk", "punch", "kick", "jump", "backflip", "getup_facedown", "getup_faceup"]


# ---- small quaternion helpers (w,x,y,z) ------------------------------------------------------
def _qmul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw])


def align_position(pos):
    """(x, y, z) -> (x, -z, y)   (math_utils.py:15-22)."""
    pos = np.asarray(pos, dtype=np.float64)
    assert pos.shape[0] == 3
    return np.array([pos[0], -pos[2], pos[1]])


def align_rotation(rot):
    """Similarity by Rx(+90deg): (w, x, y, z) -> (w, x, -z, y)   (math_utils.py:25-42)."""
    c = math.sqrt(0.5)
    left = np.array([c, c, 0.0, 0.0])
    right = np.array([c, -c, 0.0, 0.0])
    return _qmul(_qmul(left, np.asarray(rot, dtype=np.float64)), right)


def calc_rot_vel(seg_0, seg_1, dura):
    """axis*angle/dura of conj(seg_0)*seg_1 with pyquaternion's conventions (math_utils.py:46-65)."""
    q0 = np.asarray(seg_0, dtype=np.float64)
    q1 = np.asarray(seg_1, dtype=np.float64)
    qd = _qmul(np.array([q0[0], -q0[1], -q0[2], -q0[3]]), q1)
    n = np.linalg.norm(qd)
    if abs(1.0 - n * n) >= 1e-14 and n > 0:
        qd = qd / n
    vn = np.linalg.norm(qd[1:])
    axis = qd[1:] / vn if vn >= 1e-17 else np.zeros(3)
    angle = 2.0 * math.atan2(vn, qd[0])
    angle = ((angle + math.pi) % (2.0 * math.pi)) - math.pi
    if angle == -math.pi:
        angle = math.pi
    with np.errstate(divide="ignore", invalid="ignore"):
        w = angle / dura * axis
    return [w[0], w[1], w[2]]


def euler_xyz_from_quat_xyzw(q_xyzw):
    """Intrinsic X-Y-Z Euler angles of a quaternion given as (x, y, z, w)  == the reference's
    `euler_from_quaternion(q, axes='rxyz')` (transformations.py:1089, 1031-1086)."""
    x, y, z, w = [float(v) for v in q_xyzw]
    nq = x * x + y * y + z * z + w * w
    if nq < np.finfo(float).eps * 4.0:
        return (0.0, 0.0, 0.0)
    s = 2.0 / nq
    m00 = 1.0 - s * (y * y + z * z)
    m01 = s * (x * y - z * w)
    m02 = s * (x * z + y * w)
    m12 = s * (y * z - x * w)
    m22 = 1.0 - s * (x * x + y * y)
    cy = math.sqrt(m22 * m22 + m12 * m12)
    if cy > np.finfo(float).eps * 4.0:
        ex = -math.atan2(m12, m22)
        ey = math.atan2(m02, cy)
        ez = -math.atan2(m01, m00)
    else:  # gimbal lock: transformations.py falls back to az = 0 (pre-swap)
        m10 = s * (x * y + z * w)
        m11 = 1.0 - s * (x * x + z * z)
        ex = 0.0
        ey = math.atan2(m02, cy)
        ez = math.atan2(m10, m11)
    return (ex, ey, ez)


class SimpleConverter(object):
    """Same public surface as /root/reference/utils/SimpleConverter.py:15."""

    def __init__(self, file) -> None:
        self.pos_dim = 3
        self.rot_dim = 4
        self.file = file
        self.load_data()

    def load_data(self):
        with open(self.file, "r") as fin:
            data = json.load(fin)
        self.loop = data.get("Loop", "none")
        self.motions = np.array(data["Frames"], dtype=np.float64)
        self.data = np.full(self.motions.shape, np.nan)
        self.dt = self.motions[0][0]

    def load_mocap(self):
        self.read_raw_data()
        self.convert_to_mujoco_data()
        self.create_dict_duration()

    def read_raw_data(self):
        self.all_states = []
        self.durations = []
        self.total_time = 0.0
        for frame in self.motions:
            self.calculate_duration(frame[0])
            state = {"root_pos": align_position(frame[1:4]), "root_rot": align_rotation(frame[4:8])}
            self.align_joints(frame, 8, 8, state)
        self.durations = np.array(self.durations)

    def align_joints(self, frame, curr_idx, offset_idx, state):
        """The joint entries of one raw frame, from column `offset_idx` on, in the file's joint order: a hinge is one angle, a
        ball joint a wxyz quaternion whose axes are re-aligned (positions are relative to the parent and need nothing); the
        filled `state` joins `all_states` (SimpleConverter.py:193-216; `curr_idx` is overwritten there too)."""
        for joint in BODY_JOINTS_IN_DP_ORDER:
            curr_idx = offset_idx
            if DOF_DEF[joint] == 1:
                offset_idx += 1
                state[joint] = frame[curr_idx:offset_idx].copy()
            else:
                offset_idx += 4
                state[joint] = align_rotation(frame[curr_idx:offset_idx])
        self.all_states.append(state)

    def calculate_duration(self, frame):
        """`frame` is a frame's duration (its first column): add it to `total_time` and to `durations` (SimpleConverter.py:219-223)."""
        self.total_time += frame
        self.durations.append(frame)

    def convert_to_mujoco_data(self):
        self.data_vel, self.data_pos = [], []
        for k, state in enumerate(self.all_states):
            dura = self.durations[k] if k == 0 else self.durations[k - 1]
            vel, ang = [], []
            self.data[k, 0] = dura
            self.data[k, 1:4] = state["root_pos"]
            self.data[k, 4:8] = state["root_rot"]
            if k == 0:
                vel += [0.0] * 6
            else:
                with np.errstate(divide="ignore", invalid="ignore"):
                    vel += ((self.data[k, 1:4] - self.data[k - 1, 1:4]) * 1.0 / dura).tolist()
                vel += calc_rot_vel(self.data[k, 4:8], self.data[k - 1, 4:8], dura)
            ang += state["root_pos"].tolist() + state["root_rot"].tolist()
            off = 8
            for joint in BODY_JOINTS:
                n = 1 if DOF_DEF[joint] == 1 else 4
                self.data[k, off:off + n] = state[joint]
                if n == 1:
                    if k == 0:
                        vel += [0.0]
                    else:
                        with np.errstate(divide="ignore", invalid="ignore"):
                            vel += ((self.data[k, off:off + 1] - self.data[k - 1, off:off + 1]) / dura).tolist()
                    ang += state[joint].tolist()
                else:
                    if k == 0:
                        vel += [0.0, 0.0, 0.0]
                    else:
                        vel += calc_rot_vel(self.data[k, off:off + 4], self.data[k - 1, off:off + 4], dura)
                    q = state[joint]
                    ang += list(euler_xyz_from_quat_xyzw([q[1], q[2], q[3], q[0]]))
                off += n
            self.data_vel.append(np.array(vel))
            self.data_pos.append(np.array(ang))

    def create_dict_duration(self):
        # the reference accumulates with jp.cumsum, i.e. in float32 (SimpleConverter.py:226-238)
        cum = np.cumsum(self.durations.astype(np.float32), dtype=np.float32)
        cum = np.concatenate([np.zeros(1, dtype=np.float32), cum[:-1]])
        self.duration_dict = {i: [float(cum[i]), float(self.durations[i])] for i in range(len(self.durations))}


# --------------------------------------------------------------------------------------------
# Build-defined: packed motion tables for mimic mode
# --------------------------------------------------------------------------------------------
FRAME_STRIDE = 104  # floats per frame in the device table (416 B, a 16-byte multiple for cp.async.bulk)
# frame layout (include/pbrl.h PBRL_FRAME_STRIDE):
#   [0:35]    qpos: root pos, root quat (normalised), 28 hinge angles (intrinsic XYZ Euler per 3-dof joint)
#   [35:69]   qvel in MuJoCo convention by forward differences (root linear: world, root angular: body-local,
#             hinges: Euler-angle rates)
#   [69:101]  unit quaternions (w,x,y,z) of the eight 3-dof joints in qpos order
#             (chest, neck, right_shoulder, left_shoulder, right_hip, right_ankle, left_hip, left_ankle)
#   [101]     frame start time   [102] frame duration   [103] pad
SPHERICAL_JOINTS = [j for j in BODY_JOINTS if DOF_DEF[j] == 3]
FALL_ENABLED = {"walk": 1, "punch": 1, "kick": 1, "jump": 1, "backflip": 1, "getup_facedown": 0, "getup_faceup": 0}


def _quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


class MotionTable:
    """All clips packed into one float32 table + per-clip header, for the device (mimic mode; build-defined)."""

    def __init__(self, clips):
        """clips: list of (name, SimpleConverter with all_states / data_pos / durations / loop)."""
        self.names = [c[0] for c in clips]
        frames, hdr = [], []
        start = 0
        for name, conv in clips:
            pos = np.asarray(conv.data_pos, dtype=np.float64).copy()
            pos[:, 3:7] /= np.linalg.norm(pos[:, 3:7], axis=1, keepdims=True)
            dur = np.asarray(conv.durations, dtype=np.float64)
            F = pos.shape[0]
            vel = self.reference_qvel(pos, dur)
            t0 = np.concatenate([[0.0], np.cumsum(dur)[:-1]])
            tab = np.zeros((F, FRAME_STRIDE), dtype=np.float32)
            tab[:, 0:35] = pos
            tab[:, 35:69] = vel
            for k, st in enumerate(conv.all_states):
                for j, jn in enumerate(SPHERICAL_JOINTS):
                    q = np.asarray(st[jn], dtype=np.float64)
                    tab[k, 69 + 4 * j:73 + 4 * j] = q / np.linalg.norm(q)
            tab[:, 101] = t0
            tab[:, 102] = dur
            frames.append(tab)
            total = float(dur[:-1].sum())  # the last frame has dt = 0 in every DeepMimic file
            hdr.append((start, F, total, 1.0 if conv.loop == "wrap" else 0.0, float(dur[0]), float(FALL_ENABLED.get(name, 1))))
            start += F
        self.frames = np.concatenate(frames, axis=0)
        # (nclip, 6): first frame, n frames, duration, wrap, frame dt, fall termination enabled
        self.header = np.asarray(hdr, dtype=np.float64)

    @staticmethod
    def reference_qvel(pos, dur):
        """MuJoCo-convention qvel by forward differences; the last frame repeats the previous velocity."""
        F = pos.shape[0]
        vel = np.zeros((F, 34))
        for k in range(F - 1):
            h = dur[k] if dur[k] > 0 else dur[max(k - 1, 0)]
            vel[k, 0:3] = (pos[k + 1, 0:3] - pos[k, 0:3]) / h
            dq = _qmul(_quat_conj(pos[k, 3:7]), pos[k + 1, 3:7])
            if dq[0] < 0:
                dq = -dq
            vn = np.linalg.norm(dq[1:])
            ang = 2.0 * math.atan2(vn, dq[0])
            vel[k, 3:6] = (dq[1:] / vn * ang / h) if vn > 1e-12 else 0.0
            d = pos[k + 1, 7:] - pos[k, 7:]
            d = (d + np.pi) % (2 * np.pi) - np.pi
            vel[k, 6:] = d / h
        if F > 1:
            vel[F - 1] = vel[F - 2]
        return vel

    def save(self, path):
        np.savez_compressed(path, frames=self.frames, header=self.header, names=np.array(self.names))

    @staticmethod
    def load(path):
        z = np.load(path, allow_pickle=False)
        t = MotionTable.__new__(MotionTable)
        t.frames, t.header, t.names = z["frames"], z["header"], [str(s) for s in z["names"]]
        return t

    def clip_index(self, name):
        return self.names.index(name)

    # offsets of the eight 3-dof joints (table order) and of the four 1-dof joints inside qpos[7:35]
    _SPH_OFFSET = (0, 3, 6, 10, 14, 18, 21, 25)
    _HINGE_OFFSET = (9, 13, 17, 24)

    def reference_joint_angles(self, clip_ids, t):
        """Reference joint angles (N, 28) of clip `clip_ids[e]` at clip time `t[e]`, interpolated like the device does
        (pbrl_mimic.cuh: frame_lookup, slerp of the 3-dof joints then intrinsic-XYZ Euler angles, lerp of the 1-dof joints).
        Host-side helper: PD targets that track the clip."""
        clip_ids = np.asarray(clip_ids, dtype=np.int64)
        t = np.asarray(t, dtype=np.float64)
        h = self.header[clip_ids]
        first, nfr, T, wrap, dt = h[:, 0].astype(np.int64), h[:, 1].astype(np.int64), h[:, 2], h[:, 3] > 0, h[:, 4]
        tc = np.where(wrap, t - np.floor(t / T) * T, np.minimum(t, T))
        tc = np.maximum(tc, 0.0)
        fi = np.minimum((tc / dt).astype(np.int64), nfr - 2)
        u = np.clip(tc / dt - fi, 0.0, 1.0)[:, None]
        f0, f1 = self.frames[first + fi].astype(np.float64), self.frames[first + fi + 1].astype(np.float64)
        out = np.zeros((len(t), 28))
        for o in self._HINGE_OFFSET:
            out[:, o] = (1 - u[:, 0]) * f0[:, 7 + o] + u[:, 0] * f1[:, 7 + o]
        for j, o in enumerate(self._SPH_OFFSET):
            a, b = f0[:, 69 + 4 * j:73 + 4 * j], f1[:, 69 + 4 * j:73 + 4 * j]
            d = np.sum(a * b, axis=1, keepdims=True)
            sg = np.where(d < 0, -1.0, 1.0)
            d = np.abs(d)
            th = np.arccos(np.minimum(d, 1.0))
            lin = d > 0.9995
            sth = np.where(lin, 1.0, np.sin(th))
            wa = np.where(lin, 1 - u, np.sin((1 - u) * th) / sth)
            wb = np.where(lin, u, np.sin(u * th) / sth)
            q = wa * a + sg * wb * b
            q /= np.linalg.norm(q, axis=1, keepdims=True)
            w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
            m00, m01, m02 = 1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)
            m12, m22 = 2 * (y * z - x * w), 1 - 2 * (x * x + y * y)
            out[:, o] = np.arctan2(-m12, m22)
            out[:, o + 1] = np.arctan2(m02, np.sqrt(m22 * m22 + m12 * m12))
            out[:, o + 2] = np.arctan2(-m01, m00)
        return out


def load_motion_dir(motion_dir, names=MOTION_NAMES):
    clips = []
    for n in names:
        c = SimpleConverter(str(Path(motion_dir) / f"humanoid3d_{n}.txt"))
        c.load_mocap()
        clips.append((n, c))
    return clips
