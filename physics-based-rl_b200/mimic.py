"""Mimic-mode configuration: builds the C-ABI `PbrlMimicDesc` (include/pbrl.h) from a MotionTable and a model.

Build-defined part of the env (the reference's reward is zeros(3)): DeepMimic weights
0.65 / 0.1 / 0.15 / 0.1 with scales 2 / 0.1 / 40 / 10, end effectors = wrists and feet, fall = any geom other
than the feet touching the floor.
"""
import ctypes

import numpy as np

from ._lib import MimicDesc
from .mocap import FRAME_STRIDE, MotionTable

END_EFFECTORS = [("right_elbow", "right_wrist"), ("left_elbow", "left_wrist"), ("right_ankle", "right_ankle"),
                 ("left_ankle", "left_ankle")]  # (body, geom whose centre is the end-effector point)
FOOT_GEOMS = ("right_ankle", "left_ankle")


def make_mimic_desc(model, table: MotionTable, clips=None, w=(0.65, 0.1, 0.15, 0.1), k=(2.0, 0.1, 40.0, 10.0)):
    """Returns (MimicDesc, keepalive).  `clips`: names to include (default: all clips of the table, table order)."""
    names = list(table.names) if clips is None else list(clips)
    if len(names) > 8:
        raise ValueError("at most 8 clips")
    if table.frames.shape[1] != FRAME_STRIDE:
        raise ValueError("motion table stride mismatch")
    frames = np.ascontiguousarray(table.frames, dtype=np.float32)
    d = MimicDesc()
    d.nclip = len(names)
    d.nframes_total = frames.shape[0]
    d.frames = frames.ctypes.data
    for c, n in enumerate(names):
        h = table.header[table.names.index(n)]
        d.clip_first[c], d.clip_nframes[c] = int(h[0]), int(h[1])
        d.clip_duration[c], d.clip_wrap[c], d.clip_dt[c], d.clip_fall_enabled[c] = float(h[2]), int(h[3]), float(h[4]), int(h[5])
    d.w_pose, d.w_vel, d.w_ee, d.w_com = w
    d.k_pose, d.k_vel, d.k_ee, d.k_com = k
    d.n_ee = len(END_EFFECTORS)
    for e, (body, geom) in enumerate(END_EFFECTORS):
        d.ee_body[e] = model.body_names.index(body)
        off = model.geom_pos[model.geom_names.index(geom)]
        for c in range(3):
            d.ee_offset[3 * e + c] = float(off[c])
    mask = 0
    for g, name in enumerate(model.geom_names):
        if model.geom_type[g] != 0 and name not in FOOT_GEOMS:
            mask |= 1 << g
    d.fall_geom_mask = mask
    return d, frames
