"""Skeleton tables, PD gains and index helpers.

Same names and values as /root/reference/utils/util_data.py:4-87 (they are the domain's data:
DeepMimic joint order, dof counts, kp/kd per joint), returned as numpy instead of jax arrays.
The index helpers take a `ModelConst` (mjcf.py) where the reference takes a `mujoco.MjModel`.
"""
import numpy as np

BODIES = ["world", "root", "chest", "neck", "right_shoulder", "right_elbow", "left_shoulder",
          "left_elbow", "right_hip", "right_knee", "right_ankle", "left_hip", "left_knee", "left_ankle"]

# joint order inside a DeepMimic frame
BODY_JOINTS_IN_DP_ORDER = ["chest", "neck", "right_hip", "right_knee", "right_ankle", "right_shoulder",
                           "right_elbow", "left_hip", "left_knee", "left_ankle", "left_shoulder", "left_elbow"]

# joint order inside qpos (MJCF order); elbows and knees are 1-dof
BODY_JOINTS = ["chest", "neck", "right_shoulder", "right_elbow", "left_shoulder", "left_elbow",
               "right_hip", "right_knee", "right_ankle", "left_hip", "left_knee", "left_ankle"]

DOF_DEF = {"root": 3, "chest": 3, "neck": 3, "right_shoulder": 3, "right_elbow": 1, "right_wrist": 0,
           "left_shoulder": 3, "left_elbow": 1, "left_wrist": 0, "right_hip": 3, "right_knee": 1,
           "right_ankle": 3, "left_hip": 3, "left_knee": 1, "left_ankle": 3}

PARAMS_KP_KD = {"chest": [1000, 100], "neck": [100, 10], "right_shoulder": [400, 40], "right_elbow": [300, 30],
                "left_shoulder": [400, 40], "left_elbow": [300, 30], "right_hip": [500, 50],
                "right_knee": [500, 50], "right_ankle": [400, 40], "left_hip": [500, 50],
                "left_knee": [500, 50], "left_ankle": [400, 40]}

BODY_DEFS = ["root", "chest", "neck", "right_hip", "right_knee", "right_ankle", "right_shoulder",
             "right_elbow", "right_wrist", "left_hip", "left_knee", "left_ankle", "left_shoulder",
             "left_elbow", "left_wrist"]

_AXIS_OFFSET = {"X": 0, "Y": 1, "Z": 2}


def _jntadr(model, name_body):
    return int(model.body_jntadr[model.body_names.index(name_body)])


def get_joint_index(model, name_body, axis):
    """qpos index of a hinge: jntadr + 6 (util_data.py:48-56)."""
    return _jntadr(model, name_body) + _AXIS_OFFSET.get(axis, 0) + 6


def get_vel_indx(model, name_body, axis):
    """qvel index of a hinge: jntadr + 5 (util_data.py:61-67)."""
    return _jntadr(model, name_body) + _AXIS_OFFSET.get(axis, 0) + 5


def get_actuator_indx(model, name, axis):
    """actuator index of a hinge: jntadr - 1 (util_data.py:80-87)."""
    return _jntadr(model, name) + _AXIS_OFFSET.get(axis, 0) - 1


def generate_kp_kd_gains():
    """28 kp and 28 kd in actuator order (util_data.py:70-76)."""
    kp, kd = [], []
    for joint in BODY_JOINTS:
        kp += [PARAMS_KP_KD[joint][0]] * DOF_DEF[joint]
        kd += [PARAMS_KP_KD[joint][1]] * DOF_DEF[joint]
    return np.array(kp, dtype=np.float32), np.array(kd, dtype=np.float32)
