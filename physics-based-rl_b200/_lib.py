"""ctypes binding of libpbrl.so (include/pbrl.h).  There is no fallback: if the CUDA library is missing
or no B200 is visible, every call raises."""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

from ._desc import PbrlBuffers, build_model_desc

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["PBRL_LIB"]) if os.environ.get("PBRL_LIB") else PKG / "libpbrl.so"  # PBRL_LIB: dev builds
CSRC = PKG / "csrc"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


class PbrlError(RuntimeError):
    pass


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile csrc/pbrl_lib.cu for sm_100a into physics-based-rl_b200/libpbrl.so (in-tree)."""
    if out is not None:
        cmd = ["nvcc", *NVCC_FLAGS, *extra_flags, "-o", str(out), str(CSRC / "pbrl_lib.cu")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise PbrlError("nvcc failed:\n" + r.stderr)
        return Path(out)
    srcs = [CSRC / "pbrl_lib.cu", CSRC / "pbrl_core.cuh", CSRC / "pbrl_model.h", CSRC / "pbrl_desc_reflect.h",
            CSRC / "pbrl_mimic.cuh", PKG.parent / "include" / "pbrl.h"]
    srcs = [s for s in srcs if s.exists()]
    if not force and LIB_PATH.exists() and LIB_PATH.stat().st_mtime >= max(s.stat().st_mtime for s in srcs):
        return LIB_PATH
    cmd = ["nvcc", *NVCC_FLAGS, "-o", str(LIB_PATH), str(CSRC / "pbrl_lib.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        print(" ".join(cmd), r.stdout, r.stderr, sep="\n")
    if r.returncode:
        raise PbrlError("nvcc failed:\n" + r.stderr)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise PbrlError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a). No CPU fallback exists.")
        L = ctypes.CDLL(str(LIB_PATH))
        L.pbrl_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise PbrlError(lib().pbrl_last_error().decode())


EXPORTS = ["pbrl_model_desc_size", "pbrl_model_desc_field", "pbrl_last_error", "pbrl_device_count",
           "pbrl_model_create", "pbrl_model_destroy", "pbrl_reset", "pbrl_balance_order", "pbrl_autoreset", "pbrl_step", "pbrl_step_launches", "pbrl_replay", "pbrl_replay_rollout", "pbrl_pd_torque",
           "pbrl_debug_forward", "pbrl_debug_stage_profile", "pbrl_batch_create", "pbrl_batch_destroy", "pbrl_batch_reset_host",
           "pbrl_batch_step_host", "pbrl_batch_get_state_host"]


class MimicDesc(ctypes.Structure):
    """Mirror of PbrlMimicDesc (include/pbrl.h)."""
    _fields_ = [("nclip", ctypes.c_int), ("nframes_total", ctypes.c_int), ("frames", ctypes.c_void_p),
                ("clip_first", ctypes.c_int * 8), ("clip_nframes", ctypes.c_int * 8),
                ("clip_duration", ctypes.c_float * 8), ("clip_dt", ctypes.c_float * 8),
                ("clip_wrap", ctypes.c_int * 8), ("clip_fall_enabled", ctypes.c_int * 8),
                ("w_pose", ctypes.c_float), ("w_vel", ctypes.c_float), ("w_ee", ctypes.c_float), ("w_com", ctypes.c_float),
                ("k_pose", ctypes.c_float), ("k_vel", ctypes.c_float), ("k_ee", ctypes.c_float), ("k_com", ctypes.c_float),
                ("n_ee", ctypes.c_int), ("ee_body", ctypes.c_int * 4), ("ee_offset", ctypes.c_float * 12),
                ("fall_geom_mask", ctypes.c_uint)]


class Model:
    """Device-resident model handle (PbrlModel*)."""

    def __init__(self, model_const, device=0, mimic=None):
        L = lib()
        if L.pbrl_device_count() <= 0:
            raise PbrlError("no CUDA device visible: the simulator has no CPU path")
        self.const = model_const
        self.device = device
        desc = build_model_desc(model_const, L.pbrl_model_desc_size, L.pbrl_model_desc_field)
        self._mimic_keep = mimic
        h = ctypes.c_void_p()
        check(L.pbrl_model_create(desc.ctypes.data_as(ctypes.c_void_p),
                                  ctypes.byref(mimic) if mimic is not None else None, ctypes.c_int(device),
                                  ctypes.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib().pbrl_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def make_buffers(**tensors):
    b = PbrlBuffers()
    for name, _ in PbrlBuffers._fields_:
        t = tensors.get(name)
        setattr(b, name, None if t is None else t.data_ptr())
    return b


def current_stream_ptr(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
