"""dev study (CPU, host emulation built with -DPBRL_EMU_TRACE): Newton iterations and dense-Hessian iterations per (env, substep) on the
bench workload, and how often the convex narrow phase runs.  Output: /tmp/trace.npy for tools/dev/grouping_analysis.py.
usage: python tools/dev/solver_trace.py [envs]"""
import sys, ctypes, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import bench
from helpers import *
import helpers
model, clips, mim, n_env, pd_mode, descr = bench.load_assets("mimic_mixed")
table, mdesc, keep = mim
N = int(sys.argv[1]) if len(sys.argv)>1 else 1024
STEPS = 25
motion_id, ref_t0, actions = bench.synth_inputs(model, clips, mim, N, 0, STEPS, "random")
import subprocess
subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DPBRL_EMU_TRACE", "-o", "/tmp/libpbrl_emu_trace.so", "/root/repo/tests/emu/emu.cpp"], check=True)
helpers.build_emu = lambda: "/tmp/libpbrl_emu_trace.so"
b = EmuBackend(model); b.set_mimic(mdesc)
s = b.new_state(N); s.motion_id[:] = motion_id; s.ref_t0[:] = ref_t0
b.reset(s, reward_mode=1)
kp, kd = util_data.generate_kp_kd_gains()
trace = np.zeros((STEPS, N, 8, 2), np.int32)
t=time.time()
for k in range(STEPS):
    b.lib.emu_set_trace(trace[k].ctypes.data_as(ctypes.c_void_p))
    b.step(s, actions[k % len(actions)], kp, kd, pd_mode=3, pd_dt=0.002, n_frames=8, reward_mode=1)
print("time", time.time()-t)
np.save('/tmp/trace.npy', trace)
b.lib.emu_counters.restype = ctypes.POINTER(ctypes.c_long)
c = b.lib.emu_counters()
print("substeps", c[0], "need", c[1]/c[0], "need cb", c[2]/c[0], "need bb", c[3]/c[0], "pairs/need", c[4]/max(1,c[1]), "act", c[5]/c[0], "contacts/act", c[6]/max(1,c[5]), "act bb", c[7]/c[0])
