"""dev: debug_forward on the GPU vs the host emulation vs the f64 oracle on the states of check_solver_rule (N=128, seed=72)."""
import sys; sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
import numpy as np, copy
from helpers import *
import parity_cases as pc
m = load_model_asset("final_humanoid")
N, seed = 128, 72
rng = np.random.default_rng(seed)
qpos, qvel = random_states(m, N, seed, gravity_heights=True, vel_scale=1.0, joint_frac=0.6)
ctrl, warm = rng.uniform(-1, 1, (N, 28)), rng.normal(size=(N, 34))
q32, v32, c32, w32 = [np.asarray(x, np.float32) for x in (qpos, qvel, ctrl, warm)]
g = GpuBackend(m).debug_forward(q32, v32, c32, w32)
e = EmuBackend(m).debug_forward(q32, v32, c32, w32)
orc = pc._oracle(m)
get = lambda out, i, k: out[i, DBG[k][0]:DBG[k][0] + DBG[k][1]]
rows = []
for i in range(N):
    d = orc.new_data()
    d.qpos[:35], d.qvel[:34], d.ctrl[:28], d.qacc_warmstart[:34] = q32[i], v32[i], c32[i], w32[i]
    orc.forward(d)
    rows.append((relerr(get(g, i, "qacc"), d.qacc[:34]), relerr(get(e, i, "qacc"), d.qacc[:34]), relerr(get(g, i, "qacc"), get(e, i, "qacc")), i))
rows.sort(reverse=True)
for r in rows[:6]:
    i = r[3]
    print("env", i, "gpu-vs-f64 %.2e emu-vs-f64 %.2e gpu-vs-emu %.2e" % r[:3], "nself", get(g, i, "nself"), get(e, i, "nself"), "niter", get(g, i, "niter"), get(e, i, "niter"))
    print("   self_dist gpu", get(g, i, "self_dist")); print("   self_dist emu", get(e, i, "self_dist"))
    print("   self_force gpu", get(g, i, "self_force").reshape(8, 4).sum(1)); print("   self_force emu", get(e, i, "self_force").reshape(8, 4).sum(1))
    for k in ("xpos", "qM", "bias", "qacc_smooth", "qfrc_constraint", "con_dist"):
        print("   ", k, relerr(get(g, i, k), get(e, i, k)))
