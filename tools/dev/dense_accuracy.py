"""dev: accuracy of the body-vs-body (dense Hessian) solve: GPU, host emulation and the fp32 oracle against the f64 oracle,
binned by the deepest body-vs-body penetration of the state."""
import sys; sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
import numpy as np
from helpers import *
import parity_cases as pc
m = load_model_asset("final_humanoid")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
use_gpu = len(sys.argv) > 2 and sys.argv[2] == "gpu"
rng = np.random.default_rng(5)
qpos, qvel = random_states(m, N, 5, gravity_heights=True, vel_scale=1.0, joint_frac=0.6)
ctrl, warm = rng.uniform(-1, 1, (N, 28)), rng.normal(size=(N, 34))
q32, v32, c32, w32 = [np.asarray(x, np.float32) for x in (qpos, qvel, ctrl, warm)]
outs = {"emu": EmuBackend(m).debug_forward(q32, v32, c32, w32)}
if use_gpu:
    outs["gpu"] = GpuBackend(m).debug_forward(q32, v32, c32, w32)
orc, orc32 = pc._oracle(m), Oracle(m, "f32")
get = lambda out, i, k: out[i, DBG[k][0]:DBG[k][0] + DBG[k][1]]
pen, errs = [], {k: [] for k in list(outs) + ["orc32"]}
for i in range(N):
    d = orc.new_data()
    d.qpos[:35], d.qvel[:34], d.ctrl[:28], d.qacc_warmstart[:34] = q32[i], v32[i], c32[i], w32[i]
    orc.forward(d)
    nc, ns = int(d.ncon[0]), int(d.nself[0])
    if ns == 0 or ns > 8:
        continue
    pen.append(-min(d.con_dist[nc - ns:nc]))
    d2 = orc32.new_data()
    d2.qpos[:35], d2.qvel[:34], d2.ctrl[:28], d2.qacc_warmstart[:34] = q32[i], v32[i], c32[i], w32[i]
    orc32.forward(d2)
    errs["orc32"].append(relerr(d2.qacc[:34], d.qacc[:34]))
    for k, o in outs.items():
        errs[k].append(relerr(get(o, i, "qacc"), d.qacc[:34]))
pen = np.array(pen)
edges = [0, 0.002, 0.01, 0.03, 0.06, 1.0]
print("envs with body-vs-body contacts:", len(pen))
for a, b in zip(edges[:-1], edges[1:]):
    sel = (pen >= a) & (pen < b)
    if sel.sum() == 0: continue
    line = f"pen [{a*1e3:5.0f},{b*1e3:5.0f}) mm  n={sel.sum():4d}"
    for k, v in errs.items():
        v = np.array(v)[sel]
        line += f" | {k}: med {np.median(v):.1e} p95 {np.quantile(v, .95):.1e} max {v.max():.1e}"
    print(line)
