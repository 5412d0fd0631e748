"""dev study: what the lock-step grouping of 16 envs costs, from the trace of tools/dev/solver_trace.py -- the group maximum of the
solver work per substep under the identity order, the balanced order of the previous env step (what k_step uses), a per-substep
regrouping inside pools, and the perfect order; plus the imbalance over CTAs of a static round-robin against dynamic hand-out."""
import numpy as np
tr = np.load('/tmp/trace.npy')  # steps, N, 8, 2
S, N, F, _ = tr.shape
tr = tr[5:]  # after warm-up
S = tr.shape[0]
nit = tr[..., 0]; nd = tr[..., 1]
DW = 3
work = nit + (DW-1)*nd   # round cost in sparse-iteration units (dense iteration = 3)
print("mean niter", nit.mean(), "dense it", nd.mean(), "hist", np.bincount(nit.ravel())/nit.size)
G = 16
def rounds_cost(order_fn, label):
    # cost model: per substep per group: sum over rounds r of max over members participating of cost(round) ; simplified: group cost = max over members of work (approx) 
    tot = 0; cnt = 0
    for s in range(S):
        for f in range(F):
            order = order_fn(s, f)
            w = work[s, order, f].reshape(-1, G)
            tot += w.max(1).sum(); cnt += w.shape[0]
    print(f"{label:50s} mean group max-work per substep {tot/cnt:.3f}")
    return tot/cnt
ident = np.arange(N)
rounds_cost(lambda s,f: ident, "identity")
# previous env step's max work (current balance order, cost_mode 0)
def prev_step_max(s, f):
    if s == 0: return ident
    return np.argsort(-work[s-1].max(1), kind='stable')
rounds_cost(prev_step_max, "prev env-step max (current)")
def prev_substep(s, f):
    if f == 0:
        if s == 0: return ident
        key = work[s-1, :, F-1]
    else: key = work[s, :, f-1]
    return np.argsort(-key, kind='stable')
rounds_cost(prev_substep, "prev substep work, global pool")
def prev_substep_pool(P):
    def fn(s, f):
        if f == 0:
            if s == 0: return ident
            key = work[s-1, :, F-1]
        else: key = work[s, :, f-1]
        order = ident.copy()
        for p0 in range(0, N, P):
            seg = order[p0:p0+P]
            order[p0:p0+P] = seg[np.argsort(-key[seg], kind='stable')]
        return order
    return fn
for P in (112, 224, 512):
    rounds_cost(prev_substep_pool(P), f"prev substep work, pool {P}")
rounds_cost(lambda s,f: np.argsort(-work[s,:,f], kind='stable'), "oracle (perfect)")
print("mean work", work.mean())
print("---- with dense weight w")
for DWx in (1, 1.5, 2, 3):
    work = nit + (DWx-1)*nd
    def prev_step_max(s, f):
        if s == 0: return ident
        return np.argsort(-work[s-1].max(1), kind='stable')
    print("DW", DWx); rounds_cost(prev_step_max, "prev env-step max (current)")
    rounds_cost(prev_substep_pool(112), "prev substep pool 112")
# fraction of envs with any dense iteration in a substep, and how persistent
hasd = nd > 0
print("frac substeps with dense", hasd.mean(), "P(dense | dense prev substep)", (hasd[:,:,1:] & hasd[:,:,:-1]).sum()/hasd[:,:,:-1].sum())
print("frac of envs with any dense in env step", hasd.any(2).mean())

tr = np.load('/tmp/trace.npy'); S,N,F,_ = tr.shape
nit = tr[...,0]; nd = tr[...,1]; work = nit + 2*nd
G=16; U=0.091
rng = np.random.default_rng(0)
res=[]
for s in range(6, S):
    key = work[s-1].max(1)
    # tile to 16384 envs: bootstrap envs
    idx = rng.integers(0, N, 16384)
    k = key[idx]; order = np.argsort(-k, kind='stable')
    w = work[s][idx][order].reshape(-1, G, F)          # groups x 16 x 8
    gcost = (1 + U*w.max(1)).sum(1)                     # per group: sum over substeps
    ng = len(gcost); P=148
    # static round robin
    rr = np.zeros(P); 
    for g in range(ng): rr[g % P] += gcost[g]
    # snake
    sn = np.zeros(P)
    for g in range(ng):
        p, k2 = divmod(g, P); sn[k2 if p % 2 == 0 else P-1-k2] += gcost[g]
    # dynamic (LPT in list order)
    dy = np.zeros(P)
    for g in range(ng): dy[np.argmin(dy)] += gcost[g]
    res.append((rr.max()/rr.mean(), sn.max()/sn.mean(), dy.max()/dy.mean(), gcost.max()/gcost.mean(), gcost.min()/gcost.mean()))
print("max/mean: round-robin %.3f snake %.3f dynamic %.3f ; group cost max/mean %.2f min/mean %.2f" % tuple(np.mean(res,0)))
