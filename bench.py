#!/usr/bin/env python
"""bench.py -- humanoid-mimic env-substeps/s on 1..8 B200 (BASELINE.json metric), plus the CPU reference arm.

  python bench.py --gpus N --steps K --warmup W                 our arm (CUDA, libpbrl through the C ABI)
  python bench.py --impl reference --gpus N --steps K --warmup W  the reference's CPU path (oracle restatement)

One "step" = one env step (PD law + n_frames=8 physics substeps + obs/reward) over the per-GPU batch of
environments.  Prints ONE JSON line on rank 0.  See DESIGN.md section "Measurement" for the definitions.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
PKG = "physics-based-rl_b200"

METRIC = "humanoid-mimic env-substeps/sec"
UNIT = "env-substeps/s"
N_FRAMES = 8
BYTES_PER_ENV_LAUNCH = 1248  # 528 in + 416 state out + 304 epilogue, fp32, unpadded
# pbrl_replay_rollout (k_replay_gather / k_replay_rollout) writes, per (env, frame) item: obs 276 + reward 12 + done 4 + xpos 168 + xquat 224 = 684 B; it reads the clip
# (10 KB, L1/L2-resident), not 276 B per item.  SURVEY 8(d)'s 580 (+364) B figure describes the per-step launch k_replay, which
# does read qpos/qvel rows: 276 in + 304 out + 364 poses = 944 B.
BYTES_PER_REPLAY_ROLLOUT_ITEM = 276 + 12 + 4 + 168 + 224
BYTES_PER_REPLAY_STEP_ENV = 580 + 364
STRONG_TOTAL_ENVS = 65536    # BASELINE configs[4]: 65536 envs sharded at 2/4/8 GPUs (--scaling strong)

WORKLOADS = {
    # name: (model asset, clips, envs per GPU, pd_mode, description)
    "mimic_mixed": ("final_humanoid", ["walk", "punch", "kick", "jump", "backflip", "getup_facedown", "getup_faceup"], 16384, 3,
                    "BASELINE configs[4]/[3]: final_humanoid.xml, contact-rich, reference-state init from all 7 motions "
                    "(env_id mod 7), stable-PD, random target actions, 16384 envs per GPU (65536 at 4 GPUs)"),
    "contact_rich": ("final_humanoid", ["backflip", "jump"], 16384, 3, "BASELINE configs[3]: backflip+jump, 16384 envs"),
    "contact_free": ("final_humanoid_no_gravity", ["punch"], 16384, 3, "BASELINE configs[2]: no-gravity punch, 16384 envs"),
    "cpu64": ("final_humanoid", ["walk"], 64, 3, "BASELINE configs[0]: walk, stable-PD, 64 envs"),
}


def load_assets(workload):
    mjcf = importlib.import_module(PKG + ".mjcf")
    mocap = importlib.import_module(PKG + ".mocap")
    mimic = importlib.import_module(PKG + ".mimic")
    asset, clips, n_env, pd_mode, descr = WORKLOADS[workload]
    model = mjcf.load_model(ROOT / PKG / "assets" / f"{asset}.model.json")
    if os.environ.get("PBRL_BENCH_NO_SELF_COLLISION"):  # ablation: floor contacts only (both arms)
        model.self_collision = 0
    if os.environ.get("PBRL_BENCH_NO_BOX_PAIRS"):  # ablation: drop the 15 capsule-box / box-box candidates (round-1 coverage: 74 of 89)
        keep = [i for i in range(len(model.pair_geom1))
                if not (model.geom_type[model.pair_geom2[i]] == 6 and model.geom_type[model.pair_geom1[i]] in (3, 6))]
        model.pair_geom1, model.pair_geom2 = model.pair_geom1[keep], model.pair_geom2[keep]
        model.n_self_pairs_evaluated = 74
    table = mocap.MotionTable.load(ROOT / PKG / "assets" / "motions.npz")
    mdesc, keep = mimic.make_mimic_desc(model, table, clips)
    return model, clips, (table, mdesc, keep), n_env, pd_mode, descr


def self_collision_label(model):
    if not getattr(model, "self_collision", 1):
        return "off"
    n_eval = int(getattr(model, "n_self_pairs_evaluated", 89))
    return f"{n_eval} of 89 body-vs-body candidates evaluated"


def synth_inputs(model, clips, mim, n_env, first_env, steps, actions_mode, seed=0):
    """Seeded, shard-independent synthetic inputs keyed by the GLOBAL env id: clip = id mod nclip, reference-state
    init at a uniform phase of the clip, and per env step either random target angles inside the joint ranges or the
    clip's own joint angles at the end of the step (target = reference joint angles at time + dt)."""
    table, mdesc, _ = mim
    gid = np.arange(first_env, first_env + n_env)
    motion_id = (gid % len(clips)).astype(np.int32)
    dur = np.array([mdesc.clip_duration[c] for c in range(len(clips))], np.float64)
    phase = np.array([np.random.default_rng([seed, int(g)]).random() for g in gid])
    ref_t0 = (phase * dur[motion_id]).astype(np.float32)
    lo, hi = model.jnt_range[1:, 0].astype(np.float32), model.jnt_range[1:, 1].astype(np.float32)
    rs = np.random.default_rng([seed, 1000, first_env])
    if actions_mode == "reference":  # one batch per step of the rollout (warm-up and timed loops both start from the initial state)
        n_act = min(steps + 4, 96)
        table_id = np.array([table.names.index(c) for c in clips])[motion_id]   # clip index inside the table, not inside `clips`
        actions = np.stack([table.reference_joint_angles(table_id, ref_t0 + (k + 1) * N_FRAMES * 0.002) for k in range(n_act)]).astype(np.float32)
        return motion_id, ref_t0, actions
    n_act = min(steps, 16)  # a ring of action batches, reused
    actions = rs.uniform(lo, hi, (n_act, n_env, 28)).astype(np.float32)
    return motion_id, ref_t0, actions


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference_run(workload, steps, warmup, actions_mode="random", sample_envs=None, threads=None, budget_s=45.0):
    """The reference's CPU path for this workload: the oracle restatement of mjx.step + the PD law
    (oracle/, fp32, OpenMP over envs) on a bounded sample of the same synthetic inputs.  Runs exactly `warmup` untimed and
    `steps` timed env steps; the sample is the first `n` envs of the workload, n chosen from a calibration step so that the
    whole run takes about `budget_s` seconds on this host (capped at the workload's batch)."""
    from oracle.oracle import Oracle, OracleEnv
    util_data = importlib.import_module(PKG + ".util_data")
    model, clips, mim, n_env, pd_mode, descr = load_assets(workload)
    cores = threads or os.cpu_count() or 1
    motion_id, ref_t0, actions = synth_inputs(model, clips, mim, n_env, 0, steps + warmup, actions_mode)
    orc = Oracle(model, "f32")
    kp, kd = util_data.generate_kp_kd_gains()

    def fresh(n):
        e = OracleEnv(orc, n, nthreads=cores)
        e.reset_mimic(mim[1], motion_id[:n], ref_t0[:n])
        return e
    if sample_envs:
        n = min(n_env, sample_envs)
    else:  # calibration: one env step of a small batch -> seconds per env-step per env on this host
        n0 = min(n_env, 16 * cores)
        e0 = fresh(n0)
        e0.step_mimic(actions[0][:n0], kp, kd, pd_mode=pd_mode, pd_dt=0.002, n_frames=N_FRAMES)
        t0 = time.perf_counter()
        e0.step_mimic(actions[1 % len(actions)][:n0], kp, kd, pd_mode=pd_mode, pd_dt=0.002, n_frames=N_FRAMES)
        per_env = (time.perf_counter() - t0) / n0
        # the cost per env grows along the rollout (more bodies on the floor): leave a factor 2
        n = int(budget_s / (2.0 * per_env * max(1, steps + warmup)))
        n = max(min(n_env, 8 * cores), min(n_env, (n // cores) * cores))
    env = fresh(n)
    for k in range(warmup):
        env.step_mimic(actions[k % len(actions)][:n], kp, kd, pd_mode=pd_mode, pd_dt=0.002, n_frames=N_FRAMES)
    t0 = time.perf_counter()
    for k in range(warmup, warmup + steps):
        env.step_mimic(actions[k % len(actions)][:n], kp, kd, pd_mode=pd_mode, pd_dt=0.002, n_frames=N_FRAMES)
    dt = time.perf_counter() - t0
    return {"value": n * N_FRAMES * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"the first {n} envs of workload '{workload}' ({n_env} envs) x {steps} env steps x {N_FRAMES} substeps after {warmup} warm-up steps "
                      f"(oracle/pbrl_oracle.c, fp32, OpenMP over envs; a CPU restatement of MJX semantics, not MJX)",
            "seconds": dt, "ms_per_step": 1e3 * dt / steps, "sample_envs": n, "self_collision": self_collision_label(model),
            "description": descr}


def replay_bench(args, local_rank, first_env=0, quiet=False):
    """BASELINE configs[1]: agent_replay_motion kinematic replay of humanoid3d_walk.txt, 4096 envs, FK + obs/reward + poses.
    A step = one launch of pbrl_replay_rollout = K_INNER replay steps of every env (env e shows frame (e + k) mod F);
    the per-step launch (pbrl_replay, what HumanoidReplay.step does) is timed beside it."""
    import torch
    L = importlib.import_module(PKG + "._lib")
    mjcf = importlib.import_module(PKG + ".mjcf")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    model = mjcf.load_model(ROOT / PKG / "assets" / "final_humanoid.model.json")
    z = np.load(ROOT / PKG / "assets" / "mocap_walk.npz")
    n_env = args.envs or 4096
    K_INNER = 256   # 256 x 4096 items x 684 B = 0.72 GB written per launch: far larger than the 126 MB L2
    F = z["data_pos"].shape[0]
    dpos = torch.from_numpy(z["data_pos"].astype(np.float32)).to(dev)
    dvel = torch.from_numpy(z["data_vel"].astype(np.float32)).to(dev)
    frame0 = torch.from_numpy(((np.arange(n_env) + first_env) % F).astype(np.int32)).to(dev)
    M = L.Model(model, device=local_rank)
    e = lambda *s_: torch.empty(s_, device=dev, dtype=torch.float32)
    obs, rew, done, xpos, xquat = e(K_INNER, n_env, 69), e(K_INNER, n_env, 3), e(K_INNER, n_env), e(K_INNER, n_env, 14, 3), e(K_INNER, n_env, 14, 4)
    lib, stream = L.lib(), L.current_stream_ptr(dev)
    steps, warmup = (args.steps, max(args.warmup, 3)) if not quiet else (min(args.steps, 16), 3)

    def one():
        L.check(lib.pbrl_replay_rollout(M.handle, n_env, K_INNER, L._ptr(dpos), L._ptr(dvel), F, L._ptr(frame0), L._ptr(obs), L._ptr(rew),
                                        L._ptr(done), L._ptr(xpos), L._ptr(xquat), stream))
    def timed_launches(n):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for k in range(n):
            evs[k][0].record(); one(); evs[k][1].record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)
    # the other form of the same entry point, for the record: kinematics per item (what the launch does when the items do not
    # outnumber the clip's rows); the default below evaluates them once per clip row per launch and gathers
    def form_ms(var, val):
        before = os.environ.get(var)
        os.environ[var] = val
        for _ in range(warmup):
            one()
        torch.cuda.synchronize()
        t = timed_launches(min(steps, 16)) / min(steps, 16)
        if before is None:
            os.environ.pop(var)
        else:
            os.environ[var] = before
        return t
    t_item_ms = form_ms("PBRL_REPLAY_PER_ITEM", "1")
    # the two store paths of the frame-table form: ordinary 4-byte-per-lane stores, or rows staged in shared memory and handed to
    # the bulk-copy engine in 64-byte aligned blocks of 16 items (the launch below runs whichever the library defaults to)
    t_lsu_ms, t_bulk_ms = form_ms("PBRL_REPLAY_BULK", "0"), form_ms("PBRL_REPLAY_BULK", "1")
    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if not quiet:
        sampler.start()
    t_ms = timed_launches(steps)
    clocks = sampler.stop() if not quiet else None
    # the per-step launch the env facade makes (4096 envs x 944 B = 3.9 MB: launch-latency bound), L2 flushed in between
    qpos1, qvel1 = dpos[frame0.long()].contiguous(), dvel[frame0.long()].contiguous()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    ev1 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(32)]
    for k in range(35):
        flush.zero_()
        if k >= 3:
            ev1[k - 3][0].record()
        L.check(lib.pbrl_replay(M.handle, n_env, L._ptr(qpos1), L._ptr(qvel1), L._ptr(obs), L._ptr(rew), L._ptr(done), L._ptr(xpos),
                                L._ptr(xquat), stream))
        if k >= 3:
            ev1[k - 3][1].record()
    torch.cuda.synchronize()
    t1_ms = sum(a.elapsed_time(b) for a, b in ev1) / 32
    # end to end: frame indices from pinned host memory in, the last step's observations back, every launch
    h_f0, h_obs = frame0.cpu().pin_memory(), torch.empty(n_env, 69).pin_memory()
    torch.cuda.synchronize()
    e2e_steps = max(4, steps)
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        frame0.copy_(h_f0, non_blocking=True)
        one()
        h_obs.copy_(obs[K_INNER - 1], non_blocking=True)
        torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    peak, src = measured_peak()
    items = n_env * K_INNER
    ach = items * BYTES_PER_REPLAY_ROLLOUT_ITEM / (t_ms * 1e-3 / steps) / 1e9
    bulk = os.environ.get("PBRL_REPLAY_BULK", "1") != "0"   # the library's default store path (csrc/pbrl_lib.cu: PBRL_REPLAY_BULK_DEFAULT)
    gather = "k_replay_gather_bulk" if bulk else "k_replay_gather"
    tr = measured_traffic(gather)
    traffic = tr["dram_bytes_per_launch"] if tr and tr.get("envs_per_gpu") == n_env and tr.get("replay_steps_per_launch") == K_INNER else None
    return {"metric": "humanoid replay env-steps/sec", "value": items * steps / (t_ms * 1e-3), "unit": "env-steps/s",
            "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": t_ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "replay", "description": "BASELINE configs[1]: kinematic replay of walk, FK + obs/reward + body poses",
                       "envs_per_gpu": n_env, "replay_steps_per_launch": K_INNER, "clip_rows": int(F),
                       "kinematics": "once per clip row per launch (frame table), rows gathered per item; per-item form reported beside it",
                       "stores": "rows of 16 consecutive items staged in shared memory, written by the bulk-copy engine (64-byte aligned blocks); ordinary stores reported beside it" if bulk else "ordinary 4-byte-per-lane streaming stores",
                       "l2": "each launch writes 0.72 GB (> 126 MB L2); inputs are the 10 KB clip"},
            "e2e": {"value": items / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": n_env * 4, "d2h_bytes_per_step": n_env * 69 * 4, "steps": e2e_steps,
                    "api": "pbrl_replay_rollout, frame indices from pinned host memory, last observations back"},
            "gpu_launches": 2 * (steps + warmup + e2e_steps) + 5 * (warmup + min(steps, 16)) + 35,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "peak_source": src, "kernel": gather + " (+ k_replay_frames, both inside the timed launch)",
                         "algorithmic_bytes_per_env_launch": BYTES_PER_REPLAY_ROLLOUT_ITEM,
                         "note": "per (env, frame) item: obs 276 + reward 12 + done 4 + xpos 168 + xquat 224 B written; the 10 KB clip and the frame table it reads stay on chip"},
            "per_item_kinematics": {"kernel": "k_replay_rollout (PBRL_REPLAY_PER_ITEM=1: kinematics per (env, frame) item, no frame table)", "ms_per_step": t_item_ms,
                                    "env_steps_per_s": items / (t_item_ms * 1e-3),
                                    "roofline_frac": items * BYTES_PER_REPLAY_ROLLOUT_ITEM / (t_item_ms * 1e-3) / 1e9 / peak},
            "store_paths": {"ordinary_stores": {"kernel": "k_replay_gather (PBRL_REPLAY_BULK=0)", "ms_per_step": t_lsu_ms,
                                                "roofline_frac": items * BYTES_PER_REPLAY_ROLLOUT_ITEM / (t_lsu_ms * 1e-3) / 1e9 / peak},
                            "bulk_stores": {"kernel": "k_replay_gather_bulk (PBRL_REPLAY_BULK=1: cp.async.bulk.global.shared::cta)", "ms_per_step": t_bulk_ms,
                                            "roofline_frac": items * BYTES_PER_REPLAY_ROLLOUT_ITEM / (t_bulk_ms * 1e-3) / 1e9 / peak}},
            "per_step_launch": {"kernel": "k_replay", "ms": t1_ms, "env_steps_per_s": n_env / (t1_ms * 1e-3),
                                "algorithmic_bytes_per_env_launch": BYTES_PER_REPLAY_STEP_ENV,
                                "achieved_GBps": n_env * BYTES_PER_REPLAY_STEP_ENV / (t1_ms * 1e-3) / 1e9},
            "clocks": clocks}


def measured_traffic(kernel="k_step"):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture of this round."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        p = ROOT / "profiles" / name
        if p.exists():
            try:
                d = json.loads(p.read_text())
                if kernel in d:
                    return d[kernel]
                if d.get("kernel") == kernel:
                    return d
            except Exception:
                pass
    return None


class Dynamics:
    """One dynamics workload on one GPU: model, state buffers, synthetic inputs, the step closure."""

    def __init__(self, workload, n_env, first_env, local_rank, steps, actions_mode, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.world = torch, dist, world
        L = self.L = importlib.import_module(PKG + "._lib")
        util_data = importlib.import_module(PKG + ".util_data")
        self.dev = dev = torch.device("cuda", local_rank)
        self.model, self.clips, self.mim, default_env, self.pd_mode, self.descr = load_assets(workload)
        self.n_env = n_env = n_env or default_env
        self.workload = workload
        self.motion_id, self.ref_t0, self.actions = synth_inputs(self.model, self.clips, self.mim, n_env, first_env, steps, actions_mode)
        self.M = L.Model(self.model, device=local_rank, mimic=self.mim[1])
        z = lambda *s: torch.zeros(s, device=dev, dtype=torch.float32)
        self.stats2 = z(2, 8)   # double-buffered statistics: the all-reduce of step k overlaps the kernel of step k + 1
        st = self.st = dict(qpos=z(n_env, 35), qvel=z(n_env, 34), qpos_lag=z(n_env, 35), qvel_lag=z(n_env, 34), time=z(n_env),
                            warm=z(n_env, 34), ctrl=z(n_env, 28), obs=z(n_env, 69), reward=z(n_env, 3), done=z(n_env), metrics=z(n_env, 3),
                            status=torch.zeros(n_env, device=dev, dtype=torch.int32),
                            motion_id=torch.from_numpy(self.motion_id).to(dev), ref_t0=torch.from_numpy(self.ref_t0).to(dev),
                            # processing order: envs of the same clip share a CTA (similar contact work -> balanced phase barriers)
                            env_order=torch.from_numpy(np.argsort(self.motion_id, kind="stable").astype(np.int32)).to(dev),
                            cost=torch.zeros(n_env, device=dev, dtype=torch.int32))
        self.bufs = [L.make_buffers(stats=self.stats2[i], **st) for i in range(2)]
        self.lib = L.lib()
        self.stream = L.current_stream_ptr(dev)
        L.check(self.lib.pbrl_reset(self.M.handle, n_env, 1, ctypes.byref(self.bufs[0]), self.stream))  # reference-state init
        torch.cuda.synchronize()
        self.init = {k: v.clone() for k, v in st.items()}
        self.act_dev = torch.from_numpy(self.actions).to(dev)
        self.kp, self.kd = [torch.from_numpy(x).to(dev) for x in util_data.generate_kp_kd_gains()]
        self.dur = torch.tensor([self.mim[1].clip_duration[c] for c in range(len(self.clips))], device=dev)[st["motion_id"].long()]
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(1234 + first_env)
        self.pending = [None, None]
        self.balance = not os.environ.get("PBRL_BENCH_NO_BALANCE")
        self.launches = 0
        self.step_launches = int(self.lib.pbrl_step_launches(self.M.handle, n_env, N_FRAMES, ctypes.byref(self.bufs[0])))

    def restart(self):
        for k2, v in self.init.items():
            self.st[k2].copy_(v)

    def step(self, k, autoreset=False):
        L, lib, st, i = self.L, self.lib, self.st, k & 1
        if self.pending[i] is not None:   # the all-reduce that last used this statistics buffer (two steps ago) must have landed
            self.pending[i].wait()
            self.pending[i] = None
        a = self.act_dev[k % self.act_dev.shape[0]]
        self.stats2[i].zero_()
        L.check(lib.pbrl_step(self.M.handle, self.n_env, N_FRAMES, self.pd_mode, 1, L._ptr(a), L._ptr(self.kp), L._ptr(self.kd), ctypes.c_float(0.002),
                              None, ctypes.c_uint(0x0FFFFFFF), 1, ctypes.byref(self.bufs[i]), self.stream))
        self.launches += self.step_launches   # 1, or 3 when the library regroups the envs inside the step (two k_step parts + the sort)
        if self.world > 1:
            # per-step reward / termination statistics all-reduce (32 B, NCCL over NVLink): asynchronous on NCCL's own stream,
            # which waits for k_step; the compute stream goes straight on to the next step and joins before the buffer is reused
            self.pending[i] = self.dist.all_reduce(self.stats2[i], async_op=True)
        if autoreset:  # finished envs restart at a fresh clip phase (brax AutoResetWrapper semantics): the workload stays stationary
            nxt = self.torch.rand(self.n_env, generator=self.gen, device=self.dev) * self.dur
            L.check(lib.pbrl_autoreset(self.M.handle, self.n_env, 1, L._ptr(st["done"]), L._ptr(nxt), None, ctypes.byref(self.bufs[i]), self.stream))
            self.launches += 1
        if self.balance:  # processing order of the next step from this step's work estimates
            L.check(lib.pbrl_balance_order(self.M.handle, self.n_env, L._ptr(st["cost"]), L._ptr(st["env_order"]), self.stream))
            self.launches += 1

    def join_stats(self):
        for i in range(2):
            if self.pending[i] is not None:
                self.pending[i].wait()
                self.pending[i] = None

    def sync_all(self):
        self.join_stats()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, steps, warmup, flush, autoreset=False, rollout_warmup=0):
        """`warmup` untimed steps, restart from the initial states, (`rollout_warmup` untimed steps that are part of the rollout,)
        then EXACTLY `steps` timed steps: CUDA events on the launch stream around each step, L2 evicted in between (outside
        the events).  The statistics all-reduce of a step is joined inside the timed region of the next one (the last one
        inside its own).  Returns the sum of the per-step times in ms (this rank)."""
        torch = self.torch
        for k in range(warmup):
            self.step(k, autoreset)
        self.sync_all()
        self.restart()
        for k in range(rollout_warmup):
            self.step(k, autoreset)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.sync_all()
        for k in range(steps):
            flush.zero_()  # evict the 126 MB L2 between timed iterations (outside the timed events)
            ev[k][0].record()
            self.step(rollout_warmup + k, autoreset)
            if k == steps - 1:
                self.join_stats()
            ev[k][1].record()
        self.sync_all()
        return sum(a.elapsed_time(b) for a, b in ev)

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def status_counts(self):
        s = self.st["status"]
        return {"nan_envs": int((s & 1).sum().item()), "pivot_envs": int(((s & 2) != 0).sum().item()),
                "solver_maxiter_envs": int(((s & 4) != 0).sum().item()), "self_contact_overflow_envs": int(((s & 8) != 0).sum().item())}

    def e2e(self, steps, flush):
        """The same metric through the host-buffer C-ABI layer: pinned host actions in, obs / reward / done / stats out, every
        step; wall clock around each call (the call returns after its own stream synchronize), L2 evicted between calls
        outside the clock, like the event-timed loop."""
        torch, L, lib, n_env = self.torch, self.L, self.lib, self.n_env
        util_data = importlib.import_module(PKG + ".util_data")
        batch = ctypes.c_void_p()
        L.check(lib.pbrl_batch_create(self.M.handle, n_env, ctypes.byref(batch)))
        h_act = torch.from_numpy(self.actions[: min(4, len(self.actions))].copy()).pin_memory()
        h_obs, h_rew, h_done = [torch.empty(s, dtype=torch.float32).pin_memory() for s in ((n_env, 69), (n_env, 3), (n_env,))]
        h_kp, h_kd = [torch.from_numpy(x.copy()).pin_memory() for x in util_data.generate_kp_kd_gains()]
        h_mid, h_t0, h_stats = torch.from_numpy(self.motion_id.copy()).pin_memory(), torch.from_numpy(self.ref_t0.copy()).pin_memory(), torch.zeros(8).pin_memory()
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        L.check(lib.pbrl_batch_reset_host(batch, 1, p(h_mid), p(h_t0), p(h_obs)))

        def e2e_step(k):
            L.check(lib.pbrl_batch_step_host(batch, N_FRAMES, self.pd_mode, 1, p(h_act[k % h_act.shape[0]]), p(h_kp), p(h_kd), ctypes.c_float(0.002),
                                             1, p(h_obs), p(h_rew), p(h_done), p(h_stats)))
        for k in range(3):
            e2e_step(k)
        L.check(lib.pbrl_batch_reset_host(batch, 1, p(h_mid), p(h_t0), p(h_obs)))  # the timed rollout starts from the named initial states
        self.sync_all()
        total = 0.0
        for k in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_step(k)
            total += time.perf_counter() - t0
        lib.pbrl_batch_destroy(batch)
        self.launches += (self.step_launches + 1) * (steps + 3) + 2   # per host-buffer step: pbrl_step's kernels + the balance sort
        return self.max_over_ranks(total)


def run_extra(workload, args, local_rank, flush):
    """A secondary workload (BASELINE configs 0-3) measured inside the default run, so that the driver's BENCH record carries it."""
    steps = min(args.steps, 16)
    d = Dynamics(workload, 0, 0, local_rank, steps, args.actions, 1)
    t_ms = d.timed(steps, 3, flush)
    out = {"value": d.n_env * N_FRAMES * steps / (t_ms * 1e-3), "unit": UNIT, "ms_per_step": t_ms / steps, "steps": steps, "warmup": 3,
           "envs_per_gpu": d.n_env, "description": d.descr, "hbm_frac": d.n_env * BYTES_PER_ENV_LAUNCH / (t_ms * 1e-3 / steps) / 1e9 / measured_peak()[0]}
    out.update(d.status_counts())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mimic_mixed", choices=sorted(WORKLOADS) + ["replay"])
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's batch per GPU (default); strong: BASELINE configs[4] as written, 65536 envs in total divided over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads (extra_configs) and the stationary (auto-reset) variant")
    ap.add_argument("--actions", default="random", choices=["random", "reference"],
                    help="PD targets: random inside the joint ranges (default, the north star's random-action rollouts) or the clip's joint angles at time + dt")
    ap.add_argument("--stage-profile", action="store_true", help="dev: print the per-stage cycle table (needs a -DPBRL_STAGE_PROF library, PBRL_LIB=...)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = args.steps, max(args.warmup, 3)

    if args.workload == "replay" and args.impl == "ours":
        if rank == 0:
            print(json.dumps(replay_bench(args, local_rank)))
        return 0
    if args.impl == "reference":
        if rank != 0:
            return 0
        wl = "mimic_mixed" if args.workload == "replay" else args.workload
        r = cpu_reference_run(wl, steps, warmup, args.actions)
        n = r["sample_envs"]
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                # the same keys as our arm prints; envs_per_gpu / envs_total are the CPU sample actually stepped
                "config": {"workload": wl, "description": r["description"], "envs_per_gpu": n, "envs_total": n,
                           "n_frames": N_FRAMES, "pd_mode": "stable", "reward_mode": "mimic", "actions": args.actions, "l2_flush_between_steps": False,
                           "self_collision": r["self_collision"], "parallelism": f"{r['cores']} host threads (OpenMP over envs)",
                           "sample": f"first {n} of the workload's {WORKLOADS[wl][2]} envs per GPU"},
                "cpu_baseline": {k2: r[k2] for k2 in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_env = args.envs
    if args.scaling == "strong":
        n_env = STRONG_TOTAL_ENVS // world
    n_default = WORKLOADS[args.workload][2]
    D = Dynamics(args.workload, n_env, rank * (n_env or n_default), local_rank, steps, args.actions, world)
    n_env = D.n_env
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)  # > 126 MB L2
    lib, L = D.lib, D.L

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if args.stage_profile:
        L.check(lib.pbrl_debug_stage_profile(None, None, 1))
    t_ms = D.max_over_ranks(D.timed(steps, warmup, flush))
    clocks = sampler.stop() if rank == 0 else None
    if args.stage_profile and rank == 0:
        cyc, hit = (ctypes.c_ulonglong * 64)(), (ctypes.c_ulonglong * 64)()
        L.check(lib.pbrl_debug_stage_profile(cyc, hit, 0))
        names = {0: "load", 1: "pd_torque", 2: "fk", 3: "com", 4: "crb", 5: "collide", 6: "comvel", 7: "bias", 8: "rows",
                 9: "fsmooth", 10: "free:factor", 11: "free:solve", 12: "newton:init", 13: "newton:grad(JTf)", 14: "newton:hessian",
                 15: "newton:factor", 16: "newton:solve", 17: "newton:Mp,Jp", 18: "newton:linesearch", 19: "newton:update",
                 20: "fcon(JTf)", 21: "euler:factor", 22: "euler:solve", 23: "integrate", 24: "store", 25: "epilogue", 27: "barrier: substep start", 28: "barrier: after position", 29: "barrier: after rows", 30: "barrier: after solver", 31: "barrier: after collide"}
        nsub = n_env * (steps + warmup) * N_FRAMES
        tot = sum(cyc[:40])
        print(f"stage profile: {tot / nsub:.0f} cycles per env-substep (per warp, incl. stalls)", file=sys.stderr)
        print("  newton iterations per substep: " + " ".join(f"{i - 40}:{hit[i] / nsub:.3f}" for i in range(40, 56) if hit[i]), file=sys.stderr)
        if hit[60]:  # lock-step cost: Newton votes a warp attends vs iterations it runs itself
            print(f"  newton votes attended per substep {hit[60] / nsub:.3f}, iterations run {hit[61] / nsub:.3f}, of which with the dense Hessian {hit[58] / nsub:.4f}", file=sys.stderr)
            print(f"  substeps in which a warp runs the convex narrow phase (capsule-box / box-box): {hit[63] / max(hit[62], 1):.4f}", file=sys.stderr)
        for i in range(40):
            if hit[i]:
                print(f"  {i:2d} {names.get(i, '?'):22s} {100.0 * cyc[i] / tot:6.2f}%  hits/substep {hit[i] / nsub:7.3f}  "
                      f"cycles/hit {cyc[i] / hit[i]:9.0f}", file=sys.stderr)
    counts = D.status_counts()
    stats_last = D.stats2[(steps - 1) & 1].cpu().numpy().tolist()   # all-reduced over the ranks when world > 1

    # end to end through the host-buffer C-ABI layer, as many timed steps as the device-timed loop
    e2e_s = D.e2e(steps, flush)

    # stationary variant: finished envs restart inside the loop (pbrl_autoreset), timed after the rollout has run for a while
    stationary = None
    if not args.no_extra:
        pre = 16
        ts_ms = D.max_over_ranks(D.timed(steps, 3, flush, autoreset=True, rollout_warmup=pre))
        stationary = {"value": n_env * world * N_FRAMES * steps / (ts_ms * 1e-3), "unit": UNIT, "ms_per_step": ts_ms / steps, "steps": steps,
                      "rollout_steps_before_timing": pre,
                      "note": "pbrl_autoreset after every step (inside the timed region): finished envs restart at a fresh clip phase, so the "
                              "timed workload does not drift towards everybody lying on the floor"}
        stationary.update(D.status_counts())

    extra = None
    if not args.no_extra and world == 1 and args.workload == "mimic_mixed" and not args.envs and args.scaling == "weak":
        extra = {}
        for wl in ("contact_free", "contact_rich", "cpu64"):
            try:
                extra[wl] = run_extra(wl, args, local_rank, flush)
            except Exception as ex:  # a secondary workload must never cost the headline line
                extra[wl] = {"error": str(ex)[:200]}
        try:
            r = replay_bench(args, local_rank, quiet=True)
            extra["replay"] = {k2: r[k2] for k2 in ("metric", "value", "unit", "ms_per_step", "steps", "roofline", "store_paths", "per_item_kinematics", "per_step_launch", "e2e")}
            extra["replay"]["envs_per_gpu"] = r["config"]["envs_per_gpu"]
            extra["replay"]["replay_steps_per_launch"] = r["config"]["replay_steps_per_launch"]
        except Exception as ex:
            extra["replay"] = {"error": str(ex)[:200]}

    if rank == 0:
        peak, peak_src = measured_peak()
        total_envs = n_env * world
        value = total_envs * N_FRAMES * steps / (t_ms * 1e-3)
        launch_s = t_ms * 1e-3 / steps
        achieved = n_env * BYTES_PER_ENV_LAUNCH / launch_s / 1e9
        tr = measured_traffic("k_step")
        traffic = tr["dram_bytes_per_launch"] if tr and tr.get("workload") == args.workload and tr.get("envs_per_gpu") == n_env else None
        groups = -(-n_env // 16)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": t_ms / steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": args.workload, "description": D.descr, "envs_per_gpu": n_env, "envs_total": total_envs,
                           "n_frames": N_FRAMES, "pd_mode": "stable", "reward_mode": "mimic", "actions": args.actions, "l2_flush_between_steps": True,
                           "self_collision": self_collision_label(D.model),
                           "parallelism": (f"env-sharded x{world}, stats all-reduce per step (asynchronous, double-buffered, joined one step later)"
                                           if world > 1 else "single GPU"),
                           "cta_groups_per_gpu": groups, "groups_per_cta": groups / 148.0},
                "e2e": {"value": total_envs * N_FRAMES * steps / e2e_s, "unit": UNIT,
                        "h2d_bytes_per_step": n_env * 28 * 4 + 2 * 28 * 4, "d2h_bytes_per_step": n_env * (69 + 3 + 1) * 4 + 32,
                        "steps": steps, "api": "pbrl_batch_step_host (C ABI, pinned host buffers)",
                        "note": "wall clock around each host call (it returns after its own stream synchronize); L2 evicted between calls outside the clock; "
                                "the buffers are pinned and device-mapped, so k_step reads the action rows and writes obs / reward / done in place over PCIe "
                                "(the byte counts are the rows that cross the bus per step; PBRL_HOST_ZERO_COPY=0 stages them through cudaMemcpyAsync instead)"},
                # kernels launched inside the timed regions and their warm-ups: k_step (+ k_balance_order, + k_autoreset in the stationary loop), k_reset
                "gpu_launches": D.launches + 1,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src, "kernel": "k_step",
                             "algorithmic_bytes_per_env_launch": BYTES_PER_ENV_LAUNCH,
                             "note": "the dynamics step is fp32-issue/latency bound (50-300 flop/B); HBM fraction is reported per the tier rule"},
                # what actually bounds k_step: warp-instruction issue (ncu instruction count of the committed capture x live rate
                # against 4 issue slots per SM-cycle at the sampled SM clock)
                "sm_issue": None if not (tr and tr.get("warp_instructions_per_env_substep") and clocks and clocks.get("sm_mhz")) else {
                    "warp_instructions_per_env_substep": tr["warp_instructions_per_env_substep"],
                    "achieved_Ginst_per_s": tr["warp_instructions_per_env_substep"] * value / world / 1e9,
                    "peak_Ginst_per_s": 148 * 4 * clocks["sm_mhz"] * 1e6 / 1e9,
                    "frac": tr["warp_instructions_per_env_substep"] * value / world / (148 * 4 * clocks["sm_mhz"] * 1e6),
                    "source": tr.get("source")},
                "clocks": clocks,
                "reward_stats_last_step": {"mean_reward": stats_last[0] / max(stats_last[5], 1), "done_frac": stats_last[2] / max(stats_last[5], 1),
                                           "fall_frac": stats_last[3] / max(stats_last[5], 1), "n": stats_last[5]}}
        line.update(counts)
        if stationary is not None:
            line["stationary"] = stationary
        if extra is not None:
            line["extra_configs"] = extra
        if not args.no_cpu_baseline and world == 1:  # the CPU leg runs beside the single-GPU number only
            r = cpu_reference_run(args.workload, 4, 1, args.actions, budget_s=15.0)
            line["cpu_baseline"] = {k2: r[k2] for k2 in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
