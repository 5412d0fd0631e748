"""N > 1 host logic on CPU: world_size-2 gloo processes shard the envs, step their shard (oracle as the per-env
compute -- test infrastructure) and all-reduce the statistics; the result must equal the single-process run and
every env must see the same clip / phase whatever the shard count."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ASSETS, PKG, Oracle, OracleEnv, load_model_asset, mimic, mocap, util_data

pdist = importlib.import_module(PKG + ".dist")
N_GLOBAL = 13  # odd on purpose: ragged shards


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_shard(rank, world):
    model = load_model_asset("final_humanoid")
    table = mocap.MotionTable.load(ASSETS / "motions.npz")
    mdesc, keep = mimic.make_mimic_desc(model, table)
    ids = pdist.global_env_ids(N_GLOBAL, rank, world)
    durs = [float(mdesc.clip_duration[c]) for c in range(mdesc.nclip)]
    mid, t0 = pdist.env_assignment(ids, durs, seed=7)
    env = OracleEnv(Oracle(model, "f64", tolerance=1e-12), len(ids))
    env.reset_mimic(mdesc, mid, t0)
    kp, kd = util_data.generate_kp_kd_gains()
    stats = env.step_mimic(env.qpos[:, 7:].astype(np.float32), kp, kd)
    return ids, mid, t0, env.reward[:, 0].copy(), stats


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids, mid, t0, r, stats = _run_shard(rank, world)
    t = torch.tensor(stats, dtype=torch.float64)
    pdist.allreduce_stats(t)
    out[rank] = (ids.tolist(), mid.tolist(), t0.tolist(), r.tolist(), t.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    for world in (1, 2, 3, 4, 8):
        got = np.concatenate([pdist.global_env_ids(N_GLOBAL, r, world) for r in range(world)])
        assert np.array_equal(got, np.arange(N_GLOBAL))
    assert pdist.shard_range(65536, 3, 4) == (49152, 65536)


def test_two_rank_gloo_stats_allreduce_matches_single_process():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    ids1, mid1, t01, r1, stats1 = _run_shard(0, 1)
    ids = out[0][0] + out[1][0]
    assert ids == ids1.tolist()
    assert out[0][1] + out[1][1] == mid1.tolist()                       # same clip per global env id
    assert np.allclose(out[0][2] + out[1][2], t01)                      # same phase per global env id
    assert np.allclose(out[0][3] + out[1][3], r1, rtol=0, atol=1e-12)   # env i is shard-count independent
    assert np.allclose(out[0][4], out[1][4])                            # both ranks hold the reduced vector
    assert np.allclose(out[0][4][:6], stats1[:6], rtol=1e-12)
    s = pdist.summarize_stats(np.array(out[0][4]))
    assert s["n"] == N_GLOBAL and 0 < s["mean_reward"] <= 1
