"""Multi-GPU plumbing: environments shard trivially, one tiny all-reduce per env step for the reward / termination
statistics.  One process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).

Inputs are keyed by the GLOBAL env id, so env i sees the same clip, phase and action stream whatever the shard count.
"""
import numpy as np
import torch
import torch.distributed as dist

STAT_NAMES = ("sum_reward", "sum_reward_sq", "sum_done", "sum_fall", "sum_pose_error", "count")


def shard_range(n_global, rank, world):
    """Contiguous block of global env ids owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_env_ids(n_global, rank, world):
    lo, hi = shard_range(n_global, rank, world)
    return np.arange(lo, hi)


def env_assignment(global_ids, clip_durations, seed=0):
    """clip = id mod nclip, reference phase ~ U[0,1) from a counter-based stream keyed by (seed, id)."""
    nclip = len(clip_durations)
    motion_id = (global_ids % nclip).astype(np.int32)
    phase = np.array([np.random.default_rng([seed, int(g)]).random() for g in global_ids])
    ref_t0 = (phase * np.asarray(clip_durations)[motion_id]).astype(np.float32)
    return motion_id, ref_t0


def allreduce_stats(stats: torch.Tensor, group=None):
    """Sum the per-rank [8] statistics vector in place (the only collective on the path)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def summarize_stats(stats):
    s = stats.detach().cpu().numpy().astype(np.float64) if isinstance(stats, torch.Tensor) else np.asarray(stats, np.float64)
    n = max(s[5], 1.0)
    mean = s[0] / n
    return {"n": s[5], "mean_reward": mean, "var_reward": max(s[1] / n - mean * mean, 0.0), "done_frac": s[2] / n,
            "fall_frac": s[3] / n, "mean_pose_error": s[4] / n}
