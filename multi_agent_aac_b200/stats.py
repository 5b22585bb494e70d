"""Sharding and episode statistics across GPUs.

Envs never interact, so the step path has no collective: rank r owns the global env ids
[r*E_local, (r+1)*E_local) and scenario selection hashes the GLOBAL id (include/aac_env.h
`env_id_base`), so an env sees the same episodes however the job is sharded.  The only exchange is the
sum of the per-rank episode counters -- the quantities the reference prints every 100 episodes
(ATT/ma_main:581-637: collisions by type, goal reaching, returns) -- one all-reduce of 16 doubles over
NCCL (NVLink) on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from ._capi import N_STATS, STAT_NAMES


def shard_range(rank: int, world: int, n_envs_total: int):
    """Contiguous, balanced env-id range of `rank` (first `n_envs_total % world` ranks get one extra)."""
    base, extra = divmod(n_envs_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_episode_stats(local: np.ndarray, device=None, group=None) -> dict:
    """Sum the [N_STATS] counter vector over all ranks and derive the reference's report."""
    t = torch.as_tensor(np.asarray(local, dtype=np.float64))
    assert t.numel() == N_STATS
    if dist.is_available() and dist.is_initialized():
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t = t.cpu()
    v = t.numpy()
    out = {k: float(x) for k, x in zip(STAT_NAMES, v)}
    ep = max(out["episodes"], 1.0)
    out["mean_return"] = out["return_sum"] / ep
    out["mean_length"] = out["steps"] / ep
    out["crash_rate"] = (out["bound_crash"] + out["building_crash"] + out["drone_crash"]) / ep
    out["all_reached_rate"] = out["all_reached"] / ep
    return out
