"""Host-side sharding logic for one-process-per-GPU training.

Envs are split contiguously: rank r of W owns global env ids [r * n_local, (r + 1) * n_local).
Seeds are keyed by the GLOBAL id (seed + id), so an env's random streams do not depend on W.
A global minibatch is the union of the ranks' local minibatches; advantage normalisation uses the
all-reduced (sum, sum of squares, count) and each rank's gradient is already divided by the global
count, so the all-reduce(sum) of gradients is the gradient of the global-mean loss.
Pure torch / numpy: works on CPU tensors with the gloo backend (tests/test_sharding_cpu.py).
"""
from __future__ import annotations

import numpy as np
import torch


def global_env_ids(rank: int, n_local: int) -> np.ndarray:
    return rank * n_local + np.arange(n_local, dtype=np.int64)


def split_global_minibatch(global_ids: np.ndarray, rank: int, n_local: int, T: int) -> np.ndarray:
    """Entries of a GLOBAL env-major minibatch (ids n_global * T + t) that live on `rank`,
    re-indexed to the rank's local env-major ids (n_local_idx * T + t), order preserved."""
    g = np.asarray(global_ids, dtype=np.int64)
    n = g // T
    mine = (n >= rank * n_local) & (n < (rank + 1) * n_local)
    return g[mine] - rank * n_local * T


def allreduce_adv_stats(stats: torch.Tensor):
    """stats [n_mb, 3] = (sum, sum of squares, count) per local minibatch -> (global stats,
    rank_share [n_mb] = local count / global count)."""
    import torch.distributed as dist

    local_cnt = stats[:, 2].clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats)
    share = torch.where(stats[:, 2] > 0, local_cnt / stats[:, 2], torch.zeros_like(local_cnt))
    return stats, share


def mean_std_from_stats(stats: torch.Tensor):
    """Unbiased (N - 1) std like torch.Tensor.std(), from (sum, sum of squares, count)."""
    s, q, c = stats[:, 0], stats[:, 1], stats[:, 2]
    mean = s / c
    var = (q - s * mean) / (c - 1)
    return mean, torch.sqrt(torch.clamp(var, min=0))
