"""Multi-GPU equivalence check of the sharded PPO update, callable from a
torchrun worker (tests/dist_worker_gpu.py) and from bench.py after its timed region.

W ranks each own a shard of the envs of ONE global synthetic rollout.  After the same global
minibatch schedule the parameters must (a) be bit-identical on every rank and (b) equal a single-GPU
run over the whole rollout up to summation order, for both update paths: per-minibatch launches +
NCCL all-reduce ("launches"), and the fused epoch kernel with its in-kernel NVLink all-reduce
("fused").  The single-GPU side is the same library on rank 0 (the library itself is held to the
oracle by tests/test_ppo_update_gpu.py); nothing here touches oracle/.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import sharding
from .policy import sb3_initial_state_dict
from .updater import PeerExchange, PpoUpdater


def sharded_update_check(dev: torch.device, modes=("launches", "fused"), obs_dim: int = 14, T: int = 32,
                         n_local: int = 24, b_local: int = 150, epochs: int = 2) -> dict:
    rank, world = dist.get_rank(), dist.get_world_size()
    O = obs_dim
    N, B = n_local * world, b_local * world
    rng = np.random.default_rng(0)  # identical global rollout on every rank
    state = torch.get_rng_state()   # the caller's torch stream is left as it was
    torch.manual_seed(0)
    sd = sb3_initial_state_dict(O)
    torch.set_rng_state(state)
    flat0 = torch.cat([v.reshape(-1) for v in sd.values()])
    flat0[:2] = torch.tensor([-0.3, 0.2])
    full = dict(obs=rng.standard_normal((T, N, O)).astype(np.float32),
                actions=rng.standard_normal((T, N, 2)).astype(np.float32),
                log_probs=(rng.standard_normal((T, N)) * 0.1 - 2.0).astype(np.float32),
                advantages=rng.standard_normal((T, N)).astype(np.float32),
                returns=rng.standard_normal((T, N)).astype(np.float32))
    # rank-local permutations (what PPO.train draws); every rank can rebuild all of them
    local_perms = [[np.random.default_rng(100 * e + r).permutation(n_local * T).astype(np.int64)
                    for r in range(world)] for e in range(epochs)]
    n_mb = (n_local * T + b_local - 1) // b_local
    kw = dict(clip_range=0.2, ent_coef=0.05, vf_coef=0.5, normalize_advantage=True)

    def dev_buf(arrs):
        return {k: torch.as_tensor(np.ascontiguousarray(v)).to(dev) for k, v in arrs.items()}

    mine = dev_buf({k: v[:, rank * n_local:(rank + 1) * n_local] for k, v in full.items()})

    def sharded(mode):
        up = PpoUpdater(O, dev, **kw)
        up.params.copy_(flat0)
        xchg = PeerExchange(O, dev) if mode == "fused" else None
        info = torch.zeros((n_mb, 8), device=dev)
        for e in range(epochs):
            perm = torch.as_tensor(local_perms[e][rank]).to(dev)
            stats = up.adv_stats(mine["advantages"], perm, b_local, n_local, T)
            stats, share = sharding.allreduce_adv_stats(stats)
            if mode == "fused":
                up.train_epoch_fused(mine, perm, stats, b_local, n_local, T, info, xchg)
            else:
                sh = share.cpu().tolist()
                for m in range(n_mb):
                    up.compute_grad(mine, perm[m * b_local:(m + 1) * b_local], stats[m], n_local, T, sh[m])
                    dist.all_reduce(up.grad)
                    up.adam_step(info[m])
        torch.cuda.synchronize(dev)
        gathered = [torch.empty_like(up.params) for _ in range(world)]
        dist.all_gather(gathered, up.params)
        identical = all(torch.equal(gathered[0], g) for g in gathered)
        timed_out = False
        if xchg is not None:
            timed_out = xchg.timed_out()
            dist.barrier()
            xchg.close()
        return up.params.clone(), identical, timed_out

    results = {m: sharded(m) for m in modes}
    out = {"world": world, "modes": list(modes)}
    moved = torch.zeros(1, device=dev, dtype=torch.float64)
    diffs = torch.zeros(len(modes), device=dev, dtype=torch.float64)
    if rank == 0:
        one = PpoUpdater(O, dev, **kw)
        one.params.copy_(flat0)
        whole = dev_buf(full)
        for e in range(epochs):
            glob = []
            for m in range(n_mb):
                for r in range(world):
                    ids = local_perms[e][r][m * b_local:(m + 1) * b_local]
                    glob.append(ids + r * n_local * T)  # local env-major id -> global env-major id
            dp = torch.as_tensor(np.concatenate(glob)).to(dev)
            st = one.adv_stats(whole["advantages"], dp, B, N, T)
            one.train_epoch(whole, dp, st, B, N, T)
        torch.cuda.synchronize(dev)
        moved[0] = float((one.params - flat0.to(dev)).abs().max())
        for i, m in enumerate(modes):
            diffs[i] = float((one.params - results[m][0]).abs().max())
    dist.broadcast(moved, src=0)
    dist.broadcast(diffs, src=0)
    out["moved"] = float(moved[0])
    ok = out["moved"] > 1e-3
    for i, m in enumerate(modes):
        rel = float(diffs[i]) / max(out["moved"], 1e-30)
        out[m] = {"identical_across_ranks": bool(results[m][1]), "max_abs_vs_single": float(diffs[i]),
                  "max_rel_vs_single": rel, "exchange_timed_out": bool(results[m][2])}
        ok = ok and results[m][1] and rel < 2e-3 and not results[m][2]
    out["ok"] = bool(ok)
    return out
