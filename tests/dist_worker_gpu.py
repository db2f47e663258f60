"""torchrun worker: W ranks each own a shard of the envs of ONE global rollout.  After the same
global minibatch schedule the parameters must (a) be bit-identical on every rank and (b) equal a
single-GPU run on the whole rollout, for both update paths: per-minibatch launches + NCCL
all-reduce, and the fused epoch kernel with the in-kernel NVLink all-reduce."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from mobrob_b200 import sharding
from mobrob_b200.policy import sb3_initial_state_dict
from mobrob_b200.updater import PeerExchange, PpoUpdater


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    O, T, n_local, b_local, epochs = 14, 32, 24, 150, 2
    N, B = n_local * world, b_local * world
    rng = np.random.default_rng(0)  # identical global rollout on every rank
    torch.manual_seed(0)
    sd = sb3_initial_state_dict(O)
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
            share = share.to(torch.float32).contiguous()
            if mode == "fused":
                up.train_epoch_fused(mine, perm, stats, b_local, n_local, T, info, xchg)
            else:
                sh = share.cpu().tolist()
                for m in range(n_mb):
                    up.compute_grad(mine, perm[m * b_local:(m + 1) * b_local], stats[m], n_local, T, sh[m])
                    dist.all_reduce(up.grad)
                    up.adam_step(info[m])
        torch.cuda.synchronize()
        gathered = [torch.empty_like(up.params) for _ in range(world)]
        dist.all_gather(gathered, up.params)
        identical = all(torch.equal(gathered[0], g) for g in gathered)
        if xchg is not None:
            dist.barrier()
            xchg.close()
        return up.params.clone(), identical

    p_nccl, same_nccl = sharded("launches")
    p_fused, same_fused = sharded("fused")

    ok = True
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
        torch.cuda.synchronize()
        moved = float((one.params - flat0.to(dev)).abs().max())
        d_nccl = float((one.params - p_nccl).abs().max())
        d_fused = float((one.params - p_fused).abs().max())
        ok = same_nccl and same_fused and moved > 1e-3 and d_nccl < 2e-3 * moved and d_fused < 2e-3 * moved
        print(f"DIST_RESULT world={world} identical_across_ranks nccl={same_nccl} fused={same_fused} "
              f"moved={moved:.3e} max|nccl - single|={d_nccl:.3e} max|fused - single|={d_fused:.3e} ok={ok}",
              flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
