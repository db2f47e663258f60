"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the sharding helpers."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mobrob_b200 import seeding, sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local, T, B = 6, 10, 16
    n_glob = n_local * world
    rng = np.random.default_rng(0)  # same data on every rank
    adv = rng.standard_normal((T, n_glob))
    perm = rng.permutation(n_glob * T)
    n_mb = (len(perm) + B - 1) // B
    flat = adv.T.reshape(-1)  # env-major
    stats = torch.zeros((n_mb, 3), dtype=torch.float64)
    local_sets = []
    for mb in range(n_mb):
        ids = sharding.split_global_minibatch(perm[mb * B:(mb + 1) * B], rank, n_local, T)
        local_sets.append(ids)
        vals = adv[:, rank * n_local:(rank + 1) * n_local].T.reshape(-1)[ids]
        stats[mb] = torch.tensor([vals.sum(), (vals ** 2).sum(), len(vals)])
    g, share = sharding.allreduce_adv_stats(stats)
    mean, std = sharding.mean_std_from_stats(g)
    ok = True
    for mb in range(n_mb):
        ref = torch.as_tensor(flat[perm[mb * B:(mb + 1) * B]])
        ok &= abs(float(mean[mb]) - float(ref.mean())) < 1e-12
        if len(ref) > 1:
            ok &= abs(float(std[mb]) - float(ref.std())) < 1e-12
        ok &= int(g[mb, 2]) == len(ref)
    tot = share.clone()
    dist.all_reduce(tot)
    ok &= bool(torch.allclose(tot, torch.ones_like(tot)))
    # every global sample lands on exactly one rank
    cnt = torch.tensor([sum(len(s) for s in local_sets)], dtype=torch.int64)
    dist.all_reduce(cnt)
    ok &= int(cnt) == n_glob * T
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_seed_streams_are_world_size_independent():
    a_i, a_g, a_e = seeding.vec_env_streams(5, 8, first_rank=0)
    b_i, b_g, b_e = seeding.vec_env_streams(5, 4, first_rank=4)
    np.testing.assert_array_equal(a_i[4:], b_i)
    np.testing.assert_array_equal(a_g[4:], b_g)
    np.testing.assert_array_equal(a_e[4:], b_e)
    # env i's goal stream is env i+1's init stream (same PCG64(seed + i + 1)), as in the reference
    np.testing.assert_array_equal(a_g[:-1], a_i[1:])
    np.testing.assert_array_equal(sharding.global_env_ids(1, 4), [4, 5, 6, 7])
