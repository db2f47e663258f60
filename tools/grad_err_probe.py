"""Print the gradient errors of the tensor-core PPO kernels against torch-CPU autograd on the
problems of tests/test_ppo_update_gpu.py (norm-wise, relative to the largest gradient entry)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import test_ppo_update_gpu as T
from oracle import sb3_oracle

golden = os.path.join(ROOT, "tests", "golden")
for (O, TT, N, B, pre) in [(14, 16, 40, 100, False), (14, 64, 37, 999, True), (26, 32, 24, 500, False),
                           (14, 296, 64, 18944, False), (14, 296, 64, 18944, True)]:
    pol, buf = T._make_problem(O, TT, N, seed=O + TT, pretrained_dir=golden if pre else None)
    kw = dict(clip_range=0.2, ent_coef=0.05, vf_coef=0.5, normalize_advantage=True)
    up = T._updater(pol, O, **kw)
    dbuf = {k: torch.as_tensor(v).cuda().contiguous() for k, v in buf.items()}
    perm = np.random.default_rng(1).permutation(N * TT).astype(np.int64)
    dperm = torch.as_tensor(perm).cuda()
    stats = up.adv_stats(dbuf["advantages"], dperm, B, N, TT)
    idx = perm[:B]
    loss, st = sb3_oracle.ppo_loss(pol, *T._oracle_batch(buf, idx), **kw)
    pol.zero_grad()
    loss.backward()
    g_ref = pol.flat_grads().numpy()
    g = up.compute_grad(dbuf, dperm[:B], stats[0], N, TT).cpu().numpy()[:up.n_params]
    print(f"O={O} T={TT} N={N} B={B} pretrained={pre}: max|g - g_ref| / max|g_ref| = "
          f"{np.abs(g - g_ref).max() / np.abs(g_ref).max():.2e}")
    off, parts = 0, []
    for name in sb3_oracle.PARAM_ORDER:   # every tensor against its own largest entry
        n = dict(pol.named_parameters())[name].numel()
        r = g_ref[off:off + n]
        parts.append(f"{name.replace('mlp_extractor.', '')} {np.abs(g[off:off + n] - r).max() / max(np.abs(r).max(), 1e-30):.1e}")
        off += n
    print("      per tensor: " + ", ".join(parts))
