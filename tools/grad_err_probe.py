import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import test_ppo_update_gpu as t
from oracle import sb3_oracle
gd = '/root/repo/tests/golden'
for (O,T,N,B,pre) in [(14,16,40,100,False),(14,64,37,999,True),(26,32,24,500,False),(14,64,64,4096,True)]:
    pol, buf = t._make_problem(O,T,N,seed=O+T,pretrained_dir=gd if pre else None)
    kw = dict(clip_range=0.2, ent_coef=0.05, vf_coef=0.5, normalize_advantage=True)
    up = t._updater(pol, O, **kw)
    dbuf = {k: torch.as_tensor(v).cuda().contiguous() for k,v in buf.items()}
    perm = np.random.default_rng(1).permutation(N*T).astype(np.int64)
    dperm = torch.as_tensor(perm).cuda()
    stats = up.adv_stats(dbuf["advantages"], dperm, B, N, T)
    idx = perm[:B]
    loss, st = sb3_oracle.ppo_loss(pol, *t._oracle_batch(buf, idx), **kw)
    pol.zero_grad(); loss.backward()
    g_ref = pol.flat_grads().numpy()
    g = up.compute_grad(dbuf, dperm[:B], stats[0], N, T).cpu().numpy()[:up.n_params]
    print(O,T,N,B, "rel err", np.abs(g-g_ref).max()/np.abs(g_ref).max())
