"""KAT-4: the only real-stack evidence that touches the integrators.  The shipped zips carry Monitor's
last 100 training episodes of the reference's own runs (``data:ep_info_buffer``: SB3 2.0.0 +
MuJoCo 2.1.0; tests/golden/kat4_ep_info.json).  Those episodes were produced by the (almost) final
stochastic policy, i.e. by the shipped weights and log_std under training conditions
(terminate_on_goal, 1000-step limit, goal re-drawn on reach).  Running that policy on our environments
must reproduce the distribution of episode lengths and returns: a wrong integrator form (e.g. explicit
instead of implicit joint damping changes the point robot's speed by O(h d / m) per substep) or a wrong
contact model (car) moves the mean steps-to-goal.  Weak (100 reference episodes) but real.

Returns: the shipped runs' episode returns average 2.67 (point) / 2.63 (car) where the current
wrapper.py:137-154 gives 6.65 / 6.63 on the same policy -- a difference of 3.98 / 4.00, i.e. the shipped
policies were trained when reaching a goal paid +1 instead of today's +5.  The progress part of the
return (distance covered) is what can be compared: our return - 4 per reached goal against the stored one.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _episodes_gpu(zip_path, env_name, n_envs, n_steps, rollouts):
    from mobrob_b200 import GpuVecEnv
    from mobrob_b200.ppo import EP_RING, PPO

    env = GpuVecEnv(env_name, n_envs, seed=0, time_limit=1000, terminate_on_goal=True)
    model = PPO.load(zip_path, env=env, n_steps=n_steps)
    for _ in range(rollouts):
        model.collect_rollouts()          # stochastic actions (Philox noise), log_std from the zip; no update
    torch.cuda.synchronize()
    count = int(model.ep_count.item())
    assert 0 < count <= EP_RING
    return model.ep_l[:count].cpu().numpy().astype(np.int64), model.ep_r[:count].cpu().numpy()


BONUS_NOW_MINUS_THEN = 4.0   # reach bonus 5.0 (wrapper.py:151-152) against the 1.0 the stored returns imply


def _report(name, l, r, ref):
    ref_l, ref_r = np.asarray(ref["l"], float), np.asarray(ref["r"], float)
    r = r - BONUS_NOW_MINUS_THEN * (l < 1000)
    se_l, se_r = ref_l.std(ddof=1) / np.sqrt(len(ref_l)), ref_r.std(ddof=1) / np.sqrt(len(ref_r))
    z_l, z_r = (l.mean() - ref_l.mean()) / se_l, (r.mean() - ref_r.mean()) / se_r
    print(f"\nKAT-4 {name}: episode length mean {l.mean():.2f} (n={len(l)}) vs reference run {ref_l.mean():.2f} +- {se_l:.2f} "
          f"(n={len(ref_l)}), z = {z_l:+.2f}; median {np.median(l):.0f} vs {np.median(ref_l):.0f}; "
          f"95th percentile {np.percentile(l, 95):.0f} vs {np.percentile(ref_l, 95):.0f}; max {l.max()} vs {ref_l.max():.0f}; "
          f"reached {np.mean(l < 1000):.4f} vs {np.mean(ref_l < 1000):.2f}; return (minus 4 per reached goal) mean {r.mean():.3f} vs {ref_r.mean():.3f} "
          f"+- {se_r:.3f}, z = {z_r:+.2f}")
    return z_l, z_r


def test_kat4_point_episode_statistics(cuda_lib, golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "kat4_ep_info.json")))["point"]
    l, r = _episodes_gpu(os.path.join(golden_dir, "policies", "point-ppo.zip"), "point", 2048, 1000, 2)
    z_l, z_r = _report("point (CUDA)", l, r, ref)
    assert abs(z_l) < POINT_Z and abs(z_r) < POINT_Z
    assert np.mean(l < 1000) > 0.995                      # 100 / 100 reached in the reference run
    assert abs(np.median(l) - np.median(ref["l"])) < 15
    # the CPU oracle under the same conditions (torch noise): same distribution
    from oracle import point_oracle as po, sb3_oracle
    from oracle.vec_oracle import GoalVecOracle

    pol = sb3_oracle.MlpPolicyOracle(14).load_numpy(dict(np.load(os.path.join(golden_dir, "point_policy.npz"))))
    venv = GoalVecOracle(po.PointBody(256), seed=0, time_limit=1000, terminate_on_goal=True)
    ro = sb3_oracle.RolloutOracle(venv, pol, 1500)
    ro.collect(torch.randn((1500, 256, 2), generator=torch.Generator().manual_seed(0)).numpy())
    lo = np.array([e[1] for e in ro.ep_infos]); rr = np.array([e[0] for e in ro.ep_infos])
    _report("point (oracle)", lo, rr, ref)
    assert abs(lo.mean() - l.mean()) < 4 * lo.std() / np.sqrt(len(lo)) + 0.02 * l.mean()


def test_kat4_car_episode_statistics(cuda_lib, golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "kat4_ep_info.json")))["car"]
    l, r = _episodes_gpu(os.path.join(golden_dir, "policies", "car-ppo.zip"), "car", 1024, 500, 2)
    z_l, z_r = _report("car (CUDA, soft-contact stand-in)", l, r, ref)
    assert abs(z_l) < CAR_Z and abs(z_r) < CAR_Z
    assert np.mean(l < 1000) > 0.99
    assert abs(l.mean() / np.mean(ref["l"]) - 1.0) < CAR_REL   # bounds the contact model's effect on steps-to-goal


POINT_Z = 4.0
CAR_Z = 4.0
CAR_REL = 0.15
