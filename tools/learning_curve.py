"""Learning curve of the bench workload (point, 4096 envs x 296 steps, 10 epochs x 64 minibatches)
from scratch: rollout/ep_len_mean, ep_rew_mean and train/* per iteration, one JSON line each.

  python tools/learning_curve.py [--iters 60] [--permutation device|pool|sb3] [--port]

--port runs the same configuration on the CPU oracle port (numpy env + torch-CPU PPO) instead, for the
comparison of the two curves (minutes per iteration batch: use a small --iters)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=60)
ap.add_argument("--permutation", default="device")
ap.add_argument("--envs", type=int, default=bench.N_ENVS)
ap.add_argument("--port", action="store_true")
a = ap.parse_args()

if a.port:
    import torch

    from oracle import sb3_oracle

    torch.set_num_threads(os.cpu_count() or 1)
    st = bench.cpu_port_setup(a.envs, bench.N_STEPS)
    ro = st["ro"]
    for it in range(a.iters):
        bench.cpu_port_step(a.envs, bench.N_STEPS, bench.BATCH, bench.N_EPOCHS, st)
        last = ro.ep_infos[-100:]
        print(json.dumps({"iteration": it + 1, "impl": "port", "episodes": len(ro.ep_infos),
                          "ep_len_mean": float(np.mean([l for _, l in last])) if last else None,
                          "ep_rew_mean": float(np.mean([r for r, _ in last])) if last else None}), flush=True)
    sys.exit(0)

from mobrob_b200 import ppo as ppo_mod
from mobrob_b200.rl_control.ppo import PPOCtrl

cfg = dict(env_name="point", time_limit=1000, n_envs=a.envs, vec_env_type="dummy", enable_gui=False, seed=0,
           ppo_kwargs=dict(policy="MlpPolicy", n_steps=bench.N_STEPS, n_epochs=bench.N_EPOCHS, ent_coef=0.05,
                           gae_lambda=0.5, batch_size=bench.BATCH, verbose=0, permutation=a.permutation))
ctrl = PPOCtrl(cfg["ppo_kwargs"], "point", 1000, a.envs, seed=0, tensorboard_log=False)
orig = ppo_mod.Logger.dump


def dump(self, step=0):
    d = dict(self.name_to_value)
    keep = ("rollout/ep_len_mean", "rollout/ep_rew_mean", "time/iterations", "time/fps", "train/value_loss",
            "train/approx_kl", "train/clip_fraction", "train/explained_variance", "train/std", "train/entropy_loss")
    print(json.dumps({k.split("/")[1]: d[k] for k in keep if k in d}), flush=True)
    orig(self, step)


ppo_mod.Logger.dump = dump
ctrl.learn(total_timesteps=a.iters * a.envs * bench.N_STEPS)
