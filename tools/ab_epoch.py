"""Time the fused epoch kernel alone on the bench workload (CUDA events, parameters restored).

  python tools/ab_epoch.py [point|car]        # MR_PPO_EPOCH=v1 selects the round-1 kernel
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from mobrob_b200.rl_control.ppo import PPOCtrl

env_name = sys.argv[1] if len(sys.argv) > 1 else "point"
n_envs, n_steps = (bench.N_ENVS, bench.N_STEPS) if env_name == "point" else (16384, 74)
cfg = dict(env_name=env_name, time_limit=1000, n_envs=n_envs, vec_env_type="dummy", enable_gui=False, seed=0,
           ppo_kwargs=dict(policy="MlpPolicy", n_steps=n_steps, n_epochs=1, ent_coef=0.05, gae_lambda=0.5,
                           batch_size=bench.BATCH, verbose=0, permutation="device"))
model = PPOCtrl.from_config(cfg).ppo
dev = model.device
for _ in range(2):
    model.collect_rollouts()
    model.train()
torch.cuda.synchronize()
ms = [bench.time_update_kernel(model, dev)[0] / model.n_epochs for _ in range(3)]
print(f"AB_EPOCH env={env_name} kernel={os.environ.get('MR_PPO_EPOCH', 'v2')} ms_per_epoch={min(ms):.4f} all={ms}")
