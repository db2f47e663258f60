"""BASELINE.json configs[3]: car env (26-dim obs, wheel-hinge dynamics, soft contacts), 16384 envs, rollout + PPO update
on one B200.  n_steps = 74 -> 1 212 416 samples per iteration = 64 minibatches of 18 944, 10 epochs (the point bench's
batch structure)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from mobrob_b200.rl_control.ppo import PPOCtrl

n_envs, n_steps, batch = 16384, 74, 18944
cfg = dict(env_name="car", time_limit=1000, n_envs=n_envs, vec_env_type="dummy", enable_gui=False, seed=0,
           ppo_kwargs=dict(policy="MlpPolicy", n_steps=n_steps, n_epochs=10, ent_coef=0.05, gae_lambda=0.5,
                           batch_size=batch, verbose=0, permutation="device"))
model = PPOCtrl.from_config(cfg).ppo
model.tensorboard_log = None
for _ in range(2):
    model.collect_rollouts()
    model.train()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
K = 3
ro = up = 0.0
for _ in range(K):
    ev[0].record()
    model.collect_rollouts()
    ev[1].record()
    model.train()
    ev[2].record()
    torch.cuda.synchronize()
    ro += ev[0].elapsed_time(ev[1])
    up += ev[1].elapsed_time(ev[2])
ms = (ro + up) / K
print(json.dumps({"workload": f"car {n_envs} envs x {n_steps} steps, 10 epochs x 64 minibatches of {batch}",
                  "rollout_ms": ro / K, "update_ms": up / K, "env_steps_per_s": n_envs * n_steps / (ms * 1e-3)}))
