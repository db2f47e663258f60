"""Time the fused rollout kernel (bench workload) -- MR_ROLLOUT_CFG=<warps>x<envs per warp> selects a variant."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from mobrob_b200.rl_control.ppo import PPOCtrl

cfg = dict(env_name="point", time_limit=1000, n_envs=bench.N_ENVS, vec_env_type="dummy", enable_gui=False, seed=0,
           ppo_kwargs=dict(policy="MlpPolicy", n_steps=bench.N_STEPS, n_epochs=1, ent_coef=0.05,
                           gae_lambda=0.5, batch_size=bench.BATCH, verbose=0, permutation="device"))
model = PPOCtrl.from_config(cfg).ppo
for _ in range(3):
    model.collect_rollouts()
    model.train()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    model.collect_rollouts()
e1.record()
torch.cuda.synchronize()
print(f"MR_ROLLOUT_CFG={os.environ.get('MR_ROLLOUT_CFG', 'default')}: {e0.elapsed_time(e1) / 10:.3f} ms per rollout+GAE "
      f"({bench.N_ENVS} envs x {bench.N_STEPS} steps)")
