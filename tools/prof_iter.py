"""One profiled PPO iteration (bench workload, n_epochs configurable) for ncu:
  ncu --profile-from-start off ... python tools/prof_iter.py [n_epochs] [standalone_env_n] [point|car]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from mobrob_b200 import GpuVecEnv
from mobrob_b200.rl_control.ppo import PPOCtrl

n_epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
env_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
env_name = sys.argv[3] if len(sys.argv) > 3 else "point"
n_envs, n_steps = bench.DEFAULTS[env_name]
cfg = dict(env_name=env_name, time_limit=1000, n_envs=n_envs, vec_env_type="dummy", enable_gui=False, seed=0,
           ppo_kwargs=dict(policy="MlpPolicy", n_steps=n_steps, n_epochs=n_epochs, ent_coef=0.05,
                           gae_lambda=0.5, batch_size=bench.BATCH, verbose=0, permutation="device"))
ctrl = PPOCtrl.from_config(cfg)
model = ctrl.ppo
for _ in range(3):
    model.collect_rollouts()
    model.train()
env = None
if env_n:
    env = GpuVecEnv("point", env_n, seed=0 if env_n <= 65536 else None, time_limit=1000, terminate_on_goal=True)
    env.reset_tensor()
    act = (torch.rand((env_n, 2), device="cuda") * 2 - 1).sign().contiguous()
    for _ in range(3):
        env.step_tensor(act)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model.collect_rollouts()
model.train()
if env is not None:
    env.step_tensor(act)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled iteration done")
