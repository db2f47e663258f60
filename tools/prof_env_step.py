"""Stand-alone env-step kernel at n envs for ncu:
  ncu --profile-from-start off -k regex:point_step --set full ... python tools/prof_env_step.py [n_envs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from mobrob_b200 import GpuVecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
env = GpuVecEnv("point", n, seed=0 if n <= 65536 else None, time_limit=1000, terminate_on_goal=True)
env.reset_tensor()
act = (torch.rand((n, 2), device="cuda") * 2 - 1).sign().contiguous()
for _ in range(3):
    env.step_tensor(act)
torch.cuda.synchronize()
torch.cuda.profiler.start()
env.step_tensor(act)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled env step done")
