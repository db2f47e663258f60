"""Stand-alone env-step kernel at n envs for ncu:
  ncu --profile-from-start off -k regex:point_step --set full ... python tools/prof_env_step.py [n_envs] [point|car]

Also prints the kernel's CUDA-event time (20 launches) before the profiled launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from mobrob_b200 import GpuVecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
name = sys.argv[2] if len(sys.argv) > 2 else "point"
env = GpuVecEnv(name, n, seed=0 if n <= 65536 else None, time_limit=1000, terminate_on_goal=True)
env.reset_tensor()
act = (torch.rand((n, 2), device="cuda") * 2 - 1).sign().contiguous()
for _ in range(30 if name == "car" else 3):   # the car settles onto its wheels first (it is dropped from z = 0.1)
    env.step_tensor(act)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    env.step_tensor(act)
e1.record()
torch.cuda.synchronize()
print(f"STEP_TIME env={name} n={n}: {e0.elapsed_time(e1) / 20:.4f} ms per env-step launch "
      f"({n / (e0.elapsed_time(e1) / 20 * 1e-3):.3e} env-steps/s)")
torch.cuda.profiler.start()
env.step_tensor(act)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled env step done")
