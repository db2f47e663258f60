"""Time the stand-alone env-step kernel at 2^22 envs (same routine as bench.py's roofline_env_step)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
r = bench.time_env_step_kernel(dev, bench.measured_peaks()[0])
print(os.environ.get("MR_STEP_MINB", "-"), json.dumps({k: r[k] for k in ("ms_per_launch", "frac", "achieved")}))
