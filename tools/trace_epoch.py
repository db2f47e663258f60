"""Phase timeline of the persistent epoch kernel (debug build with -DMR_TRACE).

  python tools/trace_epoch.py build        # here: nvcc -DMR_TRACE -> mobrob_b200/lib/libmobrob_b200_trace.so
  python tools/trace_epoch.py [n_epochs]   # on the GPU box: run, print per-phase durations of CTA 0 / 1
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mobrob_b200 import build as B

TRACE_LIB = os.path.join(B.LIB_DIR, "libmobrob_b200_trace.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    print(B.build(lib_path=TRACE_LIB, defines=("-DMR_TRACE",)))
    sys.exit(0)

os.environ["MR_LIB_PATH"] = TRACE_LIB
import numpy as np
import torch

import bench
from mobrob_b200 import _lib
from mobrob_b200.rl_control.ppo import PPOCtrl

world = int(os.environ.get("WORLD_SIZE", "1"))   # under torchrun: the exchange phases ("slice exchanged", "barrier B passed")
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
cfg = dict(env_name="point", time_limit=1000, n_envs=bench.N_ENVS, vec_env_type="dummy", enable_gui=False, seed=0,
           ppo_kwargs=dict(policy="MlpPolicy", n_steps=bench.N_STEPS, n_epochs=1, ent_coef=0.05,
                           gae_lambda=0.5, batch_size=bench.BATCH, verbose=0, permutation="device"))
model = PPOCtrl.from_config(cfg).ppo
lib = _lib.load()
lib.mr_trace_read.restype = ctypes.c_int
lib.mr_trace_read.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(4096, dtype=np.uint64)
for _ in range(3):
    model.collect_rollouts()
    model.train()
torch.cuda.synchronize()
for cta in (0, 1):
    lib.mr_trace_read(cta, buf.ctypes.data, 4096)   # clear
model.collect_rollouts()
model.train()
torch.cuda.synchronize()
NAMES = {1: "mb start", 2: "mb start", 3: "partial reduced", 4: "barrier A passed", 5: "slice exchanged", 6: "barrier B passed",
         7: "adam done", 8: "barrier C passed", 22: "partial staged", 10: "tile start", 11: "gathered+sync", 12: "Z1 done", 13: "H1 stored+sync",
         14: "Z2 done", 15: "heads/dZ2 stored+sync", 16: "dH done", 17: "dZ1 computed", 18: "dW2 done",
         19: "dZ1 stored+sync", 20: "tiles issued", 21: "dW1 done", 30: "gradient loaded", 31: "norm known",
         32: "slice summed", 33: "norm known", 34: "adam ctx", 35: "adam pass 1", 36: "adam pass 1 synced",
         37: "restaged"}
for cta in (0, 1) if rank == 0 else ():
    n = lib.mr_trace_read(cta, buf.ctypes.data, 4096)
    ids = (buf[:n] >> np.uint64(56)).astype(int)
    ts = (buf[:n] & np.uint64((1 << 56) - 1)).astype(np.int64)
    print(f"== CTA {cta}: {n} marks")
    # first 2 minibatches verbatim, then per-phase averages
    starts = [i for i in range(n) if ids[i] == 2]
    for i in range(starts[1] if len(starts) > 1 else 0, starts[3] if len(starts) > 3 else n):
        print(f"   {NAMES.get(ids[i], ids[i]):28s} +{(ts[i] - ts[i - 1]) if i else 0:7d} ns")
    agg = {}
    for i in range(1, n):
        agg.setdefault((ids[i - 1], ids[i]), []).append(ts[i] - ts[i - 1])
    print("   -- mean per transition (ns), count")
    for (a, b), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"   {NAMES.get(a, a):26s} -> {NAMES.get(b, b):26s} {np.mean(v):9.0f} x{len(v):4d}  total {sum(v) / 1e3:8.1f} us")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
