"""Where does the host time of PPOCtrl.learn() go?  (e2e arm of bench.py)
  python tools/e2e_probe.py [iterations]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from mobrob_b200 import permfeed
from mobrob_b200.rl_control.ppo import PPOCtrl

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(env_name="point", time_limit=1000, n_envs=bench.N_ENVS, vec_env_type="dummy", enable_gui=False, seed=0,
           ppo_kwargs=dict(policy="MlpPolicy", n_steps=bench.N_STEPS, n_epochs=bench.N_EPOCHS, ent_coef=0.05,
                           gae_lambda=0.5, batch_size=bench.BATCH, verbose=0, permutation="pool"))
ctrl = PPOCtrl.from_config(cfg)
model = ctrl.ppo
model.tensorboard_log = None
T = {}


def timed(obj, name, key=None):
    fn = getattr(obj, name)
    key = key or name

    def wrap(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        T.setdefault(key, []).append(time.perf_counter() - t0)
        return r

    setattr(obj, name, wrap)


draw = permfeed.PermutationFeeder._draw


def draw_timed(self, slot, it, e):
    t0 = time.perf_counter()
    draw(self, slot, it, e)
    T.setdefault("draw(worker)", []).append(time.perf_counter() - t0)


permfeed.PermutationFeeder._draw = draw_timed
f = model._get_feeder()
for n in ("stage", "prefetch"):
    timed(f, n, "feeder." + n)
for n in ("collect_rollouts", "train", "_drain_episodes", "_log_train"):
    timed(model, n)
timed(model.logger, "dump", "logger.dump")
steps = bench.N_ENVS * bench.N_STEPS
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model.learn(total_timesteps=steps * warm)
T.clear()
t0 = time.perf_counter()
model.learn(total_timesteps=steps * iters, reset_num_timesteps=True)
wall = time.perf_counter() - t0
print(f"{iters} iterations: {1e3 * wall / iters:.1f} ms / iteration, {steps * iters / wall:.3e} env-steps/s, "
      f"{os.cpu_count()} cpus, {f.pool._max_workers} workers")
for k, v in sorted(T.items(), key=lambda kv: -sum(kv[1])):
    print(f"  {k:22s} n={len(v):4d}  mean {1e3 * np.mean(v):8.2f} ms  max {1e3 * np.max(v):8.2f} ms  total/iter {1e3 * sum(v) / iters:8.2f} ms")
