"""Small end-to-end pass for compute-sanitizer (memcheck): every kernel family once at tiny sizes.
  compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from mobrob_b200 import GpuVecEnv
from mobrob_b200.ppo import PPO

flags = {"observe_goal_dist": True, "observe_qpos": True, "observe_qvel": True, "observe_ctrl": True}
for name, n, T, cfg in (("point", 300, 12, None), ("point", 130, 10, flags), ("car", 70, 6, None), ("car", 40, 4, {"observe_ctrl": True})):
    env = GpuVecEnv(name, n, seed=1, time_limit=7, terminate_on_goal=True, robot_config=cfg)
    env.reset()
    for _ in range(9):
        env.step(np.sign(np.random.default_rng(0).standard_normal((n, 2))).astype(np.float32))
    env.get_state(); env.get_pos(); env.get_obs_tensor()
    # whole minibatches (one launch for the update) and a ragged last minibatch (one launch per epoch)
    for batch in (n * T // 2 // 128 * 128 or 128, 256):
        model = PPO("MlpPolicy", env, n_steps=T, batch_size=batch, n_epochs=2, seed=1, ent_coef=0.05, gae_lambda=0.5)
        model.learn(total_timesteps=2 * n * T)
        model.predict(env.reset(), deterministic=True)
    torch.cuda.synchronize()
    print("ok", name, n, T, cfg is not None, flush=True)
print("sanitize pass done")
