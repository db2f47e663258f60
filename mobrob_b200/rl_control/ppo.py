"""PPOCtrl with the reference's constructor and methods (src/mobrob/rl_control/ppo.py:14-77).

``vec_env_type`` ("dummy" / "subproc") is accepted for config compatibility; both map onto the
single HBM-resident GpuVecEnv (there are no worker processes to choose between).  With
torch.distributed initialised, each rank owns ``n_env`` environments with global ranks
``rank * n_env ...`` and PPO all-reduces gradients.
"""
from __future__ import annotations

import os

import torch

from ..ppo import PPO
from ..utils import DATA_DIR
from ..vec_env import GpuVecEnv

try:
    import tensorboard  # noqa: F401
except ImportError:  # pragma: no cover
    tensorboard = None


class PPOCtrl:
    def __init__(self, ppo_kwargs: dict, env_name: str, time_limit: int, n_env: int,
                 vec_env_type: str = "dummy", enable_gui: bool = False, seed: int = 0,
                 tensorboard_log: bool | None = None) -> None:
        self.ppo_kwargs = dict(ppo_kwargs)
        self.env_name, self.time_limit, self.n_env = env_name, time_limit, n_env
        if vec_env_type not in ("subproc", "dummy"):
            raise ValueError(f"Unknown vec_env_type: {vec_env_type}")
        rank = 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank = torch.distributed.get_rank()
        vec_env = GpuVecEnv(env_name, n_env, seed=seed, time_limit=time_limit, terminate_on_goal=True,
                            first_rank=rank * n_env)
        kw = dict(self.ppo_kwargs)
        kw.pop("device", None)  # the reference's "cpu" has no meaning here: the path is CUDA only
        use_tb = tensorboard is not None if tensorboard_log is None else tensorboard_log
        self.ppo = PPO(env=vec_env, seed=seed,
                       tensorboard_log=(f"{DATA_DIR}/policies/tmp/{env_name}-ppo/tensorboard" if use_tb else None),
                       **kw)

    @classmethod
    def from_config(cls, config: dict) -> "PPOCtrl":
        return cls(ppo_kwargs=config["ppo_kwargs"], env_name=config["env_name"],
                   time_limit=config["time_limit"], n_env=config["n_envs"],
                   vec_env_type=config["vec_env_type"], enable_gui=config["enable_gui"],
                   seed=config["seed"])

    def learn(self, *args, **kwargs) -> None:
        self.ppo.learn(*args, **kwargs)

    def save_model(self, save_path: str) -> None:
        self.ppo.save(save_path)
