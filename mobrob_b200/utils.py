"""DATA_DIR / load_policy (src/mobrob/utils.py:11-16)."""
from __future__ import annotations

import os
from os.path import abspath, dirname

DATA_DIR = os.environ.get("MOBROB_DATA_DIR", os.path.join(dirname(dirname(abspath(__file__))), "data"))
PROJ_DIR = dirname(abspath(__file__))


def load_policy(env_name: str, policy_name: str):
    from .ppo import PPO

    return PPO.load(f"{DATA_DIR}/policies/{env_name}-{policy_name}.zip")
