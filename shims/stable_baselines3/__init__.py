"""Name shim: ``from stable_baselines3 import PPO`` (examples/train.py:6, src/mobrob/utils.py:7)."""
from mobrob_b200.ppo import PPO  # noqa: F401

__version__ = "2.0.0"
