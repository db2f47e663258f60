"""Name shim: ``from mobrob.rl_control.ppo import PPOCtrl`` (examples/train.py:9)."""
from mobrob_b200.rl_control.ppo import PPOCtrl  # noqa: F401
