"""Name shim: ``from stable_baselines3.common.callbacks import CheckpointCallback`` (examples/train.py:7)."""
from mobrob_b200.callbacks import BaseCallback, CheckpointCallback  # noqa: F401
