"""The [SB3 2.0.0] callback surface mobrob uses: ``CheckpointCallback`` (examples/train.py:36-41).

SB3 calls ``callback.on_step()`` after every ``VecEnv.step``; here a rollout is one kernel launch,
so ``PPO.learn`` reports whole rollouts with ``on_rollout_steps(n)`` and the callback replays the
n step events it would have seen.  A checkpoint that SB3 would write in the middle of a rollout
holds the same policy (parameters only change in ``train``) and gets the same file name
(``{name_prefix}_{num_timesteps}_steps.zip`` with the step's own ``num_timesteps``).
"""
from __future__ import annotations

import os


class BaseCallback:
    def __init__(self, verbose: int = 0):
        self.verbose = verbose
        self.model = None
        self.n_calls = 0
        self.num_timesteps = 0
        self.locals, self.globals = {}, {}
        self.parent = None

    # -- SB3 protocol ---------------------------------------------------------------------------
    def init_callback(self, model) -> None:
        self.model = model
        self._init_callback()

    def _init_callback(self) -> None:
        pass

    def on_training_start(self, locals_, globals_) -> None:
        self.locals, self.globals = locals_, globals_
        self.num_timesteps = self.model.num_timesteps
        self._on_training_start()

    def _on_training_start(self) -> None:
        pass

    def on_rollout_start(self) -> None:
        pass

    def on_step(self) -> bool:
        self.n_calls += 1
        self.num_timesteps = self.model.num_timesteps
        return self._on_step()

    def _on_step(self) -> bool:
        return True

    def on_rollout_end(self) -> None:
        pass

    def on_training_end(self) -> None:
        pass

    # -- batched form used by mobrob_b200.PPO.learn ---------------------------------------------------
    def on_rollout_steps(self, n_steps: int) -> bool:
        """n_steps VecEnv steps have just been taken (model.num_timesteps already counts them)."""
        end = self.model.num_timesteps
        per_step = self.model.n_envs * self.model._world_size()
        keep_going = True
        for k in range(n_steps):
            self.n_calls += 1
            self.num_timesteps = end - (n_steps - 1 - k) * per_step
            if self._on_step() is False:
                keep_going = False
        return keep_going


class CheckpointCallback(BaseCallback):
    """Save the model every ``save_freq`` calls of ``env.step()`` (per-env steps), like SB3's."""

    def __init__(self, save_freq: int, save_path: str, name_prefix: str = "rl_model",
                 save_replay_buffer: bool = False, save_vecnormalize: bool = False, verbose: int = 0):
        super().__init__(verbose)
        self.save_freq = max(int(save_freq), 1)
        self.save_path = save_path
        self.name_prefix = name_prefix

    def _init_callback(self) -> None:
        if self.save_path is not None:
            os.makedirs(self.save_path, exist_ok=True)

    def _checkpoint_path(self, checkpoint_type: str = "", extension: str = "") -> str:
        return os.path.join(self.save_path, f"{self.name_prefix}_{checkpoint_type}{self.num_timesteps}_steps.{extension}")

    def _on_step(self) -> bool:
        if self.n_calls % self.save_freq == 0:
            path = self._checkpoint_path(extension="zip")
            if self.model._is_rank0():
                self.model.save(path)
                if self.verbose >= 2:
                    print(f"Saving model checkpoint to {path}")
        return True
