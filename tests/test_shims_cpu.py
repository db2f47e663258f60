"""shims/: the import surface of the reference's examples (SURVEY.md 8b "Import surface") resolves to the
native implementation without stable_baselines3 / gymnasium / mobrob installed."""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the import statements of examples/train.py:6-10 and examples/control.py:5-8, as (module, names)
SURFACE = [("stable_baselines3", ["PPO"]),
           ("stable_baselines3.common.callbacks", ["CheckpointCallback"]),
           ("mobrob.rl_control.ppo", ["PPOCtrl"]),
           ("mobrob.utils", ["DATA_DIR", "BulletVideoRecorder", "load_policy"]),
           ("gymnasium.wrappers", ["RecordVideo", "TimeLimit"]),
           ("mobrob", ["get_env", "load_policy"])]


def test_import_surface_resolves_to_native_classes():
    code = ["import sys", f"sys.path[:0] = [{ROOT!r}, {os.path.join(ROOT, 'shims')!r}]"]
    for mod, names in SURFACE:
        code.append(f"from {mod} import {', '.join(names)}")
    code += ["import mobrob_b200.ppo, mobrob_b200.callbacks, mobrob_b200.rl_control.ppo",
             "assert PPO is mobrob_b200.ppo.PPO",
             "assert CheckpointCallback is mobrob_b200.callbacks.CheckpointCallback",
             "assert PPOCtrl is mobrob_b200.rl_control.ppo.PPOCtrl",
             "print('ok')"]
    out = subprocess.run([sys.executable, "-c", "\n".join(code)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_checkpoint_callback_replays_step_events(tmp_path):
    sys.path.insert(0, ROOT)
    cb_mod = importlib.import_module("mobrob_b200.callbacks")

    class FakeModel:
        n_envs, num_timesteps, saved = 2, 0, []

        def _world_size(self):
            return 1

        def _is_rank0(self):
            return True

        def save(self, path):
            self.saved.append(os.path.basename(path))

    m = FakeModel()
    cb = cb_mod.CheckpointCallback(save_freq=5, save_path=str(tmp_path), name_prefix="timestep")
    cb.init_callback(m)
    cb.on_training_start({}, {})
    for _ in range(3):                 # three rollouts of 4 vec-steps
        m.num_timesteps += 4 * m.n_envs
        assert cb.on_rollout_steps(4) is True
    # SB3 saves at n_calls = 5 and 10, i.e. num_timesteps = 10 and 20
    assert m.saved == ["timestep_10_steps.zip", "timestep_20_steps.zip"]
    assert cb.n_calls == 12
