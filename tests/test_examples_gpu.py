"""The reference's example scripts, UNCHANGED (byte copies under tests/golden/examples, written by
tests/golden/make_golden.py from /root/reference/examples/{train,control}.py), run end to end through
shims/ on the CUDA path: north_star "examples/train.py and examples/control.py run unchanged and the
pretrained data/policies load directly".  Only the data directory is redirected (MOBROB_DATA_DIR) so
that the runs are small and write into a scratch directory.
"""
import os
import re
import shutil
import subprocess
import sys

import pytest
import yaml

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = os.path.join(ROOT, "tests", "golden", "examples")


def _run(script, args, data_dir):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "shims"), ROOT, env.get("PYTHONPATH", "")])
    env["MOBROB_DATA_DIR"] = str(data_dir)
    return subprocess.run([sys.executable, os.path.join(EXAMPLES, script), *args], env=env, capture_output=True,
                          text=True, timeout=900, cwd=str(data_dir))


def _config(data_dir, env_name, total_timesteps):
    """data/configs/{env}-ppo.yaml of the reference (same keys; n_steps / batch / budget scaled down)."""
    cfg = dict(env_name=env_name, time_limit=1000, n_envs=8, vec_env_type="subproc", enable_gui=False, seed=0,
               ppo_kwargs=dict(policy="MlpPolicy", n_steps=64, n_epochs=3, ent_coef=0.05, gae_lambda=0.5,
                               batch_size=128, verbose=1, device="cpu"),
               total_timesteps=total_timesteps)
    os.makedirs(os.path.join(data_dir, "configs"), exist_ok=True)
    with open(os.path.join(data_dir, "configs", f"{env_name}-ppo.yaml"), "w") as f:
        yaml.safe_dump(cfg, f)


@pytest.mark.parametrize("env_name", ["point", "car"])
def test_train_script_runs_unchanged(cuda_lib, tmp_path, golden_dir, env_name):
    """examples/train.py:16-49: from_config -> CheckpointCallback(save_freq // n_envs) -> learn(progress_bar) ->
    save_model; then --finetune, which loads the shipped zip's state dict into the fresh policy first."""
    from mobrob_b200.ppo import PPO

    _config(tmp_path, env_name, total_timesteps=2048)
    out = _run("train.py", ["--env-name", env_name, "--save-freq", "1024"], tmp_path)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    models = tmp_path / "policies" / "tmp" / f"{env_name}-ppo" / "models"
    # CheckpointCallback names: {name_prefix}_{num_timesteps}_steps.zip every save_freq // n_envs vec-steps
    assert sorted(os.listdir(models)) == ["timestep_1024_steps.zip", "timestep_2048_steps.zip"]
    final = tmp_path / "policies" / f"{env_name}-ppo.zip"
    assert final.exists()
    m = PPO.load(str(final))
    assert m.num_timesteps == 2048 and m.n_steps == 64 and m._n_updates == 4 * 3
    ck = PPO.load(str(models / "timestep_1024_steps.zip"))
    assert ck.num_timesteps == 1024
    # SB3's log table (verbose: 1) with the reference run's keys
    # (rollout/ep_* only appear once an episode has ended: not guaranteed in 2048 steps of an untrained policy)
    for key in ("time/", "fps", "total_timesteps", "train/", "approx_kl", "clip_fraction",
                "entropy_loss", "explained_variance", "learning_rate", "policy_gradient_loss", "value_loss", "n_updates"):
        assert key in out.stdout, key
    # tensorboard events under data/policies/tmp/{env}-ppo/tensorboard (src/mobrob/rl_control/ppo.py:53-57)
    tb = tmp_path / "policies" / "tmp" / f"{env_name}-ppo" / "tensorboard"
    assert any(f.startswith("events.out.tfevents") for _, _, fs in os.walk(tb) for f in fs)

    # --finetune: the shipped policy is the starting point
    shutil.copyfile(os.path.join(golden_dir, "policies", f"{env_name}-ppo.zip"), final)
    _config(tmp_path, env_name, total_timesteps=512)
    out = _run("train.py", ["--env-name", env_name, "--finetune", "--save-freq", "100000"], tmp_path)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    ft = PPO.load(str(final))
    shipped = PPO.load(os.path.join(golden_dir, "policies", f"{env_name}-ppo.zip"))
    a, b = ft.policy.state_dict(), shipped.policy.state_dict()
    moved = max(float((a[k] - b[k]).abs().max()) for k in a)
    assert 0.0 < moved < 0.05, moved   # one iteration of 3 epochs x 4 minibatches at lr 3e-4 away from the shipped weights


def test_control_script_runs_unchanged(cuda_lib, tmp_path, golden_dir):
    """examples/control.py:11-63 --no-gui: get_env + load_policy + 1000 deterministic steps per epoch."""
    os.makedirs(tmp_path / "policies")
    shutil.copyfile(os.path.join(golden_dir, "policies", "point-ppo.zip"), tmp_path / "policies" / "point-ppo.zip")
    out = _run("control.py", ["--env-name", "point", "--no-gui", "--epochs", "2"], tmp_path)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    m = re.search(r"average reward: ([-0-9.e+]+)", out.stdout)
    assert m, out.stdout[-2000:]
    # the shipped policy reaches a goal every ~80-120 steps: >= 5 goals x (+5 bonus + progress) per 1000 steps
    assert float(m.group(1)) > 25.0, out.stdout[-500:]
    assert "rewards: [" in out.stdout


def test_fix_pickle_warning_script_runs_unchanged(cuda_lib, tmp_path, golden_dir):
    """examples/fix_pickle_warning.py:1-20: load_policy + save over all five robots.  point / car go through the CUDA
    PPO object; doggo, drone and turtlebot3 (observations 58 / 12 / 43, actions 12 / 18 / 2: outside the path) come
    back as archive handles that re-save what they hold and refuse to run.  The three out-of-scope zips here are
    stand-ins with the shipped shapes, built from the point zip's entries."""
    import io
    import json
    import zipfile

    import torch

    from mobrob_b200.ppo import PPO, StoredPolicy

    pol = tmp_path / "policies"
    os.makedirs(pol)
    for name in ("point", "car"):
        shutil.copyfile(os.path.join(golden_dir, "policies", f"{name}-ppo.zip"), pol / f"{name}-ppo.zip")
    with zipfile.ZipFile(os.path.join(golden_dir, "policies", "point-ppo.zip")) as z:
        data = json.loads(z.read("data"))
    for name, (o, a) in {"doggo": (58, 12), "drone": (12, 18), "turtlebot3": (43, 2)}.items():
        d = json.loads(json.dumps(data))
        d["observation_space"]["_shape"] = [o]
        d["action_space"]["_shape"] = [a]
        sd = {"log_std": torch.zeros(a), "mlp_extractor.policy_net.0.weight": torch.randn(64, o), "action_net.weight": torch.randn(a, 64)}
        bio = io.BytesIO()
        torch.save(sd, bio)
        with zipfile.ZipFile(pol / f"{name}-ppo.zip", "w") as z:
            z.writestr("data", json.dumps(d))
            z.writestr("policy.pth", bio.getvalue())
            z.writestr("_stable_baselines3_version", "2.0.0")
    out = _run("fix_pickle_warning.py", [], tmp_path)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    for name in ("point", "car"):
        m = PPO.load(str(pol / f"{name}-ppo.zip"))
        ref = PPO.load(os.path.join(golden_dir, "policies", f"{name}-ppo.zip"))
        assert isinstance(m, PPO) and m.num_timesteps == ref.num_timesteps
        for k, v in ref.policy.state_dict().items():
            assert torch.equal(v, m.policy.state_dict()[k]), k
    for name, o in (("doggo", 58), ("drone", 12), ("turtlebot3", 43)):
        m = PPO.load(str(pol / f"{name}-ppo.zip"))
        assert isinstance(m, StoredPolicy) and m.obs_shape == (o,)
        assert sorted(zipfile.ZipFile(pol / f"{name}-ppo.zip").namelist()) == sorted(
            ["data", "pytorch_variables.pth", "policy.pth", "policy.optimizer.pth", "_stable_baselines3_version", "system_info.txt"])
        assert m.state_dict["mlp_extractor.policy_net.0.weight"].shape == (64, o)
        with pytest.raises(NotImplementedError):
            m.predict(None)
