"""Optional Engine.obs() keys (observe_goal_dist / observe_qpos / observe_qvel / observe_ctrl,
src/mobrob/envs/mujoco_robots/robots/engine.py:125,140-142,1179-1180,1243-1259) of the CUDA envs against the oracle's
restatement, through the VecEnv API, the get_env-style wrapper (get_robot_config, wrapper.py:235-240) and PPO.
"""
import itertools

import numpy as np
import pytest
import torch

from oracle import car_oracle as co, point_oracle as po
from oracle.vec_oracle import GoalVecOracle

pytestmark = pytest.mark.gpu

KEYS = ("observe_goal_dist", "observe_qpos", "observe_qvel", "observe_ctrl")
SIZES = {"point": dict(base=14, observe_goal_dist=1, observe_qpos=3, observe_qvel=3, observe_ctrl=2),
         "car": dict(base=26, observe_goal_dist=1, observe_qpos=13, observe_qvel=11, observe_ctrl=2)}


def _configs():
    yield {k: True for k in KEYS}
    for k in KEYS:
        yield {k: True}
    yield {"observe_qvel": True, "observe_ctrl": True, "observe_goal_dist": False}


@pytest.mark.parametrize("cfg", list(_configs()), ids=lambda c: "+".join(k[8:] for k, v in c.items() if v))
def test_point_optional_keys_match_oracle(cuda_lib, cfg):
    from mobrob_b200 import GpuVecEnv

    n, seed, tl = 48, 3, 40   # short time limit: truncations, goal hits, full and goal-only resets all occur
    ora = GoalVecOracle(po.PointBody(n), seed=seed, time_limit=tl, terminate_on_goal=True, observe=cfg)
    gpu = GpuVecEnv("point", n, seed=seed, time_limit=tl, terminate_on_goal=True, robot_config=cfg)
    dim = SIZES["point"]["base"] + sum(SIZES["point"][k] for k, v in cfg.items() if v)
    assert gpu.obs_dim == dim and gpu.observation_space.shape == (dim,)
    o_ref, o_gpu = ora.reset(), gpu.reset()
    assert o_gpu.shape == o_ref.shape == (n, dim)
    np.testing.assert_allclose(o_gpu, o_ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gpu.get_obs_tensor().cpu().numpy(), o_ref, rtol=1e-5, atol=1e-6)
    rng = np.random.default_rng(9)
    n_term = 0
    for t in range(200):
        a = (np.sign(rng.standard_normal((n, 2))) * rng.choice([0.4, 1.0, 1.7], (n, 2))).astype(np.float32)
        o_ref, r_ref, d_ref, info = ora.step(a)
        o_gpu, r_gpu, d_gpu, infos = gpu.step(a)
        np.testing.assert_array_equal(d_gpu, d_ref)
        np.testing.assert_allclose(o_gpu, o_ref, rtol=1e-5, atol=1e-6, err_msg=f"obs step {t}")
        np.testing.assert_allclose(r_gpu, r_ref, rtol=1e-5, atol=1e-7)
        for i in np.nonzero(d_ref)[0]:
            n_term += 1
            np.testing.assert_allclose(infos[i]["terminal_observation"], info["terminal_obs"][i], rtol=1e-5, atol=1e-6)
    assert n_term > n


def test_row_layout_is_the_sorted_key_order(cuda_lib):
    """accelerometer | ctrl | goal_compass | goal_dist | gyro | magnetometer | qpos | qvel | velocimeter: every slice of
    the extended row equals the corresponding quantity (default row, state export, goal distance)."""
    from mobrob_b200 import GpuVecEnv

    n = 16
    full = GpuVecEnv("point", n, seed=5, time_limit=None, terminate_on_goal=False, robot_config={k: True for k in KEYS})
    base = GpuVecEnv("point", n, seed=5, time_limit=None, terminate_on_goal=False)
    full.reset(); base.reset()
    rng = np.random.default_rng(0)
    for _ in range(30):
        a = rng.uniform(-1.5, 1.5, (n, 2)).astype(np.float32)
        o_f, *_ = full.step(a)
        o_b, *_ = base.step(a)
    st = full.get_state().cpu().numpy()   # qpos(3) qvel(3) body_xy(2) psi0 ctrl(2) goal(2) ...
    pos = full.get_pos().cpu().numpy()
    np.testing.assert_array_equal(o_f[:, 0:3], o_b[:, 0:3])                         # accelerometer
    np.testing.assert_array_equal(o_f[:, 3:5], np.clip(a, -1, 1))                   # ctrl = clipped last action
    np.testing.assert_array_equal(o_f[:, 5:7], o_b[:, 3:5])                         # goal_compass
    dist = np.sqrt(np.sum((st[:, 11:13] - pos[:, :2]) ** 2, axis=1))
    np.testing.assert_allclose(o_f[:, 7], np.exp(-dist), rtol=2e-7)                 # goal_dist
    np.testing.assert_array_equal(o_f[:, 8:14], o_b[:, 5:11])                       # gyro, magnetometer
    np.testing.assert_array_equal(o_f[:, 14:17], st[:, 0:3].astype(np.float32))     # qpos
    np.testing.assert_array_equal(o_f[:, 17:20], st[:, 3:6].astype(np.float32))     # qvel
    np.testing.assert_array_equal(o_f[:, 20:23], o_b[:, 11:14])                     # velocimeter


def test_car_optional_keys_match_oracle(cuda_lib):
    from mobrob_b200 import GpuVecEnv

    cfg = {k: True for k in KEYS}
    n, seed, tl = 10, 2, 20
    ora = GoalVecOracle(co.CarBody(n), seed=seed, time_limit=tl, terminate_on_goal=True, observe=cfg)
    gpu = GpuVecEnv("car", n, seed=seed, time_limit=tl, terminate_on_goal=True, robot_config=cfg)
    assert gpu.obs_dim == 26 + 1 + 13 + 11 + 2
    o_ref, o_gpu = ora.reset(), gpu.reset()
    np.testing.assert_allclose(o_gpu, o_ref, rtol=1e-5, atol=2e-6)
    rng = np.random.default_rng(4)
    n_done = 0
    for t in range(45):
        a = np.sign(rng.standard_normal((n, 2))).astype(np.float32)
        o_ref, r_ref, d_ref, info = ora.step(a)
        o_gpu, r_gpu, d_gpu, infos = gpu.step(a)
        np.testing.assert_array_equal(d_gpu, d_ref)
        np.testing.assert_allclose(o_gpu, o_ref, rtol=1e-4, atol=1e-4, err_msg=f"obs step {t}")   # contacts: test_car_gpu's bound
        for i in np.nonzero(d_ref)[0]:
            n_done += 1
            np.testing.assert_allclose(infos[i]["terminal_observation"], info["terminal_obs"][i], rtol=1e-4, atol=1e-4)
    assert n_done >= n
    # layout: ... ballquat_rear (6:15) | ctrl | goal_compass | goal_dist | gyro mag | qpos 13 | qvel 11 | velocimeter
    st = gpu.get_state().cpu().numpy()
    np.testing.assert_array_equal(o_gpu[:, 15:17], st[:, 24:26].astype(np.float32))
    np.testing.assert_array_equal(o_gpu[:, 26:39], st[:, 0:13].astype(np.float32))
    np.testing.assert_array_equal(o_gpu[:, 39:50], st[:, 13:24].astype(np.float32))


def test_wrapper_subclass_switches_keys_through_get_robot_config(cuda_lib):
    """The reference's extension point: a MujocoGoalEnv subclass overrides get_robot_config (wrapper.py:235-240)."""
    from mobrob_b200.envs.wrapper import PointEnv

    class PointWithState(PointEnv):
        def get_robot_config(self):
            return {**super().get_robot_config(), "observe_qpos": True, "observe_goal_dist": True}

    env = PointWithState(terminate_on_goal=False)
    assert env.observation_space.shape == (14 + 3 + 1,)
    env.seed(3)
    obs, _ = env.reset()
    assert obs.shape == (18,)
    goal, pos = np.asarray(env.get_goal(), np.float64), np.asarray(env.get_pos(), np.float64)
    np.testing.assert_allclose(obs[5], np.exp(-np.linalg.norm(goal[:2] - pos[:2])), rtol=2e-7)   # goal_dist after goal_compass
    np.testing.assert_array_equal(obs[12:15], np.zeros(3, np.float32))   # qpos is relative to the placed body: 0 after reset
    obs2, rew, term, trunc, info = env.step(np.array([1.0, 0.3], np.float32))
    assert obs2.shape == (18,) and np.any(obs2[12:15] != 0)

    class WithLidar(PointEnv):
        def get_robot_config(self):
            return {**super().get_robot_config(), "observe_hazards": True}

    with pytest.raises(NotImplementedError):
        WithLidar()


def test_ppo_trains_on_the_extended_observation(cuda_lib):
    """PPO over a point env with every optional key (23 floats): the stand-alone kernels collect the rollout (the fused
    rollout kernel is built for the default row), the tensor-core update takes the wider rows; the first rollout row is
    the env's reset observation and the update moves the parameters.  The car with qpos / qvel (53 floats) is refused."""
    from mobrob_b200 import GpuVecEnv
    from mobrob_b200.ppo import PPO

    cfg = {k: True for k in KEYS}
    env = GpuVecEnv("point", 256, seed=1, time_limit=100, terminate_on_goal=True, robot_config=cfg)
    model = PPO("MlpPolicy", env, n_steps=32, batch_size=1024, n_epochs=2, seed=1, verbose=0)
    first = env.reset_tensor().clone()
    model._last_obs = None
    before = model.updater.params.clone()
    model.learn(total_timesteps=2 * 256 * 32)
    assert model.updater.obs_dim == 23 and model.buf["obs"].shape[-1] == 23
    assert torch.isfinite(model.updater.params).all() and not torch.equal(before, model.updater.params)
    act, _ = model.predict(first.cpu().numpy(), deterministic=True)
    assert act.shape == (256, 2)

    car = GpuVecEnv("car", 8, seed=1, robot_config={"observe_qpos": True})
    with pytest.raises(ValueError, match="1..31"):
        PPO("MlpPolicy", car, n_steps=8, batch_size=64, n_epochs=1, verbose=0)
