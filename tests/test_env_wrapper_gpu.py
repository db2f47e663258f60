"""The reference's single-environment API -- ``get_env(...)`` -> EnvWrapper / TimeLimit
(src/mobrob/envs/wrapper.py:95-228, 290-326, 549-571) -- driven the way examples/control.py and
SB3's DummyVecEnv drive it, against the CPU oracle of the same stack with one environment.

Covers SURVEY 8a rows E1-E9 / 8b-B1 through the Python mirror (mobrob_b200/envs/wrapper.py):
seed, reset (full vs goal-only), step's 5-tuple, reward_fn, reached, set_pos / set_goal / get_pos /
get_obs, the space getters, reset_init_space / reset_goal_space and the TimeLimit wrapper.
"""
import numpy as np
import pytest

from oracle import car_oracle as co, point_oracle as po
from oracle.vec_oracle import GoalVecOracle

pytestmark = pytest.mark.gpu


def _drive(env_name, body, seed, time_limit, steps, hold, rtol=1e-5, atol=2e-6):
    """DummyVecEnv.step_wait semantics on both sides: step, and reset (no seed) when done."""
    from mobrob_b200 import get_env

    env = get_env(env_name, enable_gui=False, terminate_on_goal=True, time_limit=time_limit)
    ora = GoalVecOracle(body, seed=seed, time_limit=time_limit, terminate_on_goal=True)
    obs, info = env.reset(seed=seed)
    assert info == {}
    obs_ref = ora.reset()
    assert obs.dtype == np.float32 and obs.shape == env.observation_space.shape
    np.testing.assert_allclose(obs, obs_ref[0], rtol=rtol, atol=atol)
    np.testing.assert_array_equal(env.get_goal(), ora.goal[0])
    rng = np.random.default_rng(seed)
    n_term = n_trunc = 0
    a = np.zeros(2, np.float32)
    for t in range(steps):
        if t % hold == 0:
            a = np.sign(rng.standard_normal(2)).astype(np.float32)
        o, r, terminated, truncated, info = env.step(a)
        o_ref, r_ref, d_ref, i_ref = ora.step(a[None])
        assert isinstance(terminated, bool) and isinstance(truncated, bool)
        assert info.get("cost") == 0.0
        assert (terminated or truncated) == bool(d_ref[0]), f"done flag t={t}"
        assert terminated == bool(i_ref["terminated"][0]) and (truncated and not terminated) == bool(i_ref["truncated"][0])
        np.testing.assert_allclose(r, r_ref[0], rtol=rtol, atol=atol, err_msg=f"reward t={t}")
        if terminated or truncated:
            np.testing.assert_allclose(o, i_ref["terminal_obs"][0], rtol=rtol, atol=atol, err_msg=f"terminal obs t={t}")
            assert env.reached() == terminated
            n_term += terminated
            n_trunc += truncated and not terminated
            o, _ = env.reset()
            np.testing.assert_array_equal(env.get_goal(), ora.goal[0])
        np.testing.assert_allclose(o, o_ref[0], rtol=rtol, atol=atol, err_msg=f"obs t={t}")
        np.testing.assert_allclose(env.get_pos(), ora.body.pos()[0], rtol=rtol, atol=atol)
    env.close()
    return n_term, n_trunc, int(ora.n_full[0]), int(ora.n_resets[0])


def test_point_env_wrapper_matches_oracle(cuda_lib):
    n_term, n_trunc, n_full, n_resets = _drive("point", po.PointBody(1), seed=3, time_limit=60, steps=900, hold=25)
    assert n_trunc > 0 and n_resets == 1 + n_term + n_trunc
    # a goal-only reset follows a reached goal; a full reset everything else
    assert n_full == 1 + n_trunc


def test_point_env_reaches_goals_with_shipped_policy(cuda_lib, golden_dir):
    """examples/control.py:33-49 -- get_env + policy.predict(obs, deterministic=True) + step, one env:
    the shipped policy reaches goals (terminated) and the wrapper resets goal-only afterwards."""
    import os

    from mobrob_b200 import get_env
    from mobrob_b200.ppo import PPO

    env = get_env("point", enable_gui=False, terminate_on_goal=True)
    policy = PPO.load(os.path.join(golden_dir, "policies", "point-ppo.zip"))
    obs, _ = env.reset(seed=0)
    reached, cum = 0, 0.0
    for _ in range(600):
        action, state = policy.predict(obs, deterministic=True)
        assert state is None and action.shape == (2,)
        obs, r, terminated, truncated, _ = env.step(action)
        cum += r
        if terminated:
            reached += 1
            pos = env.get_pos().copy()
            obs, _ = env.reset()
            np.testing.assert_allclose(env.get_pos(), pos, atol=1e-12)   # goal-only reset keeps the body where it is
    assert reached >= 2 and cum > 5.0 * reached


def test_car_env_wrapper_matches_oracle(cuda_lib):
    n_term, n_trunc, n_full, n_resets = _drive("car", co.CarBody(1), seed=5, time_limit=40, steps=200, hold=10,
                                               rtol=1e-4, atol=1e-5)   # soft contacts: tests/test_car_gpu.py's tolerance
    assert n_trunc > 0 and n_resets == 1 + n_term + n_trunc


def test_wrapper_api_surface(cuda_lib):
    """README.md:80-94 of the reference: the methods a user of EnvWrapper calls."""
    from mobrob_b200 import get_env
    from mobrob_b200.spaces import Box

    with pytest.raises(ValueError):
        get_env("no-such-robot")
    env = get_env("point", terminate_on_goal=False)
    assert env.get_observation_space().shape == (14,) and env.get_action_space().shape == (2,)
    np.testing.assert_array_equal(env.get_init_space().low, [-1, -1]); np.testing.assert_array_equal(env.get_init_space().high, [1, 1])
    np.testing.assert_array_equal(env.get_goal_space().low, [-2, -2]); np.testing.assert_array_equal(env.get_goal_space().high, [2, 2])
    env.reset(seed=11)
    env.set_pos(np.array([0.25, -0.5]))
    np.testing.assert_allclose(env.get_pos(), [0.25, -0.5], atol=1e-12)
    env.set_goal(np.array([0.3, -0.5], dtype=np.float32))
    assert env.reached() and env.reached(reach_radius=0.04) is False
    env.reward_fn()   # moves _prev_pos to the new position (set_pos does not, wrapper.py:137-154)
    # reward_fn (wrapper.py:137-154): progress towards the goal + 5 inside the radius; no termination asked for
    obs, r, terminated, truncated, _ = env.step(np.zeros(2, np.float32))
    assert terminated is False and truncated is False and 4.9 < r < 5.1
    # compass of the observation points at the goal: unit-ish vector (d / (d + 0.001))
    assert abs(np.linalg.norm(obs[3:5]) - 0.05 / 0.051) < 1e-3
    # reset_init_space / reset_goal_space (wrapper.py:209-219): later resets draw from the new boxes
    env.reset_init_space(Box(low=np.array([0.9, 0.9], np.float32), high=np.array([1.0, 1.0], np.float32), dtype=np.float32))
    env.reset_goal_space(Box(low=np.array([-2.0, -2.0], np.float32), high=np.array([-1.9, -1.9], np.float32), dtype=np.float32))
    env.reset(seed=4)   # the goal is reached: goal-only reset (wrapper.py:182-191), the body stays
    np.testing.assert_allclose(env.get_pos(), [0.25, -0.5], atol=1e-6)
    assert (env.get_goal() <= -1.9).all() and not env.reached()
    env.reset()         # not reached: full reset, position drawn from the new init space
    assert (env.get_pos() >= 0.9).all() and (env.get_goal() <= -1.9).all()
    # init_pos argument of reset (wrapper.py:193-194)
    env.reset(init_pos=np.array([-0.75, 0.125]))
    np.testing.assert_allclose(env.get_pos(), [-0.75, 0.125], atol=1e-12)
    env.close()
