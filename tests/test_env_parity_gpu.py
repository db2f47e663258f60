"""CUDA point environment vs the CPU oracle on the same seeds and actions (through the C ABI)."""
import numpy as np
import pytest
import torch

from oracle import point_oracle as po, ref_rng
from oracle.vec_oracle import GoalVecOracle

pytestmark = pytest.mark.gpu

# tolerance named by BASELINE.json north_star: 1e-5 relative on body state and observations
RTOL = 1e-5
ATOL_OBS = 1e-6  # observation entries that are exactly 0 / pass through float32 rounding


def _actions(rng, n, t):
    """bang-bang held for random durations plus un-saturated uniform phases."""
    if t % 11 == 0:
        return rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    if t % 13 == 0:
        return rng.uniform(-0.06, 0.06, (n, 2)).astype(np.float32)
    return np.sign(rng.standard_normal((n, 2))).astype(np.float32) * rng.choice([1.0, 1.7], (n, 2)).astype(np.float32)


def _oracle_state(env: GoalVecOracle):
    b = env.body
    return np.concatenate([b.state_vector(), env.goal.astype(np.float64),
                           env.elapsed[:, None].astype(np.float64), env.ep_ret[:, None]], axis=1)


@pytest.mark.parametrize("seed,n,time_limit", [(0, 64, 1000), (7, 33, 60)])
def test_point_vec_env_matches_oracle(cuda_lib, seed, n, time_limit):
    from mobrob_b200 import GpuVecEnv

    steps = 1000
    ora = GoalVecOracle(po.PointBody(n), seed=seed, time_limit=time_limit, terminate_on_goal=True)
    gpu = GpuVecEnv("point", n, seed=seed, time_limit=time_limit, terminate_on_goal=True)
    o_ref = ora.reset()
    o_gpu = gpu.reset()
    np.testing.assert_allclose(o_gpu, o_ref, rtol=RTOL, atol=ATOL_OBS)
    st = gpu.get_state().cpu().numpy()
    np.testing.assert_array_equal(st[:, 6:9], _oracle_state(ora)[:, 6:9])  # init xy, heading: bit exact
    np.testing.assert_array_equal(st[:, 11:13], ora.goal.astype(np.float64))  # goals: bit exact
    rng = np.random.default_rng(seed + 100)
    n_done = n_trunc = 0
    for t in range(steps):
        a = _actions(rng, n, t)
        o_ref, r_ref, d_ref, info = ora.step(a)
        o_gpu, r_gpu, d_gpu, infos = gpu.step(a)
        np.testing.assert_array_equal(d_gpu, d_ref, err_msg=f"done flags differ at step {t}")
        np.testing.assert_allclose(o_gpu, o_ref, rtol=RTOL, atol=ATOL_OBS, err_msg=f"obs step {t}")
        np.testing.assert_allclose(r_gpu, r_ref, rtol=RTOL, atol=1e-7, err_msg=f"reward step {t}")
        for i in np.nonzero(d_ref)[0]:
            n_done += 1
            n_trunc += int(info["truncated"][i])
            assert infos[i]["TimeLimit.truncated"] == bool(info["truncated"][i])
            assert infos[i]["episode"]["l"] == int(info["ep_l"][i])
            assert abs(infos[i]["episode"]["r"] - info["ep_r"][i]) <= 1e-5 * max(1.0, abs(info["ep_r"][i]))
            np.testing.assert_allclose(infos[i]["terminal_observation"], info["terminal_obs"][i],
                                       rtol=RTOL, atol=ATOL_OBS)
        if t % 100 == 99 or t == steps - 1:
            st = gpu.get_state().cpu().numpy()
            ref = _oracle_state(ora)
            scale = np.maximum(np.abs(ref), 1.0)
            assert np.max(np.abs(st - ref) / scale) < RTOL, f"state diverged at step {t}"
            np.testing.assert_array_equal(st[:, 13], ref[:, 13])  # episode step index: bit exact
    counts = gpu.get_reset_counts().cpu().numpy()
    np.testing.assert_array_equal(counts[:, 0], ora.n_resets)
    np.testing.assert_array_equal(counts[:, 1], ora.n_full)
    assert n_done > 0
    if time_limit < 1000:
        assert n_trunc > 0


def test_point_contact_free_1000_step_trajectory(cuda_lib):
    """north_star parity case: 1000-step trajectories, no resets (goal far away, no time limit)."""
    from mobrob_b200 import GpuVecEnv

    n = 256
    gpu = GpuVecEnv("point", n, seed=11, time_limit=None, terminate_on_goal=False)
    gpu.reset()
    st0 = gpu.get_state().cpu().numpy()
    body = po.PointBody(n)
    body.q[:] = st0[:, 0:3]; body.v[:] = st0[:, 3:6]; body.body_xy[:] = st0[:, 6:8]
    body.psi0[:] = st0[:, 8]; body.ctrl[:] = st0[:, 9:11]
    goal = st0[:, 11:13].astype(np.float32)
    rng = np.random.default_rng(5)
    hold = np.zeros((n, 2), np.int64)
    a = np.zeros((n, 2), np.float32)
    for t in range(1000):
        renew = hold <= 0
        a = np.where(renew, np.sign(rng.standard_normal((n, 2))), a).astype(np.float32)
        hold = np.where(renew, rng.integers(1, 51, (n, 2)), hold) - 1
        body.step(a)
        o_gpu, _, d, _ = gpu.step(a)
        assert not d.any()
        if t % 50 == 49:
            np.testing.assert_allclose(o_gpu, body.obs(goal), rtol=RTOL, atol=ATOL_OBS)
    st = gpu.get_state().cpu().numpy()
    ref = body.state_vector()
    err = np.abs(st[:, :11] - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < RTOL
    # the physics really ran: robots moved metres and turned many radians
    assert np.abs(ref[:, 2]).max() > 5.0 and np.abs(ref[:, 0:2]).max() > 1.0


def test_device_rng_streams_match_numpy(cuda_lib):
    """PCG64 (Box.sample) and MT19937 (Engine heading) restated on the device are bit exact."""
    from mobrob_b200 import GpuVecEnv

    n, seed = 512, 1234
    gpu = GpuVecEnv("point", n, seed=seed, time_limit=1, terminate_on_goal=True)
    gpu.reset()
    init = [ref_rng.init_box() for _ in range(n)]
    goal = [ref_rng.goal_box() for _ in range(n)]
    for i in range(n):
        init[i].seed(seed + i)
        goal[i].seed(seed + i + 1)
    eng = seed + np.arange(n)
    xy = np.zeros((n, 2)); head = np.zeros(n); g = np.zeros((n, 2))
    full = np.ones(n, bool)
    zero = np.zeros((n, 2), np.float32)
    n_goal_only = 0
    for rnd in range(8):
        for i in range(n):
            if full[i]:
                eng[i] += 2
                xy[i] = init[i].sample()
                head[i] = ref_rng.engine_heading(int(eng[i]))
            g[i] = goal[i].sample()
        st = gpu.get_state().cpu().numpy()
        np.testing.assert_array_equal(st[:, 6:8], xy)
        np.testing.assert_array_equal(st[:, 8], head)
        np.testing.assert_array_equal(st[:, 11:13], g)
        # time_limit = 1 and a zero action on a resting robot: every step ends the episode;
        # the robot is re-placed unless the goal happens to lie within the reach radius
        dx = g - xy
        full = ~(np.sqrt(dx[:, 0] * dx[:, 0] + dx[:, 1] * dx[:, 1]) < 0.3)
        n_goal_only += int((~full).sum())
        gpu.step(zero)
    assert n_goal_only > 0


def test_point_fast_spin_and_large_heading_paths(cuda_lib):
    """The env-step kernel's cold paths: |h omega| beyond the small-angle rotation (|omega| >= 5 rad/s, only
    reachable through set_state) and headings beyond the Cody-Waite range (|psi| >= 1e5 rad: library sincos)."""
    from mobrob_b200 import GpuVecEnv

    n = 96
    gpu = GpuVecEnv("point", n, seed=3, time_limit=None, terminate_on_goal=False)
    gpu.reset()
    st = gpu.get_state().cpu().numpy()
    rng = np.random.default_rng(9)
    st[:, 5] = np.where(np.arange(n) % 3 == 0, rng.uniform(-60, 60, n), st[:, 5])       # spin: up to 60 rad/s
    st[:, 5] = np.where(np.arange(n) % 3 == 1, rng.uniform(4.9, 5.1, n), st[:, 5])      # straddles the switch
    st[:, 2] = np.where(np.arange(n) % 4 == 0, rng.uniform(-3e5, 3e5, n), st[:, 2])     # hinge angle: huge heading
    st[:, 3:5] = rng.uniform(-2, 2, (n, 2))
    gpu.set_state(torch.as_tensor(st))
    body = po.PointBody(n)
    body.q[:] = st[:, 0:3]; body.v[:] = st[:, 3:6]; body.body_xy[:] = st[:, 6:8]
    body.psi0[:] = st[:, 8]; body.ctrl[:] = st[:, 9:11]
    goal = st[:, 11:13].astype(np.float32)
    for t in range(12):
        a = np.sign(rng.standard_normal((n, 2))).astype(np.float32)
        body.step(a)
        o_gpu, _, d, _ = gpu.step(a)
        assert not d.any()
        np.testing.assert_allclose(o_gpu, body.obs(goal), rtol=RTOL, atol=2e-6, err_msg=f"obs step {t}")
    got = gpu.get_state().cpu().numpy()[:, :11]
    ref = body.state_vector()
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < RTOL, err.max()


@pytest.mark.parametrize("env_name", ["point", "car"])
def test_reset_spaces_of_the_batched_env(cuda_lib, env_name):
    """EnvWrapper.reset_init_space / reset_goal_space (wrapper.py:209-219) on the batched env: resets after the
    call draw start positions and goals from the new boxes, from each env's own continuing stream -- bit-exact
    against the oracle's streams (numpy PCG64)."""
    from mobrob_b200 import GpuVecEnv
    from mobrob_b200.spaces import Box
    from oracle import car_oracle as co

    n, seed = 40, 9
    body = po.PointBody(n) if env_name == "point" else co.CarBody(n)
    env = GpuVecEnv(env_name, n, seed=seed, time_limit=7, terminate_on_goal=True)
    ora = GoalVecOracle(body, seed=seed, time_limit=7, terminate_on_goal=True)
    env.reset(); ora.reset()
    lo_i, hi_i = np.array([0.5, -0.25], np.float32), np.array([0.75, 0.0], np.float32)
    lo_g, hi_g = np.array([-2.0, 1.5], np.float32), np.array([-1.75, 2.0], np.float32)
    env.reset_init_space(Box(lo_i, hi_i, dtype=np.float32)); ora.reset_init_space(lo_i, hi_i)
    env.reset_goal_space(Box(lo_g, hi_g, dtype=np.float32)); ora.reset_goal_space(lo_g, hi_g)
    a = np.zeros((n, 2), np.float32)
    for t in range(8):          # the 7-step time limit truncates every env once: full resets into the new boxes
        o, r, d, _ = env.step(a)
        o_ref, r_ref, d_ref, _ = ora.step(a)
        np.testing.assert_array_equal(d, d_ref)
    st = env.get_state().cpu().numpy()
    goal_slot = 11 if env_name == "point" else 26
    np.testing.assert_array_equal(st[:, goal_slot:goal_slot + 2].astype(np.float32), ora.goal)
    pos = env.get_pos().cpu().numpy()
    np.testing.assert_allclose(pos, ora.body.pos(), rtol=1e-5, atol=1e-6)
    assert (pos[:, 0] >= 0.45).all() and (pos[:, 0] <= 0.8).all() and (pos[:, 1] >= -0.3).all() and (pos[:, 1] <= 0.05).all()
    assert (ora.goal[:, 0] <= -1.75).all() and (ora.goal[:, 1] >= 1.5).all()
