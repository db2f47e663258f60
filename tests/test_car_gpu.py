"""CUDA car environment vs the CPU oracle (through the C ABI)."""
import numpy as np
import pytest
import torch

from oracle import car_oracle as co
from oracle.vec_oracle import GoalVecOracle

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _state(body):
    return body.state_vector()


def _load_state(body, st):
    body.p[:] = st[:, 0:3]; body.quat[:] = st[:, 3:7]; body.th[:] = st[:, 7:9]; body.qb[:] = st[:, 9:13]
    body.v[:] = st[:, 13:16]; body.w[:] = st[:, 16:19]; body.s[:] = st[:, 19:21]; body.wb[:] = st[:, 21:24]
    body.ctrl[:] = st[:, 24:26]


def test_car_contact_free_trajectory(cuda_lib):
    """north_star parity case for the car: contact-free (airborne) trajectories, 400 steps = 4000
    substeps of the articulated free-joint + hinge + ball dynamics, motors bang-bang."""
    from mobrob_b200 import GpuVecEnv

    n = 48
    gpu = GpuVecEnv("car", n, seed=3, time_limit=None, terminate_on_goal=False)
    gpu.set_contacts(False)
    gpu.reset()
    rng = np.random.default_rng(0)
    st = gpu.get_state().cpu().numpy()
    st[:, 2] = 50.0 + rng.random(n)                      # high above the floor
    q = rng.standard_normal((n, 4)); st[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    q = rng.standard_normal((n, 4)); st[:, 9:13] = q / np.linalg.norm(q, axis=1, keepdims=True)
    st[:, 13:16] = rng.standard_normal((n, 3))
    st[:, 16:19] = rng.standard_normal((n, 3)) * 2.0
    st[:, 19:21] = rng.standard_normal((n, 2)) * 10.0
    st[:, 21:24] = rng.standard_normal((n, 3)) * 5.0
    st[:, 26:28] = 100.0                                 # goal far away
    gpu.set_state(torch.as_tensor(st))
    body = co.CarBody(n)
    body.contacts_enabled = False
    _load_state(body, st)
    goal = st[:, 26:28].astype(np.float32)
    np.testing.assert_allclose(gpu.get_obs_tensor().cpu().numpy(), body.obs(goal), rtol=RTOL, atol=2e-6)
    a = np.zeros((n, 2), np.float32)
    for t in range(400):
        if t % 7 == 0:
            a = (np.sign(rng.standard_normal((n, 2))) * rng.choice([1.0, 0.01], (n, 2))).astype(np.float32)
        body.step(a)
        o, _, d, _ = gpu.step(a)
        assert not d.any()
        if t % 50 == 49:
            np.testing.assert_allclose(o, body.obs(goal), rtol=RTOL, atol=2e-6, err_msg=f"obs t={t}")
    ref = _state(body)
    got = gpu.get_state().cpu().numpy()[:, :26]
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < RTOL, f"state error {err.max():.2e}"
    assert np.abs(ref[:, 7:9]).max() > 20.0  # wheels really spun up


def test_car_vec_env_on_the_floor(cuda_lib):
    """Driving on the floor (soft contacts, projected Gauss-Seidel): same algorithm on both sides."""
    from mobrob_b200 import GpuVecEnv

    n, seed, tl = 12, 4, 25
    ora = GoalVecOracle(co.CarBody(n), seed=seed, time_limit=tl, terminate_on_goal=True)
    gpu = GpuVecEnv("car", n, seed=seed, time_limit=tl, terminate_on_goal=True)
    o_ref, o_gpu = ora.reset(), gpu.reset()
    assert o_gpu.shape == (n, 26)
    np.testing.assert_allclose(o_gpu, o_ref, rtol=RTOL, atol=2e-6)
    st = gpu.get_state().cpu().numpy()
    np.testing.assert_array_equal(st[:, 26:28], ora.goal.astype(np.float64))     # goals bit exact
    np.testing.assert_array_equal(st[:, 0:2], ora.body.p[:, :2])                # init xy bit exact
    np.testing.assert_allclose(st[:, 3:7], ora.body.quat, rtol=0, atol=1e-15)   # heading quat (sincos ulp)
    rng = np.random.default_rng(1)
    n_done = 0
    for t in range(60):
        a = np.sign(rng.standard_normal((n, 2))).astype(np.float32)
        o_ref, r_ref, d_ref, info = ora.step(a)
        o_gpu, r_gpu, d_gpu, infos = gpu.step(a)
        np.testing.assert_array_equal(d_gpu, d_ref, err_msg=f"done flags step {t}")
        np.testing.assert_allclose(o_gpu, o_ref, rtol=1e-4, atol=1e-4, err_msg=f"obs step {t}")
        np.testing.assert_allclose(r_gpu, r_ref, rtol=1e-4, atol=1e-6)
        n_done += int(d_ref.sum())
    got = gpu.get_state().cpu().numpy()
    ref = np.concatenate([ora.body.state_vector(), ora.goal.astype(np.float64),
                          ora.elapsed[:, None].astype(np.float64), ora.ep_ret[:, None]], 1)
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < 1e-4, f"state error {err.max():.2e}"
    assert n_done >= n  # every env hit the 25-step limit at least once -> full resets exercised
    counts = gpu.get_reset_counts().cpu().numpy()
    np.testing.assert_array_equal(counts[:, 0], ora.n_resets)
    np.testing.assert_array_equal(counts[:, 1], ora.n_full)
    # KAT-3 style invariants on the device observation
    for r in o_gpu.astype(np.float64):
        R = r[6:15].reshape(3, 3)
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-5)
        assert abs(np.linalg.norm(r[20:23]) - 0.5) < 1e-5
        assert np.linalg.norm(r[15:17]) < 1.0
    # the car rests on its wheels: body height a fraction of a millimetre below the 0.1 m rest pose
    assert np.all(np.abs(got[:, 2] - 0.1) < 5e-3)
