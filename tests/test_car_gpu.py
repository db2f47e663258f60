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


def test_car_ppo_iteration_and_pretrained_policy(cuda_lib, golden_dir):
    """(a) collect_rollouts for the car (unfused kernels) replayed through the oracle step by step;
    (b) one PPO update runs on the 26-dim observation; (c) the shipped car policy drives the CUDA car
    to its goals (functional check of the contact model: the reference trained it on real MuJoCo)."""
    import os
    import sys

    from mobrob_b200 import GpuVecEnv
    from mobrob_b200.ppo import PPO
    from oracle import sb3_oracle

    n, T, seed = 6, 24, 1
    env = GpuVecEnv("car", n, seed=None, time_limit=1000, terminate_on_goal=True)
    model = PPO("MlpPolicy", env, n_steps=T, batch_size=48, n_epochs=2, gae_lambda=0.5, ent_coef=0.05, seed=seed)
    w = dict(np.load(os.path.join(golden_dir, "car_policy.npz")))
    model.policy.load_state_dict({k: torch.as_tensor(v) for k, v in w.items()})
    ref_pol = sb3_oracle.MlpPolicyOracle(26).load_numpy(w)
    venv = GoalVecOracle(co.CarBody(n), seed=seed, time_limit=1000, terminate_on_goal=True)
    obs_ref = venv.reset()
    eps = torch.randn((T, n, 2), generator=torch.Generator().manual_seed(3))
    model.collect_rollouts(eps.cuda().contiguous())
    torch.cuda.synchronize()
    b = {k: v.cpu().numpy() for k, v in model.buf.items()}
    for t in range(T):
        np.testing.assert_allclose(b["obs"][t], obs_ref, rtol=1e-4, atol=1e-4, err_msg=f"obs t={t}")
        with torch.no_grad():
            a_ref, v_ref, lp_ref = ref_pol.forward_with_noise(torch.as_tensor(b["obs"][t]), eps[t])
        assert np.abs(b["actions"][t] - a_ref.numpy()).max() <= 1e-5 * max(1.0, float(a_ref.abs().max()))
        np.testing.assert_allclose(b["values"][t], v_ref.numpy(), rtol=1e-5, atol=1e-5)
        obs_ref, rew, done, info = venv.step(np.clip(b["actions"][t], -1, 1))
        np.testing.assert_allclose(b["rewards"][t], rew, rtol=1e-4, atol=1e-5)
    p0 = model.updater.params.clone()
    model.train()
    torch.cuda.synchronize()
    assert torch.isfinite(model.updater.params).all() and float((model.updater.params - p0).abs().max()) > 0

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import eval_policy

    res = eval_policy.evaluate_gpu(os.path.join(golden_dir, "policies", "car-ppo.zip"), 512, steps=600, env_name="car")
    print("car policy:", {k: v for k, v in res.items() if not k.startswith("first_l") and k != "first_ok"})
    assert res["first_goal_success_rate"] > 0.5
