"""Pin the CPU oracle on the artefacts the reference ships."""
import json
import os

import numpy as np
import torch

from oracle import c_oracle, point_oracle as po, ref_rng, sb3_oracle


def _body_from_obs_row(r):
    psi = np.arctan2(-r[8], -r[9])
    b = po.PointBody(1)
    b.q[0, 2] = psi
    c, s = np.cos(psi), np.sin(psi)
    b.v[0] = [c * r[11] - s * r[12], s * r[11] + c * r[12], r[7]]
    return b


def test_kat1_point_accelerometer(golden_dir):
    """Real-MuJoCo sensor rows: accelerometer reproduced from velocimeter/gyro/magnetometer."""
    obs = np.load(os.path.join(golden_dir, "point_last_obs.npy")).astype(np.float64)
    assert obs.shape == (2, 14)
    for r in obs:
        b = _body_from_obs_row(r)
        b.ctrl[0] = [-1.0, 1.0]  # saturated bang-bang action of the shipped policy
        o = b.obs(np.zeros((1, 2), np.float32))[0].astype(np.float64)
        np.testing.assert_allclose(o[:2], r[:2], rtol=2e-6)
        assert np.float32(o[2]) == np.float32(9.81) == np.float32(r[2])
        np.testing.assert_allclose(o[8:10], r[8:10], atol=1e-7)
        np.testing.assert_allclose(o[11:13], r[11:13], atol=1e-7)
        assert r[5] == r[6] == r[10] == r[13] == 0.0
        assert abs(np.hypot(r[8], r[9]) - 0.5) < 1e-7


def test_kat1_rules_out_wrong_models(golden_dir):
    """The fixture is sharp enough to reject a model with sliding friction or a wrong gear."""
    obs = np.load(os.path.join(golden_dir, "point_last_obs.npy")).astype(np.float64)
    r = obs[0]
    b = _body_from_obs_row(r)
    b.ctrl[0] = [-0.5, 1.0]  # motor force still saturated at |ctrl| >= 0.05 -> identical
    o = b.obs(np.zeros((1, 2), np.float32))[0]
    np.testing.assert_allclose(o[:2], r[:2], rtol=2e-6)
    b.ctrl[0] = [-0.04, 1.0]  # un-saturated motor -> must differ
    o = b.obs(np.zeros((1, 2), np.float32))[0]
    assert abs(o[0] - r[0]) > 1e-2


def test_kat2_policy_forward(golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "kat2.json")))
    # values frozen in SURVEY.md appendix A.7
    a7 = {"point": ([[-24.10194778, -0.75789261], [-32.90559006, 7.31994486]], [1.50718296, 1.51591945]),
          "car": ([[-307.91055, -296.65634], [-307.80740, -221.70090], [287.30377, 285.68530],
                   [242.41867, -39.32915]], [1.62879467, 1.58873570, 1.68457210, 1.58394730])}
    for env, O in (("point", 14), ("car", 26)):
        w = np.load(os.path.join(golden_dir, f"{env}_policy.npz"))
        assert list(w.keys()) == sb3_oracle.PARAM_ORDER
        pol = sb3_oracle.MlpPolicyOracle(O).load_numpy(dict(w))
        obs = torch.as_tensor(np.load(os.path.join(golden_dir, f"{env}_last_obs.npy")))
        with torch.no_grad():
            mu, v = pol.mean_value(obs)
        np.testing.assert_allclose(mu.numpy(), np.array(kat[env]["mu"]), rtol=1e-5)
        np.testing.assert_allclose(v.numpy(), np.array(kat[env]["v"]), rtol=1e-5)
        np.testing.assert_allclose(mu.numpy(), np.array(a7[env][0]), rtol=2e-5)
        np.testing.assert_allclose(v.numpy(), np.array(a7[env][1]), rtol=2e-5)
        assert pol.flat_params().numel() == {14: 10437, 26: 11973}[O]


def test_kat3_car_layout(golden_dir):
    obs = np.load(os.path.join(golden_dir, "car_last_obs.npy")).astype(np.float64)
    assert obs.shape == (4, 26)
    for r in obs:
        R = r[6:15].reshape(3, 3)
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-6)
        assert abs(np.linalg.det(R) - 1) < 1e-6
        assert abs(np.linalg.norm(r[20:23]) - 0.5) < 1e-6
        assert 0.9 < np.linalg.norm(r[15:17]) < 1.0


def test_c_oracle_matches_numpy_oracle():
    rng = np.random.default_rng(0)
    n = 64
    b = po.PointBody(n)
    for i in range(n):
        b.full_reset(i, rng.uniform(-1, 1, 2).astype(np.float32), rng.uniform(0, 2 * np.pi))
    st = b.state_vector().copy()
    goal = rng.uniform(-2, 2, (n, 2)).astype(np.float32)
    for t in range(200):
        a = (np.sign(rng.standard_normal((n, 2))) if t % 5 else rng.uniform(-1, 1, (n, 2))).astype(np.float32)
        b.step(a)
        c_oracle.physics_step(st, a)
    np.testing.assert_array_equal(b.state_vector(), st)
    np.testing.assert_array_equal(b.obs(goal), c_oracle.obs(st, goal))
    np.testing.assert_array_equal(b.pos(), c_oracle.pos(st))


def test_ref_rng_fixture(golden_dir):
    fx = json.load(open(os.path.join(golden_dir, "ref_rng.json")))
    for s, h in fx["heading"].items():
        assert ref_rng.engine_heading(int(s)) == h
        assert 0.0 <= h < 2 * np.pi
    b = ref_rng.init_box()
    b.seed(0)
    for row in fx["init_seed0"]:
        s = b.sample()
        assert s.dtype == np.float32 and np.all(np.abs(s) <= 1)
        np.testing.assert_array_equal(s.astype(np.float64), np.array(row))


def test_vec_oracle_semantics():
    """Goal-only reset keeps the robot state; time-limit reset re-places it; reward telescopes."""
    from oracle.vec_oracle import GoalVecOracle

    n = 8
    env = GoalVecOracle(po.PointBody(n), seed=3, time_limit=50, terminate_on_goal=True)
    obs = env.reset()
    assert obs.shape == (n, 14) and obs.dtype == np.float32
    assert (env.n_full == 1).all() and (env.n_resets == 1).all()
    rng = np.random.default_rng(1)
    total = np.zeros(n)
    d0 = np.linalg.norm(env.goal - env.body.pos(), axis=1)
    seen_trunc = seen_term = False
    for t in range(120):
        before = env.body.state_vector().copy()
        a = np.sign(rng.standard_normal((n, 2))).astype(np.float32)
        obs, rew, done, info = env.step(a)
        assert np.float32(obs[:, 2]).tolist() == [np.float32(9.81)] * n
        for i in np.nonzero(done)[0]:
            if info["truncated"][i]:
                seen_trunc = True
                assert info["ep_l"][i] == 50
                assert np.all(env.body.v[i] == 0)
            else:
                seen_term = True
                assert info["terminated"][i]
        assert (env.elapsed[done] == 0).all()
    assert seen_trunc
    assert (env.n_resets >= env.n_full).all()
