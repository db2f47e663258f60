"""The env-step kernel's integrator, compiled for the HOST, against the oracle (no GPU needed).

mobrob_b200/csrc/point_dyn.cuh is __host__ __device__: tests/host/point_dyn_host.cu calls the very functions the
CUDA kernels call.  Checks the reformulated physics (linear form of the angular acceleration, loop carried on
the heading increment, small-angle Taylor rotation, constant-bank sincos, the cold paths for |omega| >= 5 rad/s and
|heading| >= 1e5 rad) to the north_star tolerance of 1e-5 on body state and observations."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import point_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-5


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("host") / "point_dyn_host")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "point_dyn_host.cu")])
    return exe


def _run(exe, tmp_path, state, goal, act):
    n, T = state.shape[0], act.shape[0]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([n, T], np.int64).tofile(f)
        state.astype(np.float64).tofile(f)
        goal.astype(np.float32).tofile(f)
        act.astype(np.float32).tofile(f)
    subprocess.check_call([exe, fin, fout])
    raw = np.fromfile(fout, np.uint8)
    st = raw[:n * 48].view(np.float64).reshape(n, 6)
    obs = raw[n * 48:].view(np.float32).reshape(T, n, 14)
    return st, obs


def _oracle(state, goal, act):
    n, T = state.shape[0], act.shape[0]
    body = po.PointBody(n)
    body.q[:] = state[:, 0:3]          # psi0 = 0, body_xy = 0: the joint frame is the world frame
    body.v[:] = state[:, 3:6]
    obs = np.zeros((T, n, 14), np.float32)
    for t in range(T):
        body.step(act[t])
        obs[t] = body.obs(goal)
    return np.concatenate([body.q, body.v], axis=1), obs


@pytest.mark.parametrize("case", ["bang-bang", "fast-spin", "huge-heading"])
def test_host_compiled_integrator_matches_oracle(harness, tmp_path, case):
    rng = np.random.default_rng({"bang-bang": 1, "fast-spin": 2, "huge-heading": 3}[case])
    n, T = 64, 200 if case == "bang-bang" else 12
    state = np.zeros((n, 6))
    state[:, 0:2] = rng.uniform(-1, 1, (n, 2))
    state[:, 2] = rng.uniform(0, 2 * np.pi, n)
    if case == "fast-spin":
        state[:, 3:5] = rng.uniform(-2, 2, (n, 2))
        state[:, 5] = np.where(np.arange(n) % 2 == 0, rng.uniform(-60, 60, n), rng.uniform(4.9, 5.1, n))
    if case == "huge-heading":
        state[:, 2] = rng.uniform(-3e5, 3e5, n)
        state[:, 3:6] = rng.uniform(-1, 1, (n, 3))
    goal = rng.uniform(-2, 2, (n, 2)).astype(np.float32)
    hold = np.zeros((n, 2), np.int64)
    a = np.zeros((n, 2), np.float32)
    act = np.zeros((T, n, 2), np.float32)
    for t in range(T):   # bang-bang held for U{1..50} steps, some out-of-range values to exercise the clip
        renew = hold <= 0
        a = np.where(renew, np.sign(rng.standard_normal((n, 2))) * rng.choice([1.0, 1.7], (n, 2)), a).astype(np.float32)
        hold = np.where(renew, rng.integers(1, 51, (n, 2)), hold) - 1
        act[t] = a
    st, obs = _run(harness, tmp_path, state, goal, act)
    st_ref, obs_ref = _oracle(state, goal, act)
    err = np.abs(st - st_ref) / np.maximum(np.abs(st_ref), 1.0)
    assert err.max() < RTOL, err.max()
    np.testing.assert_allclose(obs, obs_ref, rtol=RTOL, atol=2e-6)
    assert np.abs(st_ref[:, 0:2] - state[:, 0:2]).max() > 0.05   # the robots really moved
