"""The car kernels' physics, compiled for the HOST, against the oracle (no GPU needed).

mobrob_b200/csrc/car_dyn.cuh is __host__ __device__ (its contact scratch is a plain array on the host):
tests/host/car_dyn_host.cu calls the very substep / sensor routines car_step_kernel calls.  Checks the gyrostat
dynamics (contact-free, north_star tolerance 1e-5 -- in fact ~1e-12) and the closed-form Delassus matrix with the
warm-started projected Gauss-Seidel sweeps against the oracle's matrix-free form of the same algorithm (1e-4, the
bound tests/test_car_gpu.py holds the GPU to)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import car_oracle as co

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("host") / "car_dyn_host")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "car_dyn_host.cu")])
    return exe


def _pack(b):
    """car::State order: p3 quat4 v3 w3 th2 s2 qb4 wb3."""
    return np.concatenate([b.p, b.quat, b.v, b.w, b.th, b.s, b.qb, b.wb], axis=1)


def _run(exe, tmp_path, state, goal, act, contacts):
    n, T = state.shape[0], act.shape[0]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([n, T, int(contacts)], np.int64).tofile(f)
        state.astype(np.float64).tofile(f)
        goal.astype(np.float32).tofile(f)
        act.astype(np.float32).tofile(f)
    subprocess.check_call([exe, fin, fout])
    raw = np.fromfile(fout, np.uint8)
    st = raw[:n * 24 * 8].view(np.float64).reshape(n, 24)
    obs = raw[n * 24 * 8:].view(np.float32).reshape(T, n, 26)
    return st, obs


def _bodies(n, seed, settle):
    rng = np.random.default_rng(seed)
    b = co.CarBody(n)
    for i in range(n):
        b.full_reset(i, rng.uniform(-1, 1, 2), rng.uniform(0, 2 * np.pi))
    for _ in range(settle):
        b.step(np.sign(rng.standard_normal((n, 2))))
    return b, rng


def test_host_compiled_contact_free_dynamics_match_oracle(harness, tmp_path):
    n, T = 6, 12
    b, rng = _bodies(n, 0, 0)
    b.contacts_enabled = False
    b.w[:] = rng.uniform(-2, 2, (n, 3)); b.v[:] = rng.uniform(-1, 1, (n, 3)); b.s[:] = rng.uniform(-30, 30, (n, 2))
    b.wb[:] = rng.uniform(-5, 5, (n, 3))
    goal = rng.uniform(-2, 2, (n, 2)).astype(np.float32)
    act = np.sign(rng.standard_normal((T, n, 2))).astype(np.float32) * 1.5
    st, obs = _run(harness, tmp_path, _pack(b), goal, act, contacts=False)
    ref_obs = np.zeros((T, n, 26), np.float32)
    for t in range(T):
        b.step(act[t])
        ref_obs[t] = b.obs(goal)
    ref = _pack(b)
    assert np.max(np.abs(st - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-9
    np.testing.assert_allclose(obs, ref_obs, rtol=1e-5, atol=2e-6)
    assert np.abs(ref[:, 14:16]).max() > 5.0   # the wheels really spun


def test_host_compiled_contacts_match_oracle(harness, tmp_path):
    """Driving on the floor: closed-form A = J M^-1 J^T + warm-started sweeps (kernel code) vs matrix-free (oracle)."""
    n, T = 5, 8
    b, rng = _bodies(n, 1, 3)
    goal = rng.uniform(-2, 2, (n, 2)).astype(np.float32)
    act = np.sign(rng.standard_normal((T, n, 2))).astype(np.float32)
    st, obs = _run(harness, tmp_path, _pack(b), goal, act, contacts=True)
    ref_obs = np.zeros((T, n, 26), np.float32)
    for t in range(T):
        b.step(act[t])
        ref_obs[t] = b.obs(goal)
    ref = _pack(b)
    err = np.abs(st - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < 1e-4, err.max()
    np.testing.assert_allclose(obs, ref_obs, rtol=1e-4, atol=1e-4)
    assert np.all(np.abs(ref[:, 2] - 0.1) < 5e-3)          # resting on the wheels
    assert abs(float(np.median(ref_obs[:, :, 2])) - 9.81) < 0.5   # accelerometer z ~ g: the contact forces carry the car
