"""The car env's per-environment step / reset logic, compiled for the HOST, against the oracle stack (no GPU needed):
mobrob_b200/csrc/env_car.cuh on car_dyn.cuh, the functions car_step_kernel / car_reset_kernel call, driven by
tests/host/car_env_host.cu with the stream words GpuVecEnv.seed uploads.  Bounds as in tests/test_car_gpu.py: flags,
counters, start positions and goals exact; observations and rewards 1e-4 with the floor contacts."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from mobrob_b200 import seeding
from oracle import car_oracle as co
from oracle.vec_oracle import GoalVecOracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("host") / "car_env_host")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "car_env_host.cu")])
    return exe


def test_host_compiled_car_env_matches_oracle_stack(harness, tmp_path):
    n, seed, tl, T = 6, 4, 12, 30
    rng = np.random.default_rng(1)
    act = np.sign(rng.standard_normal((T, n, 2))).astype(np.float32)
    init, goal, eng = seeding.vec_env_streams(seed, n)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([n, T, tl, 1], np.int64).tofile(f)
        init.tofile(f); goal.tofile(f); eng.tofile(f); act.tofile(f)
    subprocess.check_call([harness, fin, fout])
    raw = np.fromfile(fout, np.uint8)
    off = 0

    def take(count, dtype):
        nonlocal off
        nbytes = count * np.dtype(dtype).itemsize
        out = raw[off:off + nbytes].view(dtype)
        off += nbytes
        return out

    ora = GoalVecOracle(co.CarBody(n), seed=seed, time_limit=tl, terminate_on_goal=True)
    np.testing.assert_allclose(take(n * 26, np.float32).reshape(n, 26), ora.reset(), rtol=1e-5, atol=2e-6)
    n_done = 0
    for t in range(T):
        obs = take(n * 26, np.float32).reshape(n, 26)
        rew = take(n, np.float32)
        done = take(n, np.uint8).astype(bool)
        trunc = take(n, np.uint8).astype(bool)
        tobs = take(n * 26, np.float32).reshape(n, 26)
        take(n, np.float64)
        ep_l = take(n, np.int32)
        o_ref, r_ref, d_ref, info = ora.step(act[t])
        np.testing.assert_array_equal(done, d_ref, err_msg=f"done flags, step {t}")
        np.testing.assert_array_equal(trunc, info["truncated"])
        np.testing.assert_allclose(obs, o_ref, rtol=1e-4, atol=1e-4, err_msg=f"obs, step {t}")
        np.testing.assert_allclose(rew, r_ref, rtol=1e-4, atol=1e-6)
        for i in np.nonzero(d_ref)[0]:
            n_done += 1
            assert ep_l[i] == info["ep_l"][i]
            np.testing.assert_allclose(tobs[i], info["terminal_obs"][i], rtol=1e-4, atol=1e-4)
    counts = take(n * 2, np.int32).reshape(n, 2)
    assert off == raw.size
    np.testing.assert_array_equal(counts[:, 0], ora.n_resets)
    np.testing.assert_array_equal(counts[:, 1], ora.n_full)
    assert n_done >= n   # every env hit the 12-step limit at least once: full resets exercised
