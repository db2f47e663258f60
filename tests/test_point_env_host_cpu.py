"""The point env's per-environment step / reset logic, compiled for the HOST, against the oracle stack (no GPU needed).

mobrob_b200/csrc/env_point.cuh (EnvWrapper.step / reward_fn / reached / reset, TimeLimit, Monitor bookkeeping,
DummyVecEnv auto-reset, and the device restatements of numpy's PCG64 and MT19937 streams in common.cuh) is
__host__ __device__: tests/host/point_env_host.cu drives the very functions point_step_kernel / point_reset_kernel
call, fed with the same seeded stream words GpuVecEnv.seed uploads.  Same checks as tests/test_env_parity_gpu.py:
done / truncation flags, reset counters and episode lengths bit exact, initial positions and goals bit exact,
observations / rewards / terminal observations to the north_star tolerance of 1e-5."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from mobrob_b200 import seeding
from oracle import point_oracle as po
from oracle.vec_oracle import GoalVecOracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("host") / "point_env_host")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "point_env_host.cu")])
    return exe


def _actions(rng, n, t):
    if t % 11 == 0:
        return rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    return (np.sign(rng.standard_normal((n, 2))) * rng.choice([1.0, 1.7], (n, 2))).astype(np.float32)


@pytest.mark.parametrize("seed,n,time_limit,T", [(0, 24, 1000, 400), (7, 17, 45, 300)])
def test_host_compiled_env_logic_matches_oracle_stack(harness, tmp_path, seed, n, time_limit, T):
    rng = np.random.default_rng(seed + 100)
    act = np.stack([_actions(rng, n, t) for t in range(T)])
    init, goal, eng = seeding.vec_env_streams(seed, n)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([n, T, time_limit, 1], np.int64).tofile(f)
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

    ora = GoalVecOracle(po.PointBody(n), seed=seed, time_limit=time_limit, terminate_on_goal=True)
    np.testing.assert_allclose(take(n * 14, np.float32).reshape(n, 14), ora.reset(), rtol=1e-5, atol=1e-6)
    n_done = n_trunc = 0
    for t in range(T):
        obs = take(n * 14, np.float32).reshape(n, 14)
        rew = take(n, np.float32)
        done = take(n, np.uint8).astype(bool)
        trunc = take(n, np.uint8).astype(bool)
        tobs = take(n * 14, np.float32).reshape(n, 14)
        ep_r = take(n, np.float64)
        ep_l = take(n, np.int32)
        o_ref, r_ref, d_ref, info = ora.step(act[t])
        np.testing.assert_array_equal(done, d_ref, err_msg=f"done flags, step {t}")
        np.testing.assert_array_equal(trunc, info["truncated"], err_msg=f"truncation flags, step {t}")
        np.testing.assert_allclose(obs, o_ref, rtol=1e-5, atol=1e-6, err_msg=f"obs, step {t}")
        np.testing.assert_allclose(rew, r_ref, rtol=1e-5, atol=1e-7, err_msg=f"reward, step {t}")
        for i in np.nonzero(d_ref)[0]:
            n_done += 1
            n_trunc += int(info["truncated"][i])
            assert ep_l[i] == info["ep_l"][i]
            assert abs(ep_r[i] - info["ep_r"][i]) <= 1e-5 * max(1.0, abs(info["ep_r"][i]))
            np.testing.assert_allclose(tobs[i], info["terminal_obs"][i], rtol=1e-5, atol=1e-6)
    counts = take(n * 2, np.int32).reshape(n, 2)
    assert off == raw.size
    np.testing.assert_array_equal(counts[:, 0], ora.n_resets)      # resets and FULL resets (robot re-placed): bit exact,
    np.testing.assert_array_equal(counts[:, 1], ora.n_full)        # i.e. every reached / time-out decision agreed
    assert n_done > 0 and (time_limit == 1000 or n_trunc > 0)
