"""mr_host_permutation (csrc/host_perm.cu): the host-side index stream of RolloutBuffer.get."""
import numpy as np

from mobrob_b200 import _lib


def _perm(seed, stream, n):
    out = np.full(n, -1, dtype=np.int64)
    _lib.check(_lib.load().mr_host_permutation(seed, stream, n, out.ctypes.data))
    return out


def test_is_a_permutation_for_ragged_sizes():
    for n in (0, 1, 2, 7, 100, 8192, 8193, 65537, 4096 * 296):
        p = _perm(3, 11, n)
        assert np.array_equal(np.sort(p), np.arange(n))


def test_pure_function_of_seed_and_stream():
    a, b = _perm(1, 5, 50000), _perm(1, 5, 50000)
    assert np.array_equal(a, b)
    assert (a != _perm(1, 6, 50000)).mean() > 0.99
    assert (a != _perm(2, 5, 50000)).mean() > 0.99


def test_uniform_over_small_permutations():
    # every (position, value) pair equally likely: chi-square over 20 000 draws of n = 10
    cnt = np.zeros((10, 10))
    for s in range(20000):
        cnt[np.arange(10), _perm(9, s, 10)] += 1
    chi2 = ((cnt - 2000.0) ** 2 / 2000.0).sum()
    assert chi2 < 140.0  # 81 degrees of freedom, p ~ 1e-4


def test_bucket_path_spreads_values_evenly():
    n, acc = 100000, np.zeros(16)
    for s in range(50):
        p = _perm(4, s, n)
        acc += np.bincount(np.argsort(p)[:2000] * 16 // n, minlength=16)
    frac = acc / acc.sum()
    assert np.abs(frac - 1 / 16).max() < 0.006


def test_rejects_bad_arguments():
    lib = _lib.load()
    assert lib.mr_host_permutation(0, 0, 8, None) != 0
    out = np.zeros(4, dtype=np.int64)
    assert lib.mr_host_permutation(0, 0, -1, out.ctypes.data) != 0
