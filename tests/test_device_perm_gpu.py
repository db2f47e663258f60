"""mr_device_permutation: the keyed on-device index stream used by PPO(permutation="device")."""
import numpy as np
import pytest
import torch

from mobrob_b200 import _lib

pytestmark = pytest.mark.gpu


def _perm(seed, stream, n):
    out = torch.full((n,), -1, dtype=torch.int64, device="cuda")
    _lib.check(_lib.load().mr_device_permutation(seed, stream, n, out.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream))
    return out.cpu().numpy()


def test_is_a_permutation_for_ragged_sizes(cuda_lib):
    for n in (1, 2, 3, 100, 4096, 4097, 65536, 4096 * 296):
        assert np.array_equal(np.sort(_perm(3, 7, n)), np.arange(n))


def test_pure_function_of_seed_and_stream(cuda_lib):
    a = _perm(1, 5, 100000)
    assert np.array_equal(a, _perm(1, 5, 100000))
    assert (a != _perm(1, 6, 100000)).mean() > 0.99
    assert (a != _perm(2, 5, 100000)).mean() > 0.99


def test_minibatches_are_well_mixed(cuda_lib):
    # env-major sample ids n*T + t: every minibatch should draw evenly from all parts of the buffer
    n, mb, bins = 4096 * 296, 18944, 16
    worst = 0.0
    for s in range(4):
        p = _perm(0, s, n)
        for k in (0, 17, 63):
            h = np.bincount(p[k * mb:(k + 1) * mb] * bins // n, minlength=bins)
            worst = max(worst, np.abs(h / mb - 1 / bins).max())
        # no short-range structure: neighbours in the stream are far apart in the buffer
        assert np.median(np.abs(np.diff(p[:10000]))) > n / 8
    assert worst < 0.01
