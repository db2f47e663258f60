"""CUDA policy forward and GAE vs the torch / numpy oracle (through the C ABI)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import sb3_oracle

pytestmark = pytest.mark.gpu


def _flat(policy):
    return policy.flat_params().cuda().contiguous()


def _forward(lib, params, O, obs, eps=None):
    from mobrob_b200 import _lib

    n = obs.shape[0]
    act = torch.empty((n, 2), device="cuda")
    logp = torch.empty(n, device="cuda")
    val = torch.empty(n, device="cuda")
    _lib.check(lib.mr_policy_forward(params.data_ptr(), O, obs.data_ptr(),
                                     None if eps is None else eps.data_ptr(), act.data_ptr(),
                                     logp.data_ptr(), val.data_ptr(), n,
                                     torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return act.cpu(), logp.cpu(), val.cpu()


@pytest.mark.parametrize("env,O", [("point", 14), ("car", 26)])
def test_kat2_on_device(cuda_lib, golden_dir, env, O):
    kat = json.load(open(os.path.join(golden_dir, "kat2.json")))
    w = dict(np.load(os.path.join(golden_dir, f"{env}_policy.npz")))
    pol = sb3_oracle.MlpPolicyOracle(O).load_numpy(w)
    obs = torch.as_tensor(np.load(os.path.join(golden_dir, f"{env}_last_obs.npy"))).cuda()
    act, _, val = _forward(cuda_lib, _flat(pol), O, obs)
    np.testing.assert_allclose(act.numpy(), np.array(kat[env]["mu"]), rtol=1e-5)
    np.testing.assert_allclose(val.numpy(), np.array(kat[env]["v"]), rtol=1e-5)


@pytest.mark.parametrize("O,n,pretrained", [(14, 1, True), (14, 4099, True), (14, 1000, False), (26, 777, False)])
def test_policy_forward_sampled(cuda_lib, golden_dir, O, n, pretrained):
    torch.manual_seed(0)
    pol = sb3_oracle.MlpPolicyOracle(O)
    if pretrained:
        pol.load_numpy(dict(np.load(os.path.join(golden_dir, "point_policy.npz"))))
    else:
        with torch.no_grad():
            pol.log_std.copy_(torch.tensor([-0.3, 0.4]))
    obs = torch.randn(n, O) * 2.0
    eps = torch.randn(n, 2)
    with torch.no_grad():
        a_ref, v_ref, lp_ref = pol.forward_with_noise(obs, eps)
    act, logp, val = _forward(cuda_lib, _flat(pol), O, obs.cuda(), eps.cuda())
    scale = max(1.0, float(a_ref.abs().max()))
    assert float((act - a_ref).abs().max()) <= 1e-5 * scale
    np.testing.assert_allclose(val.numpy(), v_ref.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(logp.numpy(), lp_ref.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("T,N,lam", [(1, 3, 0.5), (64, 130, 0.5), (257, 64, 0.95), (4000, 2, 0.5)])
def test_gae_bit_exact(cuda_lib, T, N, lam):
    from mobrob_b200 import _lib

    rng = np.random.default_rng(T * 1000 + N)
    rew = (rng.standard_normal((T, N)) * 0.1 + (rng.random((T, N)) < 0.02) * 5.0).astype(np.float32)
    val = (rng.standard_normal((T, N)) * 2).astype(np.float32)
    starts = (rng.random((T, N)) < 0.03).astype(np.float32)
    last_val = rng.standard_normal(N).astype(np.float32)
    dones = rng.random(N) < 0.3
    adv_ref, ret_ref = sb3_oracle.gae_numpy(rew, val, starts, last_val, dones, 0.99, lam)
    d = lambda x: torch.as_tensor(x).cuda().contiguous()
    adv = torch.empty((T, N), device="cuda")
    ret = torch.empty((T, N), device="cuda")
    t_rew, t_val, t_st, t_lv, t_dn = d(rew), d(val), d(starts), d(last_val), d(dones.astype(np.uint8))
    _lib.check(cuda_lib.mr_gae(t_rew.data_ptr(), t_val.data_ptr(), t_st.data_ptr(),
                               t_lv.data_ptr(), t_dn.data_ptr(), 0.99, lam,
                               adv.data_ptr(), ret.data_ptr(), T, N,
                               torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(adv.cpu().numpy(), adv_ref)
    np.testing.assert_array_equal(ret.cpu().numpy(), ret_ref)
