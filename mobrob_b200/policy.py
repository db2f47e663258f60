"""MlpPolicy: host-side mirror of SB3's ``ActorCriticPolicy`` ("MlpPolicy", separate 64-64 tanh
towers) whose 13 parameters are views into ONE flat CUDA vector -- the vector the kernels read.

``state_dict()`` / ``load_state_dict()`` use the reference's names and order, so
``data/policies/*.zip:policy.pth`` loads directly and examples/train.py's
``ppo.policy.load_state_dict(PPO.load(...).policy.state_dict())`` (train.py:30-33) works.
Every forward runs in libmobrob_b200 (mr_policy_forward); there is no torch forward here.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import _lib

HID = 64

PARAM_SHAPES = lambda O, A=2: OrderedDict([  # noqa: E731  (state-dict order of the shipped zips)
    ("log_std", (A,)),
    ("mlp_extractor.policy_net.0.weight", (HID, O)),
    ("mlp_extractor.policy_net.0.bias", (HID,)),
    ("mlp_extractor.policy_net.2.weight", (HID, HID)),
    ("mlp_extractor.policy_net.2.bias", (HID,)),
    ("mlp_extractor.value_net.0.weight", (HID, O)),
    ("mlp_extractor.value_net.0.bias", (HID,)),
    ("mlp_extractor.value_net.2.weight", (HID, HID)),
    ("mlp_extractor.value_net.2.bias", (HID,)),
    ("action_net.weight", (A, HID)),
    ("action_net.bias", (A,)),
    ("value_net.weight", (1, HID)),
    ("value_net.bias", (1,)),
])


def sb3_initial_state_dict(obs_dim: int, act_dim: int = 2, log_std_init: float = 0.0):
    """Fresh-policy initialisation in SB3's construction order on the torch CPU generator:
    default nn.Linear init for pi.0, pi.2, vf.0, vf.2, action_net, value_net, then
    orthogonal_(gain sqrt2 / sqrt2 / 0.01 / 1) with zero biases (ActorCriticPolicy._build)."""
    pi = [nn.Linear(obs_dim, HID), nn.Linear(HID, HID)]
    vf = [nn.Linear(obs_dim, HID), nn.Linear(HID, HID)]
    action_net = nn.Linear(HID, act_dim)
    value_net = nn.Linear(HID, 1)
    for mods, gain in ((pi + vf, math.sqrt(2)), ([action_net], 0.01), ([value_net], 1.0)):
        for m in mods:
            nn.init.orthogonal_(m.weight, gain=gain)
            m.bias.data.fill_(0.0)
    sd = OrderedDict()
    sd["log_std"] = torch.ones(act_dim) * log_std_init
    for name, mods in (("policy_net", pi), ("value_net", vf)):
        for i, m in zip((0, 2), mods):
            sd[f"mlp_extractor.{name}.{i}.weight"] = m.weight.data
            sd[f"mlp_extractor.{name}.{i}.bias"] = m.bias.data
    sd["action_net.weight"], sd["action_net.bias"] = action_net.weight.data, action_net.bias.data
    sd["value_net.weight"], sd["value_net.bias"] = value_net.weight.data, value_net.bias.data
    return OrderedDict((k, sd[k]) for k in PARAM_SHAPES(obs_dim, act_dim))


class MlpPolicy(nn.Module):
    def __init__(self, obs_dim: int, act_dim: int = 2, device=None, flat: torch.Tensor | None = None,
                 init: bool = True):
        super().__init__()
        assert act_dim == 2, "point and car have two actuators"
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        n = int(self.lib.mr_ppo_num_params(obs_dim))
        self.flat = flat if flat is not None else torch.zeros(n, dtype=torch.float32, device=self.device)
        assert self.flat.numel() == n and self.flat.is_cuda
        self._names = []
        off = 0
        for name, shape in PARAM_SHAPES(obs_dim, act_dim).items():
            cnt = int(np.prod(shape))
            p = nn.Parameter(self.flat[off:off + cnt].view(shape), requires_grad=False)
            self.register_parameter(name.replace(".", "__"), p)
            self._names.append(name)
            off += cnt
        assert off == n
        if init:
            self.load_state_dict(sb3_initial_state_dict(obs_dim, act_dim))

    # reference names (with dots) in and out
    def state_dict(self, *args, **kwargs):
        return OrderedDict((n, getattr(self, n.replace(".", "__")).data) for n in self._names)

    def load_state_dict(self, state_dict, strict: bool = True):
        missing = [n for n in self._names if n not in state_dict]
        extra = [k for k in state_dict if k not in self._names]
        if strict and (missing or extra):
            raise RuntimeError(f"state_dict mismatch: missing {missing}, unexpected {extra}")
        with torch.no_grad():
            for n in self._names:
                if n in state_dict:
                    getattr(self, n.replace(".", "__")).data.copy_(torch.as_tensor(state_dict[n]))
        return self

    def set_training_mode(self, mode: bool):
        return self

    # -- device forwards ------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def forward_tensor(self, obs: torch.Tensor, eps: torch.Tensor | None, act=None, logp=None, val=None):
        """obs [n, O] cuda f32; eps [n, 2] or None (deterministic).  Returns (act, logp, val)."""
        n = obs.shape[0]
        f32 = dict(dtype=torch.float32, device=self.device)
        act = act if act is not None else torch.empty((n, 2), **f32)
        logp = logp if logp is not None else torch.empty(n, **f32)
        val = val if val is not None else torch.empty(n, **f32)
        _lib.check(self.lib.mr_policy_forward(self.flat.data_ptr(), self.obs_dim, obs.data_ptr(),
                                              None if eps is None else eps.data_ptr(), act.data_ptr(),
                                              logp.data_ptr(), val.data_ptr(), n, self._stream()))
        return act, logp, val

    def predict(self, observation, state=None, episode_start=None, deterministic: bool = False):
        """SB3 ``policy.predict``: numpy in/out, action clipped to the Box (control.py:39)."""
        obs = np.asarray(observation, dtype=np.float32)
        single = obs.ndim == 1
        o = torch.as_tensor(obs.reshape(-1, self.obs_dim)).to(self.device).contiguous()
        eps = None if deterministic else torch.randn((o.shape[0], 2), device=self.device)
        act, _, _ = self.forward_tensor(o, eps)
        a = np.clip(act.cpu().numpy(), -1.0, 1.0)
        return (a[0] if single else a), state

    def predict_values(self, obs: torch.Tensor):
        _, _, v = self.forward_tensor(obs.to(self.device).contiguous(), None)
        return v
