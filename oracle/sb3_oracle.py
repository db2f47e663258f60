"""Restatement of the stable-baselines3 2.0.0 PPO arithmetic used by mobrob, with
torch-CPU autograd as the arithmetic (test infrastructure only).

SB3 is an un-vendored dependency (``stable-baselines3[extra]~=2.0.0``,
requirements.txt:9; the shipped zips say 2.0.0) whose source is not on this
machine; what follows restates its published algorithm (SURVEY.md section 8a
rows S3-S8, appendix A.3-A.5) and is anchored on the reference's call sites
  * PPO(...) construction      src/mobrob/rl_control/ppo.py:50-59
  * hyper-parameters           data/configs/point-ppo.yaml:10-18
  * learn / save               examples/train.py:42-49
  * predict                    examples/control.py:39
and pinned by KAT-2: the shipped ``policy.pth`` evaluated on the shipped
``_last_obs`` (tests/golden/kat2.json).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

PARAM_ORDER = [
    "log_std",
    "mlp_extractor.policy_net.0.weight",
    "mlp_extractor.policy_net.0.bias",
    "mlp_extractor.policy_net.2.weight",
    "mlp_extractor.policy_net.2.bias",
    "mlp_extractor.value_net.0.weight",
    "mlp_extractor.value_net.0.bias",
    "mlp_extractor.value_net.2.weight",
    "mlp_extractor.value_net.2.bias",
    "action_net.weight",
    "action_net.bias",
    "value_net.weight",
    "value_net.bias",
]


class _Extractor(nn.Module):
    def __init__(self, obs_dim, hidden):
        super().__init__()
        self.policy_net = nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh())
        self.value_net = nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh())


class MlpPolicyOracle(nn.Module):
    """ActorCriticPolicy("MlpPolicy") with net_arch pi=[64,64], vf=[64,64], tanh (S3)."""

    def __init__(self, obs_dim: int, act_dim: int = 2, hidden: int = 64, ortho_init: bool = True):
        super().__init__()
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.log_std = nn.Parameter(torch.zeros(act_dim))
        self.mlp_extractor = _Extractor(obs_dim, hidden)
        self.action_net = nn.Linear(hidden, act_dim)
        self.value_net = nn.Linear(hidden, 1)
        if ortho_init:
            for mod, gain in ((self.mlp_extractor, math.sqrt(2)), (self.action_net, 0.01), (self.value_net, 1.0)):
                for m in mod.modules():
                    if isinstance(m, nn.Linear):
                        nn.init.orthogonal_(m.weight, gain=gain)
                        m.bias.data.fill_(0.0)

    def load_numpy(self, arrays: dict):
        self.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in arrays.items()})
        return self

    def latent(self, obs):
        return self.mlp_extractor.policy_net(obs), self.mlp_extractor.value_net(obs)

    def mean_value(self, obs):
        lp, lv = self.latent(obs)
        return self.action_net(lp), self.value_net(lv).flatten()

    def forward_with_noise(self, obs, eps):
        """collect_rollouts forward: actions = mu + exp(log_std) * eps (unclipped)."""
        mu, v = self.mean_value(obs)
        std = torch.ones_like(mu) * self.log_std.exp()
        actions = mu + std * eps
        return actions, v, self.log_prob(mu, actions)

    def log_prob(self, mu, actions):
        dist = torch.distributions.Normal(mu, torch.ones_like(mu) * self.log_std.exp())
        return dist.log_prob(actions).sum(dim=1)

    def evaluate_actions(self, obs, actions):
        mu, v = self.mean_value(obs)
        dist = torch.distributions.Normal(mu, torch.ones_like(mu) * self.log_std.exp())
        return v, dist.log_prob(actions).sum(dim=1), dist.entropy().sum(dim=1)

    def predict_deterministic(self, obs):
        mu, _ = self.mean_value(obs)
        return torch.clamp(mu, -1.0, 1.0)

    def flat_params(self):
        sd = dict(self.named_parameters())
        return torch.cat([sd[k].detach().reshape(-1) for k in PARAM_ORDER])

    def flat_grads(self):
        sd = dict(self.named_parameters())
        return torch.cat([sd[k].grad.reshape(-1) for k in PARAM_ORDER])


def gae_numpy(rewards, values, episode_starts, last_values, dones, gamma, gae_lambda):
    """RolloutBuffer.compute_returns_and_advantage (S5), statement for statement.

    ``dones`` must be the bool array the VecEnv returned: ``1.0 - dones`` is then float64,
    which makes ``last_gae_lam`` a float64 array for the whole loop (numpy promotion),
    while the buffers are float32.
    """
    rewards = np.asarray(rewards, np.float32)
    values = np.asarray(values, np.float32)
    episode_starts = np.asarray(episode_starts, np.float32)
    last_values = np.asarray(last_values, np.float32).flatten()
    dones = np.asarray(dones, dtype=bool)
    gamma, gae_lambda = float(gamma), float(gae_lambda)
    buffer_size = rewards.shape[0]
    advantages = np.zeros_like(rewards)
    last_gae_lam = 0
    for step in reversed(range(buffer_size)):
        if step == buffer_size - 1:
            next_non_terminal = 1.0 - dones
            next_values = last_values
        else:
            next_non_terminal = 1.0 - episode_starts[step + 1]
            next_values = values[step + 1]
        delta = rewards[step] + gamma * next_values * next_non_terminal - values[step]
        last_gae_lam = delta + gamma * gae_lambda * next_non_terminal * last_gae_lam
        advantages[step] = last_gae_lam
    returns = advantages + values
    return advantages, returns


def ppo_loss(policy, obs, actions, old_log_prob, advantages, returns, clip_range=0.2,
             ent_coef=0.05, vf_coef=0.5, normalize_advantage=True):
    """One minibatch of PPO.train (S7).  Returns (loss, stats dict of python floats)."""
    values, log_prob, entropy = policy.evaluate_actions(obs, actions)
    adv = advantages
    if normalize_advantage and len(adv) > 1:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    ratio = torch.exp(log_prob - old_log_prob)
    pl1 = adv * ratio
    pl2 = adv * torch.clamp(ratio, 1 - clip_range, 1 + clip_range)
    policy_loss = -torch.min(pl1, pl2).mean()
    clip_fraction = torch.mean((torch.abs(ratio - 1) > clip_range).float()).item()
    value_loss = torch.nn.functional.mse_loss(returns, values)
    entropy_loss = -torch.mean(entropy)
    loss = policy_loss + ent_coef * entropy_loss + vf_coef * value_loss
    with torch.no_grad():
        log_ratio = log_prob - old_log_prob
        approx_kl = torch.mean((torch.exp(log_ratio) - 1) - log_ratio).item()
    stats = dict(policy_loss=policy_loss.item(), value_loss=value_loss.item(),
                 entropy_loss=entropy_loss.item(), loss=loss.item(),
                 clip_fraction=clip_fraction, approx_kl=approx_kl)
    return loss, stats


def make_adam(policy, lr=3e-4):
    sd = dict(policy.named_parameters())
    return torch.optim.Adam([sd[k] for k in PARAM_ORDER], lr=lr, eps=1e-5)


def train_minibatch(policy, optimizer, batch, max_grad_norm=0.5, **loss_kw):
    """zero_grad, backward, clip_grad_norm_, Adam.step; returns (stats, flat grad before clip)."""
    loss, stats = ppo_loss(policy, *batch, **loss_kw)
    optimizer.zero_grad()
    loss.backward()
    g = policy.flat_grads().clone()
    total = torch.nn.utils.clip_grad_norm_(policy.parameters(), max_grad_norm)
    stats["grad_norm"] = float(total)
    optimizer.step()
    return stats, g


def flatten_env_major(x):
    """RolloutBuffer.swap_and_flatten: (T, N, ...) -> (N*T, ...), index = n*T + t (S6)."""
    x = np.asarray(x)
    shape = x.shape
    if len(shape) < 3:
        shape = (*shape, 1)
    return x.swapaxes(0, 1).reshape(shape[0] * shape[1], *shape[2:])


class RolloutOracle:
    """collect_rollouts (S4) over a GoalVecOracle with host-supplied action noise."""

    def __init__(self, venv, policy, n_steps, gamma=0.99, gae_lambda=0.5):
        self.venv, self.policy, self.T = venv, policy, n_steps
        self.gamma, self.lam = gamma, gae_lambda
        self.last_obs = venv.reset()
        self.last_starts = np.ones(venv.n, dtype=bool)
        self.ep_infos = []

    @torch.no_grad()
    def collect(self, eps):
        """eps: (T, N, A) float32 standard-normal draws.  Returns dict of (T, N, ...) arrays."""
        T, N = self.T, self.venv.n
        O, A = self.policy.obs_dim, self.policy.act_dim
        buf = dict(obs=np.zeros((T, N, O), np.float32), actions=np.zeros((T, N, A), np.float32),
                   rewards=np.zeros((T, N), np.float32), episode_starts=np.zeros((T, N), np.float32),
                   values=np.zeros((T, N), np.float32), log_probs=np.zeros((T, N), np.float32))
        for t in range(T):
            obs_t = torch.as_tensor(self.last_obs)
            actions, values, logp = self.policy.forward_with_noise(obs_t, torch.as_tensor(eps[t]))
            actions = actions.numpy()
            clipped = np.clip(actions, -1.0, 1.0)
            new_obs, rewards, dones, infos = self.venv.step(clipped)
            for i in np.nonzero(dones)[0]:
                self.ep_infos.append((float(infos["ep_r"][i]), int(infos["ep_l"][i])))
            trunc = dones & infos["truncated"]
            if trunc.any():  # timeout bootstrap
                _, tv = self.policy.mean_value(torch.as_tensor(infos["terminal_obs"][trunc]))
                rewards[trunc] += self.gamma * tv.numpy()
            buf["obs"][t] = self.last_obs
            buf["actions"][t] = actions
            buf["rewards"][t] = rewards
            buf["episode_starts"][t] = self.last_starts
            buf["values"][t] = values.numpy()
            buf["log_probs"][t] = logp.numpy()
            self.last_obs = new_obs
            self.last_starts = dones
        _, last_values = self.policy.mean_value(torch.as_tensor(new_obs))
        adv, ret = gae_numpy(buf["rewards"], buf["values"], buf["episode_starts"],
                             last_values.numpy(), dones, self.gamma, self.lam)
        buf["advantages"], buf["returns"] = adv, ret
        buf["last_values"], buf["last_dones"] = last_values.numpy(), dones.copy()
        return buf


def train_epochs(policy, optimizer, buf, n_epochs, batch_size, perms=None, rng=None, **loss_kw):
    """PPO.train (S7) over a collected buffer.  perms: optional list of index arrays per epoch."""
    flat = {k: torch.as_tensor(flatten_env_major(buf[k])) for k in
            ("obs", "actions", "log_probs", "advantages", "returns")}
    n = flat["obs"].shape[0]
    all_stats = []
    for ep in range(n_epochs):
        idx = perms[ep] if perms is not None else (rng or np.random).permutation(n)
        for s in range(0, n, batch_size):
            b = torch.as_tensor(np.asarray(idx[s:s + batch_size], dtype=np.int64))
            batch = (flat["obs"][b], flat["actions"][b], flat["log_probs"][b].flatten(),
                     flat["advantages"][b].flatten(), flat["returns"][b].flatten())
            stats, _ = train_minibatch(policy, optimizer, batch, **loss_kw)
            all_stats.append(stats)
    return all_stats
