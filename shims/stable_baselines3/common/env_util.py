"""Name shim: ``make_vec_env`` (src/mobrob/rl_control/ppo.py:1): n_envs environments in ONE GpuVecEnv."""
from mobrob_b200.vec_env import GpuVecEnv


def make_vec_env(env_id, n_envs=1, seed=None, env_kwargs=None, vec_env_cls=None, **_ignored):
    kw = dict(env_kwargs or {})
    name = kw.pop("env_name", env_id if isinstance(env_id, str) else None)
    if name is None:
        raise TypeError("make_vec_env shim: pass env_kwargs['env_name'] (the reference passes get_env as env_id)")
    return GpuVecEnv(name, n_envs, seed=seed, time_limit=kw.get("time_limit"),
                     terminate_on_goal=kw.get("terminate_on_goal", False))
