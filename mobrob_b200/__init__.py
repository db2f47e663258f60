"""mobrob_b200 -- B200-native goal-conditioned PPO hot path of ZikangXiong/mobrob."""
__all__ = ["GpuVecEnv"]


def __getattr__(name):
    if name == "GpuVecEnv":
        from .vec_env import GpuVecEnv
        return GpuVecEnv
    raise AttributeError(name)
