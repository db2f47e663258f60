"""Name shim: both vec-env classes of src/mobrob/rl_control/ppo.py:2-3 are the HBM-resident GpuVecEnv."""
from mobrob_b200.vec_env import GpuVecEnv

DummyVecEnv = SubprocVecEnv = VecEnv = GpuVecEnv
