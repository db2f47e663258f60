"""Host-side seeding glue: turn the reference's integer seeds into the raw generator words
the device-side streams start from.

EnvWrapper.seed (src/mobrob/envs/wrapper.py:95-107) seeds ``init_space`` with ``seed`` and
``goal_space`` with ``seed + 1``; gymnasium 0.28.1 ``Space.seed`` builds
``Generator(PCG64(SeedSequence(seed)))``.  The SeedSequence hashing is numpy's; the device
continues the PCG64 stream from the words computed here (csrc/common.cuh: Pcg64).
[SB3] make_vec_env / VecEnv.seed give env rank i the seed ``seed + i``
(src/mobrob/rl_control/ppo.py:37-48).
"""
from __future__ import annotations

import numpy as np

_MASK = (1 << 64) - 1


def pcg64_words(seed: int) -> np.ndarray:
    st = np.random.PCG64(np.random.SeedSequence(int(seed))).state["state"]
    return np.array([st["state"] >> 64, st["state"] & _MASK, st["inc"] >> 64, st["inc"] & _MASK],
                    dtype=np.uint64)


def vec_env_streams(seed: int, n_envs: int, first_rank: int = 0):
    """(pcg_init [N,4] u64, pcg_goal [N,4] u64, engine_seed [N] i64) for ranks first_rank.."""
    words = np.stack([pcg64_words(seed + first_rank + i) for i in range(n_envs + 1)])
    pcg_init = np.ascontiguousarray(words[:-1])
    pcg_goal = np.ascontiguousarray(words[1:])
    engine_seed = (seed + first_rank + np.arange(n_envs)).astype(np.int64)
    return pcg_init, pcg_goal, engine_seed
