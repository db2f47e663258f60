"""Reference random streams, restated with numpy (test infrastructure only).

What the reference draws, and where:

* ``EnvWrapper.seed``            wrapper.py:95-107   init_space -> PCG64(seed),
                                                      goal_space -> PCG64(seed + 1),
                                                      Engine._seed = seed
* ``Box.sample``  [gymnasium 0.28.1, call sites wrapper.py:126,191]
                                  Generator(PCG64(SeedSequence(seed))).uniform(low, high)
                                  .astype(float32)
* ``Engine.reset``               engine.py:1000-1021 ``_seed += 1; rs = RandomState(_seed)``
* ``Engine.build_layout``        engine.py:633-667   rejection sampling: robot xy then goal xy
* ``Engine.build_world_config``  engine.py:721-731   ``robot_rot = rs.uniform(0, 2*pi)``
* ``make_vec_env`` / ``VecEnv.seed`` [SB3 2.0.0, call site ppo.py:37-48]
                                  env rank i is first reset with ``seed + i``.
"""
from __future__ import annotations

import numpy as np

# engine.py:101 placements_extents, :110 robot_keepout, :175 goal_keepout
EXTENTS = (-2, -2, 2, 2)
ROBOT_KEEPOUT = 0.4
GOAL_KEEPOUT = 0.4
PLACEMENTS_MARGIN = 0.0


class BoxSampler:
    """gymnasium 0.28.1 ``Box`` restricted to bounded float32 boxes (seed/sample only)."""

    def __init__(self, low, high):
        self.low = np.asarray(low, dtype=np.float32)
        self.high = np.asarray(high, dtype=np.float32)
        self.rng = np.random.Generator(np.random.PCG64())

    def seed(self, seed):
        self.rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

    def sample(self) -> np.ndarray:
        s = self.rng.uniform(low=self.low, high=self.high, size=self.low.shape)
        return s.astype(np.float32)


def init_box():
    """MujocoGoalEnv.get_init_space, wrapper.py:250-256 (extents / 2)."""
    x0, y0, x1, y1 = EXTENTS
    return BoxSampler(np.array([x0, y0], np.float32) / 2, np.array([x1, y1], np.float32) / 2)


def goal_box():
    """MujocoGoalEnv.get_goal_space, wrapper.py:258-264."""
    x0, y0, x1, y1 = EXTENTS
    return BoxSampler(np.array([x0, y0], np.float32), np.array([x1, y1], np.float32))


def engine_heading(engine_seed: int) -> float:
    """Heading drawn by ``Engine.reset`` with ``RandomState(engine_seed)``.

    Follows engine.py:633-667 (layout rejection sampling; only the *count* of
    draws matters because set_pos / set_goal overwrite the positions) and
    engine.py:728-729 (the heading draw itself).
    """
    rs = np.random.RandomState(engine_seed & 0xFFFFFFFF)
    xmin, ymin, xmax, ymax = EXTENTS
    lo_r, hi_r = xmin + ROBOT_KEEPOUT, xmax - ROBOT_KEEPOUT
    lo_g, hi_g = xmin + GOAL_KEEPOUT, xmax - GOAL_KEEPOUT
    for _ in range(10000):
        robot = np.array([rs.uniform(lo_r, hi_r), rs.uniform(lo_r, hi_r)])
        ok = False
        for _ in range(100):
            goal = np.array([rs.uniform(lo_g, hi_g), rs.uniform(lo_g, hi_g)])
            dist = np.sqrt(np.sum(np.square(goal - robot)))
            if not dist < ROBOT_KEEPOUT + PLACEMENTS_MARGIN + GOAL_KEEPOUT:
                ok = True
                break
        if ok:
            break
    else:  # pragma: no cover - probability ~0
        raise RuntimeError("layout resampling failed")
    return float(rs.uniform(0, 2 * np.pi))


def pcg64_state_words(seed: int) -> np.ndarray:
    """(state_hi, state_lo, inc_hi, inc_lo) of ``PCG64(SeedSequence(seed))`` as uint64[4]."""
    st = np.random.PCG64(np.random.SeedSequence(seed)).state["state"]
    m = (1 << 64) - 1
    return np.array(
        [st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m], dtype=np.uint64
    )
