"""Restatement of the reference's vectorised goal environment stack
``DummyVecEnv(Monitor(TimeLimit(EnvWrapper)))`` (test infrastructure only).

Follows
  * EnvWrapper.step / reward_fn / reached / reset / seed
        src/mobrob/envs/wrapper.py:95-107, 137-207
  * MujocoGoalEnv spaces, get_pos       wrapper.py:250-273
  * PointEnv.set_pos / CarEnv.set_pos   wrapper.py:301-305, 320-326
  * get_env (+ TimeLimit)               wrapper.py:549-571
  * PPOCtrl make_vec_env call           src/mobrob/rl_control/ppo.py:37-48
  * [gymnasium 0.28.1] TimeLimit, [SB3 2.0.0] Monitor / DummyVecEnv.step_wait
        (SURVEY.md appendix A.1, A.2 -- source not on this machine)
"""
from __future__ import annotations

import numpy as np

from . import ref_rng

REACH_RADIUS = 0.3  # wrapper.py:203
REACH_BONUS = 5.0  # wrapper.py:151-152


def dist2(ax, ay, bx, by):
    dx = ax - bx
    dy = ay - by
    return np.sqrt(dx * dx + dy * dy)


def extend_obs(body, base, goal, observe):
    """Engine.obs() with the optional keys (src/mobrob/envs/mujoco_robots/robots/engine.py:1179-1180 goal_dist =
    exp(-dist_goal()), engine.py:1243-1248 qpos / qvel / ctrl), flattened in sorted key order (engine.py:1253-1259):
    accelerometer [ballangvel_rear ballquat_rear] ctrl goal_compass goal_dist gyro magnetometer qpos qvel velocimeter.
    ``base`` is the default row (body.obs); ``observe`` the config dict (observe_goal_dist / qpos / qvel / ctrl).
    Parity unpinned beyond the restated source: no shipped artefact was produced with these keys switched on."""
    observe = observe or {}
    if not any(observe.get(k) for k in ("observe_goal_dist", "observe_qpos", "observe_qvel", "observe_ctrl")):
        return base
    pre, n_base = body.obs_pre, base.shape[1]
    parts = [base[:, :pre]]
    if observe.get("observe_ctrl"):
        parts.append(body.ctrl)
    parts.append(base[:, pre:pre + 2])
    if observe.get("observe_goal_dist"):
        p = body.pos()
        g = np.asarray(goal, dtype=np.float64)
        parts.append(np.exp(-dist2(g[:, 0], g[:, 1], p[:, 0], p[:, 1]))[:, None])
    parts.append(base[:, pre + 2:n_base - 3])
    if observe.get("observe_qpos"):
        parts.append(body.qpos())
    if observe.get("observe_qvel"):
        parts.append(body.qvel())
    parts.append(base[:, n_base - 3:])
    return np.concatenate([np.asarray(x, dtype=np.float64) for x in parts], axis=1).astype(np.float32)


class GoalVecOracle:
    def __init__(self, body, seed: int = 0, time_limit: int | None = 1000,
                 terminate_on_goal: bool = True, observe: dict | None = None):
        self.body = body
        self.observe = observe
        self.n = body.n
        self.seed0 = seed
        self.time_limit = time_limit
        self.terminate_on_goal = terminate_on_goal
        self.goal = np.zeros((self.n, 2), dtype=np.float32)
        self.prev_pos = np.zeros((self.n, 2))
        self.elapsed = np.zeros(self.n, dtype=np.int64)
        self.ep_ret = np.zeros(self.n)
        self.engine_seed = np.zeros(self.n, dtype=np.int64)
        self.init_rng = [ref_rng.init_box() for _ in range(self.n)]
        self.goal_rng = [ref_rng.goal_box() for _ in range(self.n)]
        self.first_reset = True
        # counters for tests: (# resets, # full resets) per env
        self.n_resets = np.zeros(self.n, dtype=np.int64)
        self.n_full = np.zeros(self.n, dtype=np.int64)

    def reset_init_space(self, low, high):
        """EnvWrapper.reset_init_space (wrapper.py:209-213) for every env; the streams continue (the batched env's
        convention: the reference swaps in the new Box object, re-seeded by EnvWrapper.seed at the next seeded reset)."""
        for b in self.init_rng:
            b.low, b.high = np.asarray(low, np.float32), np.asarray(high, np.float32)

    def reset_goal_space(self, low, high):
        """EnvWrapper.reset_goal_space (wrapper.py:215-219)."""
        for b in self.goal_rng:
            b.low, b.high = np.asarray(low, np.float32), np.asarray(high, np.float32)

    # ------------------------------------------------------------------------
    def reached(self):
        p = self.body.pos()
        g = self.goal.astype(np.float64)
        return dist2(p[:, 0], p[:, 1], g[:, 0], g[:, 1]) < REACH_RADIUS

    def _reset_env(self, i, seed=None, init_pos=None, reached_i=False):
        """EnvWrapper.reset for env i (wrapper.py:173-201)."""
        if seed is not None:  # EnvWrapper.seed, wrapper.py:95-107
            self.engine_seed[i] = seed
            self.init_rng[i].seed(seed)
            self.goal_rng[i].seed(seed + 1)
        if self.first_reset or not reached_i:
            self.engine_seed[i] += 1  # Engine.reset(), wrapper.py:190
            xy = self.init_rng[i].sample()  # wrapper.py:191
            self._set_pos_full(i, xy)
            self.n_full[i] += 1
        if init_pos is not None:
            raise NotImplementedError("init_pos is not used on the training path")
        self.goal[i] = self.goal_rng[i].sample()  # wrapper.py:196
        self.elapsed[i] = 0  # TimeLimit.reset
        self.ep_ret[i] = 0.0  # Monitor.reset
        self.n_resets[i] += 1

    def _set_pos_full(self, i, xy):
        b = self.body
        if b.engine_resets_per_full_reset == 2:  # PointEnv.set_pos resets Engine again
            self.engine_seed[i] += 1
        heading = ref_rng.engine_heading(int(self.engine_seed[i]))
        b.full_reset(i, xy, heading)

    def reset(self):
        """VecEnv.reset(): first reset of env i is seeded with seed0 + i."""
        for i in range(self.n):
            self._reset_env(i, seed=self.seed0 + i)
        self.first_reset = False
        self.prev_pos = self.body.pos()
        return self._obs()

    def _obs(self):
        return extend_obs(self.body, self.body.obs(self.goal), self.goal, self.observe)

    # ------------------------------------------------------------------------
    def step(self, actions):
        """DummyVecEnv.step_wait over all envs.

        Returns obs (N,O) f32, rewards (N,) f32, dones (N,) bool, infos dict of
        arrays: terminal_obs (N,O) f32 (valid where done), truncated (N,) bool
        (= TimeLimit.truncated), ep_r (N,) f64 / ep_l (N,) int (valid where done).
        """
        b = self.body
        b.step(actions)
        cur = b.pos()
        g = self.goal.astype(np.float64)
        d_prev = dist2(g[:, 0], g[:, 1], self.prev_pos[:, 0], self.prev_pos[:, 1])
        d_cur = dist2(g[:, 0], g[:, 1], cur[:, 0], cur[:, 1])
        reach = d_cur < REACH_RADIUS
        reward = d_prev - d_cur
        reward = np.where(reach, reward + REACH_BONUS, reward)
        terminated = reach & self.terminate_on_goal
        self.elapsed += 1
        if self.time_limit is not None:
            truncated = self.elapsed >= self.time_limit
        else:
            truncated = np.zeros(self.n, dtype=bool)
        done = terminated | truncated
        self.ep_ret += reward
        obs = self._obs()
        infos = {
            "terminal_obs": obs.copy(),
            "truncated": truncated & ~terminated,
            "terminated": terminated.copy(),
            "ep_r": self.ep_ret.copy(),
            "ep_l": self.elapsed.copy(),
        }
        self.prev_pos = cur
        if done.any():
            for i in np.nonzero(done)[0]:
                self._reset_env(int(i), reached_i=bool(reach[i]))
            self.prev_pos = b.pos()
            new_obs = self._obs()
            obs = np.where(done[:, None], new_obs, obs)
        return obs, reward.astype(np.float32), done, infos
