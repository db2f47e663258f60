"""fp64 numpy restatement of the point robot: MuJoCo 2.1.0 ``mj_step`` on
``xmls/point.xml`` plus the Engine observation (test infrastructure only).

Follows
  * model constants      src/mobrob/envs/mujoco_robots/xmls/point.xml:1-40
  * Engine.step          src/mobrob/envs/mujoco_robots/robots/engine.py:1392-1464
                         (clip ctrl, 10 x sim.step(), sim.forward())
  * Engine.obs           engine.py:1174-1263 (sorted-key flat float32 vector)
  * Engine.obs_compass   engine.py:1059-1082
  * World start pose     world.py:52-54, 119-120 (body quat = rot about z by robot_rot)
  * PointEnv.set_pos     src/mobrob/envs/wrapper.py:301-305 (position lives in body_pos)
and the published MuJoCo pipeline: ``mj_forward`` at
(q, qdot), then Euler with joint damping treated implicitly,
``qdot+ = qdot + h (M + h D)^-1 (Q_act - D qdot - C)``, ``q+ = q + h qdot+``.

Generalised coordinates: slide x, slide y (axes fixed in the frame rotated by the
start heading psi0, because the slides precede the hinge in the joint list) and
hinge z.  No contact force ever acts (sphere bottom sits at exactly z = 0 =
margin, the contact is listed but excluded) and gravity has no DoF to act on.

Pinned by KAT-1 (tests/test_oracle_kat.py): the accelerometer rows of the
shipped ``point-ppo.zip:_last_obs`` are reproduced to <= 1e-6 relative.
"""
from __future__ import annotations

import math

import numpy as np

# ---- point.xml ----------------------------------------------------------------
TIMESTEP = 0.002  # point.xml:3
FRAME_SKIP = 10  # engine.py:293-295 (binomial(10, 1.0))
DENSITY = 1.0  # point.xml:5
R_SPHERE = 0.1  # point.xml:19
HALF_BOX = 0.05  # point.xml:20
BOX_X = 0.1  # point.xml:20
D_SLIDE = 0.01  # point.xml:16-17
D_HINGE = 0.005  # point.xml:18
GEAR = 0.3  # point.xml:37-38
FORCE_LIM = 0.05  # point.xml:7-8
KV = 1.0  # <velocity> default kv
Z_HEIGHT = 0.1  # point.xml:13
MAG_Y = -0.5  # MuJoCo default magnetic = (0, -0.5, 0)
GRAVITY = 9.81
GOAL_Z = 0.3 / 2 + 1e-2  # engine.py:794

M_SPHERE = 4.0 / 3.0 * math.pi * R_SPHERE**3 * DENSITY
M_BOX = (2 * HALF_BOX) ** 3 * DENSITY
MASS = M_SPHERE + M_BOX
COM = M_BOX * BOX_X / MASS  # COM offset along body x
I_SPHERE = 0.4 * M_SPHERE * R_SPHERE**2
I_BOX = M_BOX / 12.0 * 2 * (2 * HALF_BOX) ** 2
I_O = I_SPHERE + I_BOX + M_BOX * BOX_X**2  # about the hinge axis
MC = MASS * COM

OBS_DIM = 14
ACT_DIM = 2


def actuator_forces(ctrl_x, ctrl_z, omega):
    """fwdActuation for the two actuators (ctrl already clipped to ctrlrange)."""
    f = GEAR * np.clip(ctrl_x, -FORCE_LIM, FORCE_LIM)
    tau = GEAR * np.clip(KV * ctrl_z - KV * GEAR * omega, -FORCE_LIM, FORCE_LIM)
    return f, tau


def solve_accel(theta, vx, vy, omega, f, tau, h):
    """(M + h D)^-1 (Q_act - D qdot - C) via the Schur complement on theta."""
    s, c = np.sin(theta), np.cos(theta)
    w2 = omega * omega
    r0 = f * c - D_SLIDE * vx + MC * c * w2
    r1 = f * s - D_SLIDE * vy + MC * s * w2
    r2 = tau - D_HINGE * omega
    a = MASS + h * D_SLIDE
    dth = I_O + h * D_HINGE
    schur = dth - MC * MC / a
    t = MC * (-s * r0 + c * r1) / a
    alpha = (r2 - t) / schur
    ax = (r0 + MC * s * alpha) / a
    ay = (r1 - MC * c * alpha) / a
    return ax, ay, alpha


class PointBody:
    """State of N point robots (struct of arrays, float64)."""

    obs_dim = OBS_DIM
    act_dim = ACT_DIM
    name = "point"
    engine_resets_per_full_reset = 2  # wrapper.py:190 and wrapper.py:302

    def __init__(self, n: int):
        self.n = n
        self.q = np.zeros((n, 3))  # slide x, slide y, hinge z
        self.v = np.zeros((n, 3))
        self.body_xy = np.zeros((n, 2))  # model.body_pos[robot][:2]
        self.psi0 = np.zeros(n)  # start heading (body quat)
        self.ctrl = np.zeros((n, 2))  # data.ctrl, survives goal-only resets

    # -- resets ----------------------------------------------------------------
    def full_reset(self, i, xy, heading):
        """Engine.reset (new MjSim: qpos = qvel = ctrl = 0) + PointEnv.set_pos."""
        self.q[i] = 0.0
        self.v[i] = 0.0
        self.ctrl[i] = 0.0
        self.body_xy[i] = np.asarray(xy, dtype=np.float64)
        self.psi0[i] = heading

    # -- physics ---------------------------------------------------------------
    def step(self, action):
        """Engine.step: clip to ctrlrange, FRAME_SKIP x mj_step."""
        self.ctrl[:] = np.clip(np.asarray(action, dtype=np.float64), -1.0, 1.0)
        h = TIMESTEP
        for _ in range(FRAME_SKIP):
            f, tau = actuator_forces(self.ctrl[:, 0], self.ctrl[:, 1], self.v[:, 2])
            ax, ay, al = solve_accel(
                self.q[:, 2], self.v[:, 0], self.v[:, 1], self.v[:, 2], f, tau, h
            )
            self.v[:, 0] += h * ax
            self.v[:, 1] += h * ay
            self.v[:, 2] += h * al
            self.q += h * self.v

    # -- kinematics / sensors ----------------------------------------------------
    def pos(self):
        """robot_pos[:2] (wrapper.py:269-270)."""
        c0, s0 = np.cos(self.psi0), np.sin(self.psi0)
        x = self.body_xy[:, 0] + c0 * self.q[:, 0] - s0 * self.q[:, 1]
        y = self.body_xy[:, 1] + s0 * self.q[:, 0] + c0 * self.q[:, 1]
        return np.stack([x, y], axis=1)

    def obs(self, goal):
        """Engine.obs(): mj_forward at the current state with the current ctrl."""
        theta = self.q[:, 2]
        vx, vy, om = self.v[:, 0], self.v[:, 1], self.v[:, 2]
        f, tau = actuator_forces(self.ctrl[:, 0], self.ctrl[:, 1], om)
        ax, ay, _ = solve_accel(theta, vx, vy, om, f, tau, 0.0)
        ct, st = np.cos(theta), np.sin(theta)
        psi = self.psi0 + theta
        cp, sp = np.cos(psi), np.sin(psi)
        o = np.zeros((self.n, OBS_DIM))
        # accelerometer: R(theta)^T (ax, ay) + g ez   (site frame = body frame)
        o[:, 0] = ct * ax + st * ay
        o[:, 1] = -st * ax + ct * ay
        o[:, 2] = GRAVITY
        # goal_compass: ((goal - pos) @ R)[:2] / (norm + 0.001)
        p = self.pos()
        dx = np.asarray(goal, dtype=np.float64)[:, 0] - p[:, 0]
        dy = np.asarray(goal, dtype=np.float64)[:, 1] - p[:, 1]
        ex = cp * dx + sp * dy
        ey = -sp * dx + cp * dy
        nrm = np.sqrt(ex * ex + ey * ey) + 0.001
        o[:, 3] = ex / nrm
        o[:, 4] = ey / nrm
        # gyro
        o[:, 7] = om
        # magnetometer: R(psi)^T (0, -0.5, 0)
        o[:, 8] = sp * MAG_Y
        o[:, 9] = cp * MAG_Y
        # velocimeter: R(theta)^T (vx, vy)
        o[:, 11] = ct * vx + st * vy
        o[:, 12] = -st * vx + ct * vy
        return o.astype(np.float32)

    # -- optional Engine.obs() keys (engine.py:1243-1248): data.qpos / data.qvel are the joint coordinates ----------
    obs_pre = 3  # floats of the default row in front of goal_compass (accelerometer)

    def qpos(self):
        return self.q.copy()

    def qvel(self):
        return self.v.copy()

    # -- state export (tests) -----------------------------------------------------
    def state_vector(self):
        return np.concatenate(
            [self.q, self.v, self.body_xy, self.psi0[:, None], self.ctrl], axis=1
        )
