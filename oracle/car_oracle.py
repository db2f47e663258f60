"""fp64 numpy restatement of the car robot (``xmls/car.xml``) -- test infrastructure only.

Follows
  * model constants / tree      src/mobrob/envs/mujoco_robots/xmls/car.xml:1-57
  * Engine.step / obs           src/mobrob/envs/mujoco_robots/robots/engine.py:1392-1464, 1174-1263
                                (ballquat -> 3x3 via quat2mat engine.py:61-66, 1214-1216)
  * CarEnv.set_pos              src/mobrob/envs/wrapper.py:320-326 (free-joint qpos[0:2] only)
  * MuJoCo 2.1.0 pipeline       SURVEY.md appendix B.1, B.3, B.4 (source not on this machine)

Smooth dynamics are EXACT rigid-body mechanics for this tree: free joint (linear velocity in the
world frame, angular velocity in the body frame), two wheel hinges about body x, a ball joint.
Both wheels are axisymmetric about their hinge axis and the caster is a sphere centred on its
joint, so the system is a gyrostat: the locked inertia is constant in the chassis frame and the
equations reduce to a constant 6x6 chassis solve plus scalar / isotropic rotor equations (derived
in DESIGN.md "Car").  Joint damping (0.001 on hinges and ball) is integrated implicitly like
MuJoCo's Euler: (M + h D)^-1.

Contacts are an APPROXIMATION of MuJoCo's soft-constraint model (parity unpinned):
five points (two rim points per wheel, one under the caster), per point one normal row and two
tangential rows with MuJoCo's reference acceleration (solref 0.02/1, solimp 0.9/0.95/0.001) and
regulariser R = (1 - d)/d * A_ii, solved by projected Gauss-Seidel (box friction |f_t| <= mu f_n,
matrix free: N_SWEEPS sweeps from zero in the first substep of an env step and in obs(), N_SWEEPS_WARM
sweeps from the previous substep's forces in the other nine -- MuJoCo warm-starts its solver from the
previous qacc in the same spirit) instead of MuJoCo's pyramidal-cone Newton solver; torsional and
rolling friction are dropped.
"""
from __future__ import annotations

import math

import numpy as np

TIMESTEP = 0.004  # car.xml:3
FRAME_SKIP = 10
DENSITY = 5.0  # car.xml:5
D_ROT = 0.001  # car.xml:6 (hinges and ball; the free joint has damping 0, car.xml:15)
FORCE_LIM = 0.02  # car.xml:7
GRAV = 9.81
MAG_Y = -0.5
GOAL_Z = 0.3 / 2 + 1e-2  # engine.py:794
R_WHEEL = 0.05
HALF_LEN = 0.025
R_CASTER = 0.05
N_SWEEPS = 10  # cold solve: first substep of an env step, and every obs()
N_SWEEPS_WARM = 4  # substeps 2..10 of an env step start from the previous substep's forces
MU = 1.0  # geom friction[0] default
SOLREF_TC, SOLREF_DR = 0.02, 1.0
IMP_D0, IMP_DMAX, IMP_WIDTH = 0.9, 0.95, 0.001
OBS_DIM = 26


def _box(hx, hy, hz, pos):
    m = 8 * hx * hy * hz * DENSITY
    I = np.diag([m / 3 * (hy * hy + hz * hz), m / 3 * (hx * hx + hz * hz), m / 3 * (hx * hx + hy * hy)])
    return m, np.array(pos, float), I


def _shift(m, I, d):
    """inertia about a point displaced by d from the COM (parallel axis)."""
    return I + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))


# chassis geoms (car.xml:16-20), wheels (car.xml:21-28), caster (car.xml:29-32)
_CHASSIS = [_box(.1, .1, .05, (0, 0, 0)), _box(.1, .01, .05, (0, .15, 0)), _box(.01, .025, .03, (0, .125, 0)),
            _box(.05, .01, .05, (0, -.165, 0)), _box(.05, .03, .01, (0, -.13, .04))]
M_WHEEL = math.pi * R_WHEEL**2 * (2 * HALF_LEN) * DENSITY
I_AX = 0.5 * M_WHEEL * R_WHEEL**2
I_TR = M_WHEEL * (3 * R_WHEEL**2 + (2 * HALF_LEN) ** 2) / 12
M_CASTER = 4 / 3 * math.pi * R_CASTER**3 * DENSITY
I_S = 0.4 * M_CASTER * R_CASTER**2
POS_WL = np.array([-.1 - .03, .1, -.05])  # left body pos + cylinder centre (fromto midpoint)
POS_WR = np.array([.1 + .03, .1, -.05])
POS_C = np.array([0., -.1, -.05])
XHAT = np.array([1., 0, 0])

MASS = sum(g[0] for g in _CHASSIS) + 2 * M_WHEEL + M_CASTER
COM = (sum(g[0] * g[1] for g in _CHASSIS) + M_WHEEL * (POS_WL + POS_WR) + M_CASTER * POS_C) / MASS
# locked inertia about the body origin O, chassis frame
J_O = sum(_shift(g[0], g[2], g[1]) for g in _CHASSIS)
_I_wheel = np.diag([I_AX, I_TR, I_TR])
J_O = J_O + _shift(M_WHEEL, _I_wheel, POS_WL) + _shift(M_WHEEL, _I_wheel, POS_WR) \
    + _shift(M_CASTER, I_S * np.eye(3), POS_C)


def _cross_mat(c):
    return np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])


def _chassis_inverse(h):
    """inverse of J_c = J' + m [c]x [c]x for the (M + h D) system (h = 0: plain M)."""
    ka = I_AX / (I_AX + h * D_ROT)
    ks = I_S / (I_S + h * D_ROT)
    Jp = J_O - 2 * ka * I_AX * np.outer(XHAT, XHAT) - ks * I_S * np.eye(3)
    C = _cross_mat(COM)
    return np.linalg.inv(Jp + MASS * C @ C), ka, ks


JC_INV = {0.0: _chassis_inverse(0.0), TIMESTEP: _chassis_inverse(TIMESTEP)}


# ---- quaternion helpers (w, x, y, z), batched ---------------------------------------------
def quat_mul(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def quat_to_mat(q):
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                     np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                     np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], -2)


def quat_integrate(q, w, h):
    """mju_quatIntegrate: q <- q * exp(h w / 2) with w in the local frame, then normalise."""
    ang = np.linalg.norm(w, axis=-1) * h
    axis = w / np.maximum(np.linalg.norm(w, axis=-1, keepdims=True), 1e-300)
    dq = np.concatenate([np.cos(ang / 2)[..., None], np.sin(ang / 2)[..., None] * axis], -1)
    out = quat_mul(q, dq)
    return out / np.linalg.norm(out, axis=-1, keepdims=True)


def cross(a, b):
    return np.cross(a, b)


def mv(M, v):
    return np.einsum("...ij,...j->...i", M, v)


def mtv(M, v):
    return np.einsum("...ji,...j->...i", M, v)


class CarBody:
    obs_dim = OBS_DIM
    act_dim = 2
    name = "car"
    engine_resets_per_full_reset = 1  # CarEnv.set_pos does not rebuild the Engine

    def __init__(self, n):
        self.n = n
        self.p = np.zeros((n, 3)); self.p[:, 2] = 0.1
        self.quat = np.zeros((n, 4)); self.quat[:, 0] = 1
        self.v = np.zeros((n, 3))      # world frame
        self.w = np.zeros((n, 3))      # body frame
        self.th = np.zeros((n, 2))     # wheel angles (left, right)
        self.s = np.zeros((n, 2))      # wheel rates
        self.qb = np.zeros((n, 4)); self.qb[:, 0] = 1
        self.wb = np.zeros((n, 3))     # ball angular velocity, caster frame
        self.ctrl = np.zeros((n, 2))
        self.contacts_enabled = True
        self.last_forces = np.zeros((n, 5, 3))

    def full_reset(self, i, xy, heading):
        """Engine.reset (new MjSim at qpos0 with the body quat = heading) + CarEnv.set_pos."""
        self.p[i] = [xy[0], xy[1], 0.1]
        self.quat[i] = [math.cos(heading / 2), 0, 0, math.sin(heading / 2)]
        self.v[i] = 0; self.w[i] = 0; self.th[i] = 0; self.s[i] = 0
        self.qb[i] = [1, 0, 0, 0]; self.wb[i] = 0; self.ctrl[i] = 0

    # -- structured solve: accelerations from loads ------------------------------------------------
    def _solve(self, R, Rb, F_w, tau_O, tau_L, tau_R, tau_c, gyro, h):
        """(M + h D) qacc = loads.
        F_w: total external force, world.  tau_O: external torque about O, chassis frame.
        tau_L/R: total axial torque on each wheel (motor + damping + contact).  tau_c: total torque on
        the caster about its centre, chassis frame.  gyro: include velocity-dependent bias terms.
        Returns (vdot world, wdot body, sdot (n,2), wbdot caster frame)."""
        Jinv, ka, ks = JC_INV[h]
        w, s, wb = self.w, self.s, self.wb
        ub = mv(Rb, wb)
        fB = mtv(R, F_w)
        if gyro:
            H = mv(J_O, w) + I_AX * (s[:, 0] + s[:, 1])[:, None] * XHAT + I_S * ub
            rhs1 = fB - MASS * cross(w, cross(w, COM))
            bias_c = I_S * cross(w, ub)
            rhs2 = tau_O - ka * (tau_L + tau_R)[:, None] * XHAT - ks * (tau_c - bias_c) - cross(w, H)
        else:
            rhs1 = fB
            bias_c = 0.0
            rhs2 = tau_O - ka * (tau_L + tau_R)[:, None] * XHAT - ks * tau_c
        wdot = mv(Jinv, rhs2 - cross(COM, rhs1))
        aB = rhs1 / MASS + cross(COM, wdot) * -1.0
        # m a - m c x wdot = rhs1  ->  a = rhs1/m + c x wdot ... sign: wdot x c = -(c x wdot)
        aB = rhs1 / MASS - cross(wdot, COM)
        vdot = mv(R, aB)
        sdot = np.stack([(tau_L - I_AX * wdot[:, 0]), (tau_R - I_AX * wdot[:, 0])], 1) / (I_AX + h * D_ROT)
        rb_wbdot = (tau_c - bias_c - I_S * wdot) / (I_S + h * D_ROT)
        return vdot, wdot, sdot, mtv(Rb, rb_wbdot)

    # -- contacts -----------------------------------------------------------------------------------------
    def _contact_points(self, R):
        """5 candidate points: world position, distance, body id (0 L, 1 R, 2 caster)."""
        n = self.n
        z = np.array([0., 0, 1])
        axis = R[:, :, 0]
        d = -(z[None] - axis[:, 2:3] * axis)
        d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
        pts, body = [], []
        for centre, b in ((POS_WL, 0), (POS_WR, 1)):
            for end in (-HALF_LEN, HALF_LEN):
                cap = self.p + mv(R, centre + np.array([end, 0, 0]))
                pts.append(cap + R_WHEEL * d)
                body.append(b)
        cc = self.p + mv(R, POS_C)
        pts.append(cc - R_CASTER * z[None])
        body.append(2)
        P = np.stack(pts, 1)  # (n, 5, 3) surface points
        dist = P[:, :, 2].copy()
        P[:, :, 2] -= dist / 2  # MuJoCo puts the contact midway between the surfaces
        return P, dist, body

    def _row_jacobian_apply(self, R, Rb, r_rel_O, r_rel_body, body, direction, vdot, wdot, sdot, wbdot_c):
        """acceleration (or velocity, same linear map) of the contact point along `direction`."""
        wdot_w = mv(R, wdot)
        acc = vdot + cross(wdot_w, r_rel_O)
        if body < 2:
            acc = acc + cross(mv(R, XHAT) * sdot[:, body:body + 1], r_rel_body)
        else:
            acc = acc + cross(mv(R, mv(Rb, wbdot_c)), r_rel_body)
        return np.einsum("ni,ni->n", acc, direction)

    def _unit_load(self, R, r_rel_O, r_rel_body, body, direction):
        """generalised loads of a unit force along `direction` at the contact point."""
        F = direction
        tau_O = mtv(R, cross(r_rel_O, F))
        tL = np.zeros(self.n); tR = np.zeros(self.n); tc = np.zeros((self.n, 3))
        tb = mtv(R, cross(r_rel_body, F))  # torque about the rotor's centre, chassis frame
        if body == 0:
            tL = tb[:, 0]
        elif body == 1:
            tR = tb[:, 0]
        else:
            tc = tb
        return F, tau_O, tL, tR, tc

    def _smooth_loads(self, R, Rb):
        tau_m = np.clip(np.clip(self.ctrl, -1, 1), -FORCE_LIM, FORCE_LIM)
        tau_L = tau_m[:, 0] - D_ROT * self.s[:, 0]
        tau_R = tau_m[:, 1] - D_ROT * self.s[:, 1]
        tau_c = -D_ROT * mv(Rb, self.wb)
        F = np.zeros((self.n, 3)); F[:, 2] = -MASS * GRAV
        tau_O = mtv(R, cross(mv(R, COM), F))
        return F, tau_O, tau_L, tau_R, tau_c

    def _solve_contacts(self, R, Rb, loads, f0=None):
        """Projected Gauss-Seidel on plain M (like mj_fwdConstraint); returns per-point forces (n,5,3)
        in the frame (t1 = world x, t2 = world y, n = world z).  f0: forces to start from (warm start,
        N_SWEEPS_WARM sweeps; contacts that are not active now start from zero), None: cold."""
        n = self.n
        f = np.zeros((n, 5, 3))
        if not self.contacts_enabled:
            return f
        P, dist, bodies = self._contact_points(R)
        active = dist < 0
        if not active.any():
            return f
        dirs = np.eye(3)  # rows: tangent x, tangent y, normal z (world)
        centres = [self.p + mv(R, POS_WL), self.p + mv(R, POS_WR), self.p + mv(R, POS_C)]
        a_free = self._solve(R, Rb, *loads, gyro=True, h=0.0)
        vel = (self.v, self.w, self.s, self.wb)
        dmax = IMP_DMAX
        b_coef = 2.0 / (dmax * SOLREF_TC)
        k_coef = 1.0 / (dmax * dmax * SOLREF_TC * SOLREF_TC * SOLREF_DR * SOLREF_DR)
        rows = []
        for c in range(5):
            rO = P[:, c] - self.p
            rB = P[:, c] - centres[bodies[c]]
            x = np.minimum(np.abs(dist[:, c]) / IMP_WIDTH, 1.0)
            imp = IMP_D0 + (IMP_DMAX - IMP_D0) * np.where(x < 0.5, 2 * x * x, 1 - 2 * (1 - x) ** 2)
            for k in (2, 0, 1):  # normal first, then the two tangents
                dvec = np.broadcast_to(dirs[k], (n, 3))
                col = self._solve(R, Rb, *self._unit_load(R, rO, rB, bodies[c], dvec), gyro=False, h=0.0)
                Aii = self._row_jacobian_apply(R, Rb, rO, rB, bodies[c], dvec, *col)
                vrow = self._row_jacobian_apply(R, Rb, rO, rB, bodies[c], dvec, *vel)
                aref = -b_coef * vrow - (k_coef * imp * dist[:, c] if k == 2 else 0.0)
                Rreg = (1 - imp) / imp * Aii
                afree_row = self._row_jacobian_apply(R, Rb, rO, rB, bodies[c], dvec, *a_free)
                rows.append((c, k, rO, rB, dvec, col, Aii, Rreg, afree_row - aref))
        a_c = [np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 2)), np.zeros((n, 3))]
        if f0 is not None:
            f = np.where(active[:, :, None], f0, 0.0)
            for (c, k, rO, rB, dvec, col, Aii, Rreg, resid0) in rows:
                for q in range(4):
                    a_c[q] = a_c[q] + f[:, c, k][:, None] * col[q]
        for _ in range(N_SWEEPS if f0 is None else N_SWEEPS_WARM):
            for (c, k, rO, rB, dvec, col, Aii, Rreg, resid0) in rows:
                cur = f[:, c, k]
                res = resid0 + self._row_jacobian_apply(R, Rb, rO, rB, bodies[c], dvec, *a_c) + Rreg * cur
                new = cur - res / (Aii + Rreg)
                if k == 2:
                    new = np.maximum(new, 0.0)
                else:
                    lim = MU * f[:, c, 2]
                    new = np.clip(new, -lim, lim)
                new = np.where(active[:, c], new, 0.0)
                delta = new - cur
                for q in range(4):
                    a_c[q] = a_c[q] + delta[:, None] * col[q]
                f[:, c, k] = new
        return f

    def _loads_with_contacts(self, R, Rb, loads, f):
        F, tau_O, tau_L, tau_R, tau_c = [x.copy() for x in loads]
        if not self.contacts_enabled or not np.any(f):
            return F, tau_O, tau_L, tau_R, tau_c
        P, dist, bodies = self._contact_points(R)
        centres = [self.p + mv(R, POS_WL), self.p + mv(R, POS_WR), self.p + mv(R, POS_C)]
        for c in range(5):
            fw = f[:, c]  # world components (x, y, z)
            rO = P[:, c] - self.p
            rB = P[:, c] - centres[bodies[c]]
            F += fw
            tau_O += mtv(R, cross(rO, fw))
            tb = mtv(R, cross(rB, fw))
            if bodies[c] == 0:
                tau_L += tb[:, 0]
            elif bodies[c] == 1:
                tau_R += tb[:, 0]
            else:
                tau_c += tb
        return F, tau_O, tau_L, tau_R, tau_c

    # -- stepping -------------------------------------------------------------------------------------------
    def substep(self, warm=False):
        """One mj_step.  warm: not the first substep of the env step (start the contact solve from last_forces)."""
        h = TIMESTEP
        R, Rb = quat_to_mat(self.quat), quat_to_mat(self.qb)
        loads = self._smooth_loads(R, Rb)
        f = self._solve_contacts(R, Rb, loads, self.last_forces if warm else None)
        self.last_forces = f
        vdot, wdot, sdot, wbdot = self._solve(R, Rb, *self._loads_with_contacts(R, Rb, loads, f), gyro=True, h=h)
        self.v = self.v + h * vdot
        self.w = self.w + h * wdot
        self.s = self.s + h * sdot
        self.wb = self.wb + h * wbdot
        self.p = self.p + h * self.v
        self.quat = quat_integrate(self.quat, self.w, h)
        self.th = self.th + h * self.s
        self.qb = quat_integrate(self.qb, self.wb, h)

    def step(self, action):
        self.ctrl[:] = np.clip(np.asarray(action, dtype=np.float64), -1.0, 1.0)
        for k in range(FRAME_SKIP):
            self.substep(warm=k > 0)

    def pos(self):
        return self.p[:, :2].copy()

    def obs(self, goal):
        R, Rb = quat_to_mat(self.quat), quat_to_mat(self.qb)
        loads = self._smooth_loads(R, Rb)
        f = self._solve_contacts(R, Rb, loads)
        vdot, wdot, sdot, wbdot = self._solve(R, Rb, *self._loads_with_contacts(R, Rb, loads, f), gyro=True, h=0.0)
        o = np.zeros((self.n, OBS_DIM))
        acc = vdot.copy(); acc[:, 2] += GRAV
        o[:, 0:3] = mtv(R, acc)
        o[:, 3:6] = self.wb
        o[:, 6:15] = Rb.reshape(self.n, 9)
        g3 = np.concatenate([np.asarray(goal, np.float64), np.full((self.n, 1), GOAL_Z)], 1)
        vec = mtv(R, g3 - self.p)[:, :2]
        o[:, 15:17] = vec / (np.sqrt(np.sum(np.square(vec), 1, keepdims=True)) + 0.001)
        o[:, 17:20] = self.w
        o[:, 20:23] = mtv(R, np.broadcast_to(np.array([0, MAG_Y, 0.]), (self.n, 3)))
        o[:, 23:26] = mtv(R, self.v)
        return o.astype(np.float32)

    # -- optional Engine.obs() keys (engine.py:1243-1248), MuJoCo order: free joint, wheel hinges, rear ball ----------
    obs_pre = 15  # floats of the default row in front of goal_compass (accelerometer, ballangvel_rear, ballquat_rear)

    def qpos(self):
        return np.concatenate([self.p, self.quat, self.th, self.qb], 1)

    def qvel(self):
        return np.concatenate([self.v, self.w, self.s, self.wb], 1)

    def state_vector(self):
        """qpos(13) qvel(11) ctrl(2), MuJoCo order."""
        return np.concatenate([self.p, self.quat, self.th, self.qb, self.v, self.w, self.s, self.wb, self.ctrl], 1)

    # -- diagnostics for the conservation tests ------------------------------------------------------------------
    def momenta(self):
        """(linear momentum, angular momentum about the system COM, kinetic energy), world frame."""
        R, Rb = quat_to_mat(self.quat), quat_to_mat(self.qb)
        w, ub = self.w, mv(Rb, self.wb)
        vcom = self.v + mv(R, cross(w, COM))
        P = MASS * vcom
        H_O = mv(J_O, w) + I_AX * (self.s[:, 0] + self.s[:, 1])[:, None] * XHAT + I_S * ub
        # angular momentum about O of the moving-origin term, then shift to the COM
        vB = mtv(R, self.v)
        L_O = H_O + MASS * cross(COM, vB)
        L_com = L_O - MASS * cross(COM, mtv(R, vcom))
        T = 0.5 * MASS * np.sum(self.v * self.v, 1) + MASS * np.sum(vB * cross(w, COM), 1) \
            + 0.5 * np.sum(w * mv(J_O, w), 1) + I_AX * w[:, 0] * (self.s[:, 0] + self.s[:, 1]) \
            + 0.5 * I_AX * np.sum(self.s * self.s, 1) + I_S * np.sum(w * ub, 1) + 0.5 * I_S * np.sum(ub * ub, 1)
        return P, mv(R, L_com), T
