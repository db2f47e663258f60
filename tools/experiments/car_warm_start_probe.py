"""Probe (CPU, oracle only): would warm-starting the car's projected Gauss-Seidel from the previous substep's forces
let fewer sweeps reach the accuracy of the 10 cold sweeps the oracle / kernel use?  One-step error against a
200-sweep solve, from states on a driven trajectory.

  python tools/experiments/car_warm_start_probe.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import car_oracle as co


class Probe(co.CarBody):
    sweeps_first = 10
    sweeps_next = 10
    warm = False
    _prev = None

    def _solve_contacts(self, R, Rb, loads):
        n = self.n
        f = np.zeros((n, 5, 3))
        P, dist, bodies = self._contact_points(R)
        active = dist < 0
        if not active.any():
            self._prev = None
            return f
        dirs = np.eye(3)
        centres = [self.p + co.mv(R, co.POS_WL), self.p + co.mv(R, co.POS_WR), self.p + co.mv(R, co.POS_C)]
        a_free = self._solve(R, Rb, *loads, gyro=True, h=0.0)
        vel = (self.v, self.w, self.s, self.wb)
        b_coef = 2.0 / (co.IMP_DMAX * co.SOLREF_TC)
        k_coef = 1.0 / (co.IMP_DMAX ** 2 * co.SOLREF_TC ** 2 * co.SOLREF_DR ** 2)
        rows = []
        for c in range(5):
            rO = P[:, c] - self.p
            rB = P[:, c] - centres[bodies[c]]
            x = np.minimum(np.abs(dist[:, c]) / co.IMP_WIDTH, 1.0)
            imp = co.IMP_D0 + (co.IMP_DMAX - co.IMP_D0) * np.where(x < 0.5, 2 * x * x, 1 - 2 * (1 - x) ** 2)
            for k in (2, 0, 1):
                dvec = np.broadcast_to(dirs[k], (n, 3))
                col = self._solve(R, Rb, *self._unit_load(R, rO, rB, bodies[c], dvec), gyro=False, h=0.0)
                Aii = self._row_jacobian_apply(R, Rb, rO, rB, bodies[c], dvec, *col)
                vrow = self._row_jacobian_apply(R, Rb, rO, rB, bodies[c], dvec, *vel)
                aref = -b_coef * vrow - (k_coef * imp * dist[:, c] if k == 2 else 0.0)
                Rreg = (1 - imp) / imp * Aii
                afree_row = self._row_jacobian_apply(R, Rb, rO, rB, bodies[c], dvec, *a_free)
                rows.append((c, k, rO, rB, dvec, col, Aii, Rreg, afree_row - aref))
        a_c = [np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 2)), np.zeros((n, 3))]
        sweeps = self.sweeps_first
        if self.warm and self._prev is not None:
            f = np.where(active[:, :, None], self._prev, 0.0)
            for (c, k, rO, rB, dvec, col, *_rest) in rows:
                for q in range(4):
                    a_c[q] = a_c[q] + f[:, c, k][:, None] * col[q]
            sweeps = self.sweeps_next
        for _ in range(sweeps):
            for (c, k, rO, rB, dvec, col, Aii, Rreg, resid0) in rows:
                cur = f[:, c, k]
                res = resid0 + self._row_jacobian_apply(R, Rb, rO, rB, bodies[c], dvec, *a_c) + Rreg * cur
                new = cur - res / (Aii + Rreg)
                new = np.maximum(new, 0.0) if k == 2 else np.clip(new, -co.MU * f[:, c, 2], co.MU * f[:, c, 2])
                new = np.where(active[:, c], new, 0.0)
                delta = new - cur
                for q in range(4):
                    a_c[q] = a_c[q] + delta[:, None] * col[q]
                f[:, c, k] = new
        self._prev = f.copy()
        return f

    def step(self, action):
        self._prev = None   # cold at the first substep of every env step: no hidden state across steps
        super().step(action)


def copy_state(dst, src):
    for k in ("p", "quat", "v", "w", "th", "s", "qb", "wb", "ctrl"):
        getattr(dst, k)[:] = getattr(src, k)


def main():
    n, steps = 32, 60
    rng = np.random.default_rng(0)
    ref = Probe(n); ref.sweeps_first = ref.sweeps_next = 200
    for i in range(n):
        ref.full_reset(i, rng.uniform(-1, 1, 2), rng.uniform(0, 2 * np.pi))
    variants = {"cold 10 (current)": dict(warm=False, sweeps_first=10, sweeps_next=10),
                "cold 5": dict(warm=False, sweeps_first=5, sweeps_next=5),
                "warm 10 + 2": dict(warm=True, sweeps_first=10, sweeps_next=2),
                "warm 10 + 3": dict(warm=True, sweeps_first=10, sweeps_next=3),
                "warm 10 + 4": dict(warm=True, sweeps_first=10, sweeps_next=4),
                "warm 10 + 5": dict(warm=True, sweeps_first=10, sweeps_next=5)}
    err = {k: [] for k in variants}
    for t in range(steps):
        a = np.sign(rng.standard_normal((n, 2)))
        bodies = {}
        for name, kw in variants.items():
            b = Probe(n)
            for k, v in kw.items():
                setattr(b, k, v)
            copy_state(b, ref)
            b.step(a)
            bodies[name] = b
        ref.step(a)
        for name, b in bodies.items():
            e = max(np.abs(b.v - ref.v).max(), np.abs(b.w - ref.w).max() * 0.1, np.abs(b.s - ref.s).max() * 0.05)
            err[name].append(e)
    for name in variants:
        e = np.array(err[name])
        print(f"{name:20s} one-step velocity error vs 200 sweeps: median {np.median(e):.2e}  max {e.max():.2e}")


main()
