"""ORACLE (test infrastructure, never imported by the product): CPU restatement of the stand-alone FCT variant of the
reference, ``/root/reference/test/forward_fct.py`` -- ``solve_VOF_rudman`` :254-264, ``fct_x_sweep`` :267-308,
``fct_y_sweep`` :310-351, ``set_BC`` :218-229 -- in NumPy fp32, every expression in the order it is written there
(left to right, Python-scalar sub-expressions ``dx * dy``, ``dt * dy``, ``dt * dx`` folded in double and rounded once,
as Taichi folds them).  Each of the five loops of a sweep writes a zero-initialised array of its own time level in
the reference (``Ftd_x[t]``, ``ax[t]`` ...), so an entry the loop bounds never reach is 0.

Pinned: ``tests/test_reference_pin_cpu.py::test_forward_fct_oracle_equals_reference_run`` compares every stored
half-step level of ``tests/golden/ref_fct_*.npz`` -- the unmodified script text executed under the taichi stand-in
(``oracle/run_reference.py``) -- with this restatement, bit for bit.
"""
from __future__ import annotations

import numpy as np

R = np.float32


class FctForwardOracle:
    def __init__(self, nx, ny, dx, dy, dt, eps=1.0e-4):
        self.nx, self.ny = nx, ny
        self.dx, self.dy, self.dt = R(dx), R(dy), R(dt)
        self.dxdy = R(float(dx) * float(dy))
        self.dtdy = R(float(dt) * float(dy))        # forward_fct.py:269
        self.dtdx = R(float(dt) * float(dx))        # :312
        self.eps = R(eps)
        shp = (nx + 2, ny + 2)
        self.F = np.zeros(shp, R)
        self.u = np.zeros(shp, R)
        self.v = np.zeros(shp, R)
        self.t = 0

    # :218-229 -- loop over i first, then over j (the corner takes the second loop's value)
    def set_BC(self):
        F, nx, ny = self.F, self.nx, self.ny
        F[:, 0] = F[:, 1]
        F[:, ny + 1] = F[:, ny]
        F[0, :] = F[1, :]
        F[nx + 1, :] = F[nx, :]

    def _ratio(self, q, p):
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            return np.where(p > 0, np.minimum(R(1), q / (p + self.eps)), R(0)).astype(R)

    def _sweep(self, axis):
        """One sweep along ``axis`` (0: fct_x_sweep :267-308, 1: fct_y_sweep :310-351).  The y sweep is the x sweep with
        i and j exchanged, v for u and dt * dx for dt * dy; every other factor (dy, dx * dy, dx in q) is the same text."""
        F = self.F if axis == 0 else self.F.T
        w = self.u if axis == 0 else self.v.T
        n = self.nx if axis == 0 else self.ny
        m = self.ny if axis == 0 else self.nx
        dt, dx, dy, dxdy = self.dt, self.dx, self.dy, self.dxdy
        dtd = self.dtdy if axis == 0 else self.dtdx
        zero = R(0)
        c, lo, hi = slice(1, n + 1), slice(0, n), slice(2, n + 2)
        cj = slice(1, m + 1)
        shp = F.shape
        # loop 1 (:268-272): cells 1 .. n
        wc, wp = w[c, cj], w[hi, cj]
        with np.errstate(all="ignore"):
            dv = dxdy - dtd * (wp - wc)
            fl_L = np.where(wc >= 0, wc * dt * F[lo, cj], wc * dt * F[c, cj])
            fr_L = np.where(wp >= 0, wp * dt * F[c, cj], wp * dt * F[hi, cj])
            Ftd = np.zeros(shp, R)
            Ftd[c, cj] = F[c, cj] + (fl_L - fr_L) * dy / dxdy * dx * dy / dv
            # loop 2 (:274-277): faces 1 .. n + 1
            f = slice(1, n + 2)
            fm = slice(0, n + 1)
            wf = w[f, cj]
            a_L = np.where(wf >= 0, wf * dt * F[fm, cj], wf * dt * F[f, cj])
            a_H = np.where(wf <= 0, wf * dt * F[fm, cj], wf * dt * F[f, cj])
            a = np.zeros(shp, R)
            a[f, cj] = a_H - a_L
            # loop 3 (:279-297)
            fmax = np.maximum(np.maximum(Ftd[c, cj], Ftd[lo, cj]), Ftd[hi, cj])
            fmin = np.minimum(np.minimum(Ftd[c, cj], Ftd[lo, cj]), Ftd[hi, cj])
            rp = np.zeros(shp, R)
            rm = np.zeros(shp, R)
            pp = np.maximum(zero, a[c, cj]) - np.minimum(zero, a[hi, cj])
            rp[c, cj] = self._ratio((fmax - Ftd[c, cj]) * dx, pp)
            pm = np.maximum(zero, a[hi, cj]) - np.minimum(zero, a[c, cj])
            rm[c, cj] = self._ratio((Ftd[c, cj] - fmin) * dx, pm)
            # loop 4 (:299-303): face i + 1 from cell i = 1 .. n
            cl = np.zeros(shp, R)
            cl[hi, cj] = np.where(a[hi, cj] >= 0, np.minimum(rp[hi, cj], rm[c, cj]), np.minimum(rp[c, cj], rm[hi, cj]))
            # loop 5 (:305-308)
            out = Ftd[c, cj] - ((a[hi, cj] * cl[hi, cj] - a[c, cj] * cl[c, cj]) / dy) * dx * dy / dv
        Fn = np.zeros(shp, R)           # F[level + 1] starts from zeros; set_BC then fills every ghost cell
        Fn[c, cj] = out
        self.F = np.ascontiguousarray(Fn if axis == 0 else Fn.T)

    def fct_x_sweep(self):
        self._sweep(0)

    def fct_y_sweep(self):
        self._sweep(1)

    # :254-264
    def solve_VOF_rudman(self):
        order = (1, 0) if self.t % 2 == 0 else (0, 1)
        for axis in order:
            self._sweep(axis)
            self.set_BC()
        self.t += 1
