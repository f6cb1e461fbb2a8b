"""CPU oracle for the 2-D VOF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Order-exact fp32 NumPy restatement of the per-timestep path of the reference
solver ``/root/reference/2dvof.py`` (lines 513-528 and every kernel they call).

PARITY PINNED TO THE REFERENCE'S OWN SOURCE, EXECUTED: ``oracle/run_reference.py`` runs the unmodified text of
``/root/reference/2dvof.py`` (only the size constants substituted) under ``oracle/refshim/taichi`` -- a NumPy stand-in
for taichi==1.4.1 (requirements.txt:3; not installable in this image: py3.12, offline) -- and commits what it computed
as ``tests/golden/ref_2d_*.npz``; ``tests/test_reference_pin_cpu.py`` requires this oracle to equal those fixtures bit
for bit, after every kernel call of the first steps and at the step snapshots (-ic 1/2/3 at the reference's 200 x 200,
small non-square grids, injected synthetic states).  What stays outside the pin is the Taichi *compiler*: its
``fast_math`` may re-associate at round-off level.  The reference itself ships no golden vectors
or assertions for this path (``/root/reference/test/*.py`` are GUI demos).  The conventions the restatement follows:

* every top-level ``for`` of a ``@ti.kernel`` is "read the old arrays, write the
  whole result" (NumPy slice assignment gives exactly Taichi's barrier
  semantics between offloaded loops);
* all field arithmetic is IEEE fp32, left-to-right as written, no FMA
  contraction, correctly-rounded ``/`` and ``sqrt``;
* sub-expressions made only of Python scalars (``dxi**2``, ``dx*dy``,
  ``dt*dy``, ``-1/(2*dx)``, ``1/dx/2`` ...) are folded in double by the DSL
  front-end and rounded once to fp32;
* Python ints meeting a field value become fp32 constants.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.

Each method cites the reference lines it restates.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = ["Vof2DParams", "Vof2DOracle", "rel_linf"]


class Vof2DParams:
    """Constants block, 2dvof.py:19-50.  Defaults reproduce the reference."""

    def __init__(self, nx=200, ny=200, Lx=0.1, Ly=0.1, rho_l=1000.0, rho_g=50.0,
                 nu_l=1.0e-6, nu_g=1.5e-5, sigma=0.007, gx=0, gy=-5, dt=4e-6,
                 n_jacobi=10):
        self.nx, self.ny = int(nx), int(ny)
        self.Lx, self.Ly = float(Lx), float(Ly)
        self.rho_l, self.rho_g = float(rho_l), float(rho_g)
        self.nu_l, self.nu_g = float(nu_l), float(nu_g)
        self.sigma = float(sigma)
        self.gx, self.gy = gx, gy
        self.dt = float(dt)
        self.n_jacobi = int(n_jacobi)
        # 2dvof.py:41-50 -- node coordinates as an fp32 array, dx from two of them
        self.x = np.hstack((0.0, np.linspace(0, self.Lx, self.nx + 1), self.Lx)).astype(np.float32)
        self.y = np.hstack((0.0, np.linspace(0, self.Ly, self.ny + 1), self.Ly)).astype(np.float32)
        self.dx = float(self.x[3]) - float(self.x[2])   # python double (imin+2, imin+1)
        self.dy = float(self.y[3]) - float(self.y[2])
        self.dxi = 1 / self.dx
        self.dyi = 1 / self.dy

    @classmethod
    def scaled(cls, n, **kw):
        """Constant-dx scaling for large synthetic grids (risk 3)."""
        L = 0.1 * n / 200.0
        return cls(nx=n, ny=n, Lx=L, Ly=L, **kw)


def rel_linf(a, b):
    """Per-field relative L-inf used by the parity gates: max|a-b| / max|b|."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    num = np.max(np.abs(a - b))
    return float(num / den) if den > 0 else float(num)


class Vof2DOracle:
    """State + kernels.  ``real`` may be np.float64 for the sensitivity twin."""

    FIELDS = ("F", "u", "v", "p", "rho", "nu", "kappa", "u_star", "v_star")

    def __init__(self, params: Vof2DParams | None = None, real=np.float32):
        self.P = P = params or Vof2DParams()
        self.real = R = real
        shape = (P.nx + 2, P.ny + 2)
        z = lambda: np.zeros(shape, dtype=R)
        # 2dvof.py:53-89 (zero-initialised; Ap/rhs/V/rgb_buf are dead or display-only)
        self.F, self.Ftd = z(), z()
        self.ax, self.ay, self.cx, self.cy, self.rp, self.rm = z(), z(), z(), z(), z(), z()
        self.u, self.v, self.u_star, self.v_star = z(), z(), z(), z()
        self.p, self.pt, self.rho, self.nu = z(), z(), z(), z()
        self.mx, self.my, self.kappa = z(), z(), z()
        self.istep = 0
        self.courant_flags = 0
        # constants as seen by the kernels (double-folded, then one rounding)
        c = lambda val: R(val)
        self.c_dt = c(P.dt)
        self.c_dx, self.c_dy = c(P.dx), c(P.dy)
        self.c_dxi, self.c_dyi = c(P.dxi), c(P.dyi)
        self.c_dxi2, self.c_dyi2 = c(P.dxi ** 2), c(P.dyi ** 2)
        self.c_dxdy = c(P.dx * P.dy)
        self.c_dtdy, self.c_dtdx = c(P.dt * P.dy), c(P.dt * P.dx)
        self.c_m1_2dx, self.c_m1_2dy = c(-1 / (2 * P.dx)), c(-1 / (2 * P.dy))
        self.c_1_dx_2, self.c_1_dy_2 = c(1 / P.dx / 2), c(1 / P.dy / 2)
        self.c_sigma = c(P.sigma)
        self.c_rho_l, self.c_rho_g = c(P.rho_l), c(P.rho_g)
        self.c_nu_l, self.c_nu_g = c(P.nu_l), c(P.nu_g)
        self.c_gx, self.c_gy = c(P.gx), c(P.gy)
        self.c_cfl_x, self.c_cfl_y = c(0.25 * P.dx), c(0.25 * P.dy)

    # ------------------------------------------------------------------ helpers
    def _var(self, a, b, c):
        """2dvof.py:192-195 -- ``a + b + c - max(a,b,c) - min(a,b,c)``, left to right."""
        R = self.real
        a, b, c = (np.asarray(t, dtype=R) for t in (a, b, c))
        return ((a + b) + c) - np.maximum(np.maximum(a, b), c) - np.minimum(np.minimum(a, b), c)

    # ------------------------------------------------------------- initial state
    def _find_area(self, cx, cy, r):
        """2dvof.py:102-134 on the whole (nx+2, ny+2) index grid."""
        P, R = self.P, self.real
        dx, dy = self.c_dx, self.c_dy
        hdx, hdy = R(P.dx / 2), R(P.dy / 2)
        ii = (np.arange(P.nx + 2, dtype=np.int32) - 1).astype(R)[:, None]
        jj = (np.arange(P.ny + 2, dtype=np.int32) - 1).astype(R)[None, :]
        xc = ii * dx + hdx
        yc = jj * dy + hdy
        xl, xr = xc - hdx, xc + hdx
        yd, yu = yc - hdy, yc + hdy

        def dist(xx, yy):
            ddx = xx - cx
            ddy = yy - cy
            return np.sqrt(ddx * ddx + ddy * ddy)

        d_ct, d_lu, d_ld, d_ru, d_rd = dist(xc, yc), dist(xl, yu), dist(xl, yd), dist(xr, yu), dist(xr, yd)
        outside = (d_lu > r) & (d_ld > r) & (d_ru > r) & (d_rd > r)
        inside = (d_lu < r) & (d_ld < r) & (d_ru < r) & (d_rd < r)
        a = R(0.5) + R(0.5) * (d_ct - r) / R(math.sqrt(2.0) * P.dx)
        a = self._var(a, R(0), R(1))
        a = np.where(inside, R(0.0), a)
        a = np.where(outside, R(1.0), a)
        return a.astype(R)

    def set_init_F(self, ic: int):
        """2dvof.py:137-159."""
        P, R = self.P, self.real
        x = P.x[: P.nx + 2].astype(R)[:, None]
        y = P.y[: P.ny + 2].astype(R)[None, :]
        if ic == 1:
            x1, x2, y1, y2 = R(0.0), R(P.Lx / 3), R(0.0), R(P.Ly / 2)
            m = (x >= x1) & (x <= x2) & (y >= y1) & (y <= y2)
            self.F[m] = R(1.0)
        elif ic == 2:
            r = R(P.Lx / 12)
            cx, cy = R(P.Lx / 2), R(2) * r
            self.F[...] = self._find_area(cx, cy, r)
        elif ic == 3:
            r = R(P.Lx / 12)
            cx, cy = R(P.Lx / 2), R(P.Ly) - R(3) * r
            Fn = R(1.0) - self._find_area(cx, cy, r)
            Fn = np.where(np.broadcast_to(y < R(P.Ly * 0.37), Fn.shape), R(1.0), Fn)
            self.F[...] = Fn
        else:
            raise ValueError("ic must be 1, 2 or 3")

    # ------------------------------------------------------------------ kernels
    def set_BC(self):
        """2dvof.py:162-189 -- row loop, then column loop (corner order matters)."""
        P = self.P
        nx, ny = P.nx, P.ny
        u, v, F, p, rho = self.u, self.v, self.F, self.p, self.rho
        # loop A over i in [0, nx+1]
        u[:, 0] = u[:, 1]
        v[:, 1] = 0
        F[:, 0] = F[:, 1]
        p[:, 0] = p[:, 1]
        rho[:, 0] = rho[:, 1]
        u[:, ny + 1] = u[:, ny]
        v[:, ny + 1] = 0
        F[:, ny + 1] = F[:, ny]
        p[:, ny + 1] = p[:, ny]
        rho[:, ny + 1] = rho[:, ny]
        # loop B over j in [0, ny+1]
        u[1, :] = 0
        v[0, :] = v[1, :]
        F[0, :] = F[1, :]
        p[0, :] = p[1, :]
        rho[0, :] = rho[1, :]
        u[nx + 1, :] = 0
        v[nx + 1, :] = v[nx, :]
        F[nx + 1, :] = F[nx, :]
        p[nx + 1, :] = p[nx, :]
        rho[nx + 1, :] = rho[nx, :]

    def cal_nu_rho(self):
        """2dvof.py:198-203 -- all cells including ghosts."""
        R = self.real
        Fc = self._var(R(0.0), R(1.0), self.F)
        self.rho[...] = self.c_rho_g * (R(1) - Fc) + self.c_rho_l * Fc
        self.nu[...] = self.c_nu_l * Fc + self.c_nu_g * (R(1.0) - Fc)

    def get_normal_young(self):
        """2dvof.py:283-309."""
        P, R = self.P, self.real
        nx, ny = P.nx, P.ny
        F = self.F
        c, m, pl = slice(1, nx + 1), slice(0, nx), slice(2, nx + 2)
        cj, mj, pj = slice(1, ny + 1), slice(0, ny), slice(2, ny + 2)
        kx, ky = self.c_m1_2dx, self.c_m1_2dy
        mx1 = kx * (F[pl, pj] + F[pl, cj] - F[c, pj] - F[c, cj])
        my1 = ky * (F[pl, pj] - F[pl, cj] + F[c, pj] - F[c, cj])
        mx2 = kx * (F[pl, cj] + F[pl, mj] - F[c, cj] - F[c, mj])
        my2 = ky * (F[pl, cj] - F[pl, mj] + F[c, cj] - F[c, mj])
        mx3 = kx * (F[c, cj] + F[c, mj] - F[m, cj] - F[m, mj])
        my3 = ky * (F[c, cj] - F[c, mj] + F[m, cj] - F[m, mj])
        mx4 = kx * (F[c, pj] + F[c, cj] - F[m, pj] - F[m, cj])
        my4 = ky * (F[c, pj] - F[c, cj] + F[m, pj] - F[m, cj])
        mxsum = (mx1 + mx2 + mx3 + mx4) / R(4)
        mysum = (my1 + my2 + my3 + my4) / R(4)
        small = (np.abs(mxsum) < R(1e-10)) & (np.abs(mysum) < R(1e-10))
        mag = np.sqrt(mxsum * mxsum + mysum * mysum)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.mx[c, cj] = np.where(small, mxsum, mxsum / mag)
            self.my[c, cj] = np.where(small, mysum, mysum / mag)
        mx, my = self.mx, self.my
        self.kappa[c, cj] = -(self.c_1_dx_2 * (mx[pl, cj] - mx[m, cj]) +
                              self.c_1_dy_2 * (my[c, pj] - my[c, mj]))

    def advect_upwind(self):
        """2dvof.py:206-233."""
        P, R = self.P, self.real
        nx, ny = P.nx, P.ny
        u, v, F, kappa, nu, rho = self.u, self.v, self.F, self.kappa, self.nu, self.rho
        dt, dxi, dyi, dxi2, dyi2 = self.c_dt, self.c_dxi, self.c_dyi, self.c_dxi2, self.c_dyi2
        two = R(2)
        # ---- u*: i in [2, nx], j in [1, ny]
        c, m, pl = slice(2, nx + 1), slice(1, nx), slice(3, nx + 2)
        cj, mj, pj = slice(1, ny + 1), slice(0, ny), slice(2, ny + 2)
        uc = u[c, cj]
        v_here = R(0.25) * (v[m, cj] + v[m, pj] + v[c, cj] + v[c, pj])
        dudx = np.where(uc > 0, (uc - u[m, cj]) * dxi, (u[pl, cj] - uc) * dxi)
        dudy = np.where(v_here > 0, (uc - u[c, mj]) * dyi, (u[c, pj] - uc) * dyi)
        kappa_ave = (kappa[c, cj] + kappa[m, cj]) / R(2.0)
        fx_kappa = (-self.c_sigma) * (F[c, cj] - F[m, cj]) * kappa_ave / self.c_dx
        us = uc + dt * (
            nu[c, cj] * (u[m, cj] - two * uc + u[pl, cj]) * dxi2
            + nu[c, cj] * (u[c, mj] - two * uc + u[c, pj]) * dyi2
            - uc * dudx - v_here * dudy
            + self.c_gx + fx_kappa * two / (rho[c, cj] + rho[m, cj]))
        # ---- v*: i in [1, nx], j in [2, ny]
        c2, m2, p2 = slice(1, nx + 1), slice(0, nx), slice(2, nx + 2)
        cj2, mj2, pj2 = slice(2, ny + 1), slice(1, ny), slice(3, ny + 2)
        vc = v[c2, cj2]
        u_here = R(0.25) * (u[c2, mj2] + u[c2, cj2] + u[p2, mj2] + u[p2, cj2])
        dvdx = np.where(u_here > 0, (vc - v[m2, cj2]) * dxi, (v[p2, cj2] - vc) * dxi)
        dvdy = np.where(vc > 0, (vc - v[c2, mj2]) * dyi, (v[c2, pj2] - vc) * dyi)
        kappa_ave2 = (kappa[c2, cj2] + kappa[c2, mj2]) / R(2.0)
        fy_kappa = (-self.c_sigma) * (F[c2, cj2] - F[c2, mj2]) * kappa_ave2 / self.c_dy
        vs = vc + dt * (
            nu[c2, cj2] * (v[m2, cj2] - two * vc + v[p2, cj2]) * dxi2
            + nu[c2, cj2] * (v[c2, mj2] - two * vc + v[c2, pj2]) * dyi2
            - u_here * dvdx - vc * dvdy
            + self.c_gy + fy_kappa * two / (rho[c2, cj2] + rho[c2, mj2]))
        self.u_star[c, cj] = us
        self.v_star[c2, cj2] = vs

    def poisson_rhs(self):
        """rhs of 2dvof.py:239-241 (value-identical in all sweeps of one step)."""
        P = self.P
        nx, ny = P.nx, P.ny
        c, pl = slice(1, nx + 1), slice(2, nx + 2)
        cj, pj = slice(1, ny + 1), slice(2, ny + 2)
        us, vs = self.u_star, self.v_star
        return self.rho[c, cj] / self.c_dt * (
            (us[pl, cj] - us[c, cj]) * self.c_dxi + (vs[c, pj] - vs[c, cj]) * self.c_dyi)

    def solve_p_jacobi(self):
        """2dvof.py:236-266 -- ONE sweep (the loop at 521-522 calls it 10 times)."""
        P, R = self.P, self.real
        nx, ny = P.nx, P.ny
        c, m, pl = slice(1, nx + 1), slice(0, nx), slice(2, nx + 2)
        cj, mj, pj = slice(1, ny + 1), slice(0, ny), slice(2, ny + 2)
        rhs = self.poisson_rhs()
        ii = np.arange(1, nx + 1)[:, None]
        jj = np.arange(1, ny + 1)[None, :]
        zero = R(0.0)
        ae = np.where(ii != nx, self.c_dxi2, zero).astype(R)
        aw = np.where(ii != 1, self.c_dxi2, zero).astype(R)
        an = np.where(jj != ny, self.c_dyi2, zero).astype(R)
        a_s = np.where(jj != 1, self.c_dyi2, zero).astype(R)
        ap = R(-1.0) * (ae + aw + an + a_s)
        p = self.p
        self.pt[c, cj] = (rhs - ae * p[pl, cj] - aw * p[m, cj] - an * p[c, pj] - a_s * p[c, mj]) / ap
        self.p[c, cj] = self.pt[c, cj]

    def update_uv(self):
        """2dvof.py:269-280."""
        P, R = self.P, self.real
        nx, ny = P.nx, P.ny
        rho, p = self.rho, self.p
        c, m = slice(2, nx + 1), slice(1, nx)
        cj = slice(1, ny + 1)
        r = (rho[c, cj] + rho[m, cj]) * R(0.5)
        self.u[c, cj] = self.u_star[c, cj] - self.c_dt / r * (p[c, cj] - p[m, cj]) * self.c_dxi
        c2 = slice(1, nx + 1)
        cj2, mj2 = slice(2, ny + 1), slice(1, ny)
        r2 = (rho[c2, cj2] + rho[c2, mj2]) * R(0.5)
        self.v[c2, cj2] = self.v_star[c2, cj2] - self.c_dt / r2 * (p[c2, cj2] - p[c2, mj2]) * self.c_dyi
        # 274-275 / 279-280: device-side print only; counted here as a diagnostic
        self.courant_flags = int(np.count_nonzero(self.u[c, cj] * self.c_dt > self.c_cfl_x)
                                 + np.count_nonzero(self.v[c2, cj2] * self.c_dt > self.c_cfl_y))

    def _limit(self, q, pq):
        R = self.real
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(pq > 0, np.minimum(R(1), q / pq), R(0.0)).astype(R)

    def fct_x_sweep(self):
        """2dvof.py:321-382."""
        P, R = self.P, self.real
        nx, ny = P.nx, P.ny
        F, u, Ftd = self.F, self.u, self.Ftd
        dt, dx, dy, dxdy, dtdy = self.c_dt, self.c_dx, self.c_dy, self.c_dxdy, self.c_dtdy
        c, m, pl = slice(1, nx + 1), slice(0, nx), slice(2, nx + 2)
        cj, mj, pj = slice(1, ny + 1), slice(0, ny), slice(2, ny + 2)
        zero = R(0)
        # loop 1 (323-331)
        uc, up = u[c, cj], u[pl, cj]
        dv = dxdy - dtdy * (up - uc)
        fl_L = np.where(uc >= 0, uc * dt * F[m, cj], uc * dt * F[c, cj])
        fr_L = np.where(up >= 0, up * dt * F[c, cj], up * dt * F[pl, cj])
        t = (F[c, cj] + (fl_L - fr_L + zero - zero) * dy / dxdy) * dx * dy / dv
        t = np.where((t > R(1.)) | (t < 0), self._var(R(0), R(1), t), t)
        Ftd[c, cj] = t
        # loop 2 (333-363)
        fmax = np.maximum(np.maximum(Ftd[c, cj], Ftd[m, cj]), Ftd[pl, cj])
        fmin = np.minimum(np.minimum(Ftd[c, cj], Ftd[m, cj]), Ftd[pl, cj])
        fl_H = np.where(uc <= 0, uc * dt * F[m, cj], uc * dt * F[c, cj])
        fr_H = np.where(up <= 0, up * dt * F[c, cj], up * dt * F[pl, cj])
        self.ax[pl, cj] = fr_H - fr_L
        self.ax[c, cj] = fl_H - fl_L          # same-value double write on shared faces
        self.ay[c, pj] = 0
        self.ay[c, cj] = 0
        ax, ay = self.ax, self.ay
        pp = (np.maximum(zero, ax[c, cj]) - np.minimum(zero, ax[pl, cj])
              + np.maximum(zero, ay[c, cj]) - np.minimum(zero, ay[c, pj]))
        qp = (fmax - Ftd[c, cj]) * dx
        self.rp[c, cj] = self._limit(qp, pp)
        pm = (np.maximum(zero, ax[pl, cj]) - np.minimum(zero, ax[c, cj])
              + np.maximum(zero, ay[c, pj]) - np.minimum(zero, ay[c, cj]))
        qm = (Ftd[c, cj] - fmin) * dx
        self.rm[c, cj] = self._limit(qm, pm)
        # loop 3 (365-374)
        rp, rm = self.rp, self.rm
        self.cx[pl, cj] = np.where(ax[pl, cj] >= 0, np.minimum(rp[pl, cj], rm[c, cj]),
                                   np.minimum(rp[c, cj], rm[pl, cj]))
        self.cy[c, pj] = np.where(ay[c, pj] >= 0, np.minimum(rp[c, pj], rm[c, cj]),
                                  np.minimum(rp[c, cj], rm[c, pj]))
        # loop 4 (376-382)
        cx, cy = self.cx, self.cy
        dv = dxdy - dtdy * (u[pl, cj] - u[c, cj])
        Fn = Ftd[c, cj] - ((ax[pl, cj] * cx[pl, cj] - ax[c, cj] * cx[c, cj]
                            + ay[c, pj] * cy[c, pj] - ay[c, cj] * cy[c, cj]) / dy) * dx * dy / dv
        F[c, cj] = self._var(R(0), R(1), Fn)

    def fct_y_sweep(self):
        """2dvof.py:385-448."""
        P, R = self.P, self.real
        nx, ny = P.nx, P.ny
        F, v, Ftd = self.F, self.v, self.Ftd
        dt, dx, dy, dxdy, dtdx = self.c_dt, self.c_dx, self.c_dy, self.c_dxdy, self.c_dtdx
        c, m, pl = slice(1, nx + 1), slice(0, nx), slice(2, nx + 2)
        cj, mj, pj = slice(1, ny + 1), slice(0, ny), slice(2, ny + 2)
        zero = R(0)
        # loop 1 (387-395)
        vc, vp = v[c, cj], v[c, pj]
        dv = dxdy - dtdx * (vp - vc)
        ft_L = np.where(vp >= 0, vp * dt * F[c, cj], vp * dt * F[c, pj])
        fb_L = np.where(vc >= 0, vc * dt * F[c, mj], vc * dt * F[c, cj])
        t = (F[c, cj] + (zero - zero + fb_L - ft_L) * dy / dxdy) * dx * dy / dv
        t = np.where((t > R(1.)) | (t < 0), self._var(R(0), R(1), t), t)
        Ftd[c, cj] = t
        # loop 2 (397-427)
        fmax = np.maximum(np.maximum(Ftd[c, cj], Ftd[c, mj]), Ftd[c, pj])
        fmin = np.minimum(np.minimum(Ftd[c, cj], Ftd[c, mj]), Ftd[c, pj])
        ft_H = np.where(vp <= 0, vp * dt * F[c, cj], vp * dt * F[c, pj])
        fb_H = np.where(vc <= 0, vc * dt * F[c, mj], vc * dt * F[c, cj])
        self.ax[pl, cj] = 0
        self.ax[c, cj] = 0
        self.ay[c, pj] = ft_H - ft_L
        self.ay[c, cj] = fb_H - fb_L
        ax, ay = self.ax, self.ay
        pp = (np.maximum(zero, ax[c, cj]) - np.minimum(zero, ax[pl, cj])
              + np.maximum(zero, ay[c, cj]) - np.minimum(zero, ay[c, pj]))
        qp = (fmax - Ftd[c, cj]) * dx
        self.rp[c, cj] = self._limit(qp, pp)
        pm = (np.maximum(zero, ax[pl, cj]) - np.minimum(zero, ax[c, cj])
              + np.maximum(zero, ay[c, pj]) - np.minimum(zero, ay[c, cj]))
        qm = (Ftd[c, cj] - fmin) * dx
        self.rm[c, cj] = self._limit(qm, pm)
        # loop 3 (429-438)
        rp, rm = self.rp, self.rm
        self.cx[pl, cj] = np.where(ax[pl, cj] >= 0, np.minimum(rp[pl, cj], rm[c, cj]),
                                   np.minimum(rp[c, cj], rm[pl, cj]))
        self.cy[c, pj] = np.where(ay[c, pj] >= 0, np.minimum(rp[c, pj], rm[c, cj]),
                                  np.minimum(rp[c, cj], rm[c, pj]))
        # loop 4 (441-448)
        cx, cy = self.cx, self.cy
        dv = dxdy - dtdx * (v[c, pj] - v[c, cj])
        Fn = Ftd[c, cj] - ((ax[pl, cj] * cx[pl, cj] - ax[c, cj] * cx[c, cj]
                            + ay[c, pj] * cy[c, pj] - ay[c, cj] * cy[c, cj]) / dy) * dx * dy / dv
        F[c, cj] = self._var(R(0), R(1), Fn)

    def solve_VOF_rudman(self):
        """2dvof.py:312-318 -- sweep order alternates with istep parity."""
        if self.istep % 2 == 0:
            self.fct_y_sweep()
            self.fct_x_sweep()
        else:
            self.fct_x_sweep()
            self.fct_y_sweep()

    def post_process_f(self):
        """2dvof.py:452-455 -- all cells including ghosts."""
        R = self.real
        self.F[...] = self._var(self.F, R(0), R(1))

    # --------------------------------------------------------------- main loop
    # 2dvof.py:458-486: rgb_buf[I] = field[I // r], r = resolution[0] // nx = 2; velocities over L / 0.2
    def _upsample(self, a):
        P = self.P
        return np.repeat(np.repeat(a[:P.nx, :P.ny], 2, axis=0), 2, axis=1).astype(self.real)

    def get_vof_field(self):
        return self._upsample(self.F)

    def get_u_field(self):
        return self._upsample(self.u / self.real(self.P.Lx / 0.2))

    def get_v_field(self):
        return self._upsample(self.v / self.real(self.P.Ly / 0.2))

    def get_vnorm_field(self):
        R = self.real
        return self._upsample(np.sqrt(self.u * self.u + self.v * self.v).astype(R) / R(self.P.Ly / 0.2))

    # 2dvof.py:489-492.  The reference's i-range reaches nx+1 and reads u[nx+2, j] (out of bounds, undefined): that
    # row is left at zero here and in the CUDA kernel.
    def interp_velocity(self):
        P, R = self.P, self.real
        V = np.zeros((P.nx + 2, P.ny + 2, 2), dtype=R)
        V[1:P.nx + 1, 1:P.ny + 1, 0] = (self.u[1:P.nx + 1, 1:P.ny + 1] + self.u[2:P.nx + 2, 1:P.ny + 1]) / R(2)
        V[1:P.nx + 1, 1:P.ny + 1, 1] = (self.v[1:P.nx + 1, 1:P.ny + 1] + self.v[1:P.nx + 1, 2:P.ny + 2]) / R(2)
        return V

    def step(self):
        """One iteration of the loop body, 2dvof.py:506-528."""
        self.istep += 1
        self.cal_nu_rho()
        self.get_normal_young()
        self.advect_upwind()
        self.set_BC()
        for _ in range(self.P.n_jacobi):
            self.solve_p_jacobi()
        self.update_uv()
        self.set_BC()
        self.solve_VOF_rudman()
        self.post_process_f()
        self.set_BC()

    def run(self, nsteps):
        for _ in range(nsteps):
            self.step()

    # --------------------------------------------------------------- diagnostics
    def mass(self):
        P = self.P
        return float(np.sum(self.F[1:P.nx + 1, 1:P.ny + 1], dtype=np.float64))

    def state(self):
        return {k: getattr(self, k).copy() for k in self.FIELDS}
