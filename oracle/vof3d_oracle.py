"""CPU oracle for the 3-D VOF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Order-exact fp32 NumPy restatement of the per-timestep path of ``/root/reference/3dvof.py``
(loop body 606-623 and the kernels it calls: 141-302, 351-547).  Same conventions as oracle/vof2d_oracle.py (IEEE
fp32, left to right, no FMA contraction, Python-scalar sub-expressions folded in double) and the same pin: the
unmodified text of 3dvof.py executed under ``oracle/refshim/taichi`` (``oracle/run_reference.py``) produced
``tests/golden/ref_3d_*.npz``, and ``tests/test_reference_pin_cpu.py`` requires this oracle to equal them bit for bit
after every kernel call and at the step snapshots.

Facts of the 3-D script that differ from 2-D: curvature is never computed (get_normal_young is
commented out, 304-332 / 607), so kappa == 0 and the CSF terms are exactly +-0; only ``-ic 1``
sets F (126-138); the FCT sweep order rotates with istep % 3 (351-363).
"""
from __future__ import annotations

import numpy as np

__all__ = ["Vof3DParams", "Vof3DOracle"]


class Vof3DParams:
    """Constants block, 3dvof.py:20-68."""

    def __init__(self, nx=200, ny=200, nz=200, Lx=0.1, Ly=0.1, Lz=0.1, rho_l=1000.0, rho_g=50.0,
                 nu_l=1.0e-6, nu_g=1.5e-5, sigma=0.007, gx=0, gy=-5, gz=0, dt=4e-6, n_jacobi=10):
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.Lx, self.Ly, self.Lz = float(Lx), float(Ly), float(Lz)
        self.rho_l, self.rho_g, self.nu_l, self.nu_g = float(rho_l), float(rho_g), float(nu_l), float(nu_g)
        self.sigma = float(sigma)
        self.gx, self.gy, self.gz = gx, gy, gz
        self.dt = float(dt)
        self.n_jacobi = int(n_jacobi)
        mk = lambda L, n: np.hstack((0.0, np.linspace(0, L, n + 1), L)).astype(np.float32)
        self.x, self.y, self.z = mk(self.Lx, self.nx), mk(self.Ly, self.ny), mk(self.Lz, self.nz)
        self.dx = float(self.x[3]) - float(self.x[2])
        self.dy = float(self.y[3]) - float(self.y[2])
        self.dz = float(self.z[3]) - float(self.z[2])      # 3dvof.py:65 (indexes with jmin; same element)
        self.dxi, self.dyi, self.dzi = 1 / self.dx, 1 / self.dy, 1 / self.dz

    @classmethod
    def scaled(cls, n, **kw):
        L = 0.1 * n / 200.0
        return cls(nx=n, ny=n, nz=n, Lx=L, Ly=L, Lz=L, **kw)


class Vof3DOracle:
    FIELDS = ("F", "u", "v", "w", "p", "rho", "nu", "u_star", "v_star", "w_star")

    def __init__(self, params: Vof3DParams | None = None, real=np.float32):
        self.P = P = params or Vof3DParams()
        self.real = R = real
        shape = (P.nx + 2, P.ny + 2, P.nz + 2)
        z = lambda: np.zeros(shape, dtype=R)
        self.F, self.Ftd = z(), z()
        self.u, self.v, self.w, self.u_star, self.v_star, self.w_star = z(), z(), z(), z(), z(), z()
        self.p, self.pt, self.rho, self.nu, self.kappa = z(), z(), z(), z(), z()
        self.rp, self.rm = z(), z()
        self.istep = 0
        self.courant_flags = 0
        c = lambda val: R(val)
        dx, dy, dz, dt = P.dx, P.dy, P.dz, P.dt
        self.c_dt = c(dt)
        self.c_dx, self.c_dy, self.c_dz = c(dx), c(dy), c(dz)
        self.c_dxi, self.c_dyi, self.c_dzi = c(P.dxi), c(P.dyi), c(P.dzi)
        self.c_dxi2, self.c_dyi2, self.c_dzi2 = c(P.dxi ** 2), c(P.dyi ** 2), c(P.dzi ** 2)
        self.c_vol = c(dx * dy * dz)
        self.c_dxdy = c(dx * dy)
        self.c_dt_yz, self.c_dt_xz, self.c_dt_xy = c(dt * dy * dz), c(dt * dx * dz), c(dt * dx * dy)
        self.c_sigma = c(P.sigma)
        self.c_rho_l, self.c_rho_g, self.c_nu_l, self.c_nu_g = c(P.rho_l), c(P.rho_g), c(P.nu_l), c(P.nu_g)
        self.c_gx, self.c_gy, self.c_gz = c(P.gx), c(P.gy), c(P.gz)
        self.c_cfl_x, self.c_cfl_y = c(0.25 * dx), c(0.25 * dy)

    def _var(self, a, b, c):
        R = self.real
        a, b, c = (np.asarray(t, dtype=R) for t in (a, b, c))
        return ((a + b) + c) - np.maximum(np.maximum(a, b), c) - np.minimum(np.minimum(a, b), c)

    # 3dvof.py:126-138 -- only ic 1 writes F
    def set_init_F(self, ic: int):
        P, R = self.P, self.real
        if ic != 1:
            return
        x = P.x[: P.nx + 2].astype(R)[:, None, None]
        y = P.y[: P.ny + 2].astype(R)[None, :, None]
        z = P.z[: P.nz + 2].astype(R)[None, None, :]
        m = ((x >= R(0.0)) & (x <= R(P.Lx / 3)) & (y >= R(0.0)) & (y <= R(P.Ly / 2)) & (z >= R(0.0)) & (z <= R(P.Lz / 3)))
        self.F[m] = R(1.0)

    # 3dvof.py:141-190 -- j-faces, then i-faces, then k-faces
    def set_BC(self):
        P = self.P
        nx, ny, nz = P.nx, P.ny, P.nz
        u, v, w, F, p, rho = self.u, self.v, self.w, self.F, self.p, self.rho
        u[:, 0, :] = u[:, 1, :]; v[:, 1, :] = 0; w[:, 0, :] = w[:, 1, :]
        F[:, 0, :] = F[:, 1, :]; p[:, 0, :] = p[:, 1, :]; rho[:, 0, :] = rho[:, 1, :]
        u[:, ny + 1, :] = u[:, ny, :]; v[:, ny + 1, :] = 0; w[:, ny + 1, :] = w[:, ny, :]
        F[:, ny + 1, :] = F[:, ny, :]; p[:, ny + 1, :] = p[:, ny, :]; rho[:, ny + 1, :] = rho[:, ny, :]
        u[1, :, :] = 0; v[0, :, :] = v[1, :, :]; w[0, :, :] = w[1, :, :]
        F[0, :, :] = F[1, :, :]; p[0, :, :] = p[1, :, :]; rho[0, :, :] = rho[1, :, :]
        u[nx + 1, :, :] = 0; v[nx + 1, :, :] = v[nx, :, :]; w[nx + 1, :, :] = w[nx, :, :]
        F[nx + 1, :, :] = F[nx, :, :]; p[nx + 1, :, :] = p[nx, :, :]; rho[nx + 1, :, :] = rho[nx, :, :]
        u[:, :, 0] = u[:, :, 1]; v[:, :, 0] = v[:, :, 1]; w[:, :, 1] = 0
        F[:, :, 0] = F[:, :, 1]; p[:, :, 0] = p[:, :, 1]; rho[:, :, 0] = rho[:, :, 1]
        u[:, :, nz + 1] = u[:, :, nz]; v[:, :, nz + 1] = v[:, :, nz]; w[:, :, nz + 1] = 0
        F[:, :, nz + 1] = F[:, :, nz]; p[:, :, nz + 1] = p[:, :, nz]; rho[:, :, nz + 1] = rho[:, :, nz]

    # 3dvof.py:199-204
    def cal_nu_rho(self):
        R = self.real
        Fc = self._var(R(0.0), R(1.0), self.F)
        self.rho[...] = self.c_rho_g * (R(1) - Fc) + self.c_rho_l * Fc
        self.nu[...] = self.c_nu_l * Fc + self.c_nu_g * (R(1.0) - Fc)

    # 3dvof.py:207-258
    def advect_upwind(self):
        P, R = self.P, self.real
        nx, ny, nz = P.nx, P.ny, P.nz
        u, v, w, F, kap, nu, rho = self.u, self.v, self.w, self.F, self.kappa, self.nu, self.rho
        dt, dxi, dyi, dzi = self.c_dt, self.c_dxi, self.c_dyi, self.c_dzi
        dxi2, dyi2, dzi2 = self.c_dxi2, self.c_dyi2, self.c_dzi2
        two = R(2)

        def sl(a, b):
            return slice(a, b)

        def sh(s, d):
            return slice(s.start + d, s.stop + d)
        # ---- u*: i in [2, nx], j in [1, ny], k in [1, nz]
        I, J, K = sl(2, nx + 1), sl(1, ny + 1), sl(1, nz + 1)
        uc = u[I, J, K]
        v_here = R(0.25) * (v[sh(I, -1), J, K] + v[sh(I, -1), sh(J, 1), K] + v[I, J, K] + v[I, sh(J, 1), K])
        w_here = R(0.25) * (w[sh(I, -1), J, K] + w[sh(I, -1), J, sh(K, 1)] + w[I, J, K] + w[I, J, sh(K, 1)])
        dudx = np.where(uc > 0, (uc - u[sh(I, -1), J, K]) * dxi, (u[sh(I, 1), J, K] - uc) * dxi)
        dudy = np.where(v_here > 0, (uc - u[I, sh(J, -1), K]) * dyi, (u[I, sh(J, 1), K] - uc) * dyi)
        dudz = np.where(w_here > 0, (uc - u[I, J, sh(K, -1)]) * dzi, (u[I, J, sh(K, 1)] - uc) * dzi)
        kappa_ave = (kap[I, J, K] + kap[sh(I, -1), J, K]) / R(2.0)
        fk = (-self.c_sigma) * (F[I, J, K] - F[sh(I, -1), J, K]) * kappa_ave / self.c_dx
        us = uc + dt * (
            nu[I, J, K] * (u[sh(I, -1), J, K] - two * uc + u[sh(I, 1), J, K]) * dxi2
            + nu[I, J, K] * (u[I, sh(J, -1), K] - two * uc + u[I, sh(J, 1), K]) * dyi2
            + nu[I, J, K] * (u[I, J, sh(K, -1)] - two * uc + u[I, J, sh(K, 1)]) * dzi2
            - uc * dudx - v_here * dudy - w_here * dudz
            + self.c_gx + fk * two / (rho[I, J, K] + rho[sh(I, -1), J, K]))
        # ---- v*: i in [1, nx], j in [2, ny], k in [1, nz]
        I2, J2, K2 = sl(1, nx + 1), sl(2, ny + 1), sl(1, nz + 1)
        vc = v[I2, J2, K2]
        u_here = R(0.25) * (u[I2, sh(J2, -1), K2] + u[I2, J2, K2] + u[sh(I2, 1), sh(J2, -1), K2] + u[sh(I2, 1), J2, K2])
        w_here2 = R(0.25) * (w[I2, sh(J2, -1), sh(K2, 1)] + w[I2, sh(J2, -1), K2] + w[I2, J2, K2] + w[I2, J2, sh(K2, 1)])
        dvdx = np.where(u_here > 0, (vc - v[sh(I2, -1), J2, K2]) * dxi, (v[sh(I2, 1), J2, K2] - vc) * dxi)
        dvdy = np.where(vc > 0, (vc - v[I2, sh(J2, -1), K2]) * dyi, (v[I2, sh(J2, 1), K2] - vc) * dyi)
        dvdz = np.where(w_here2 > 0, (vc - v[I2, J2, sh(K2, -1)]) * dzi, (v[I2, J2, sh(K2, 1)] - vc) * dzi)
        kappa_ave2 = (kap[I2, J2, K2] + kap[I2, sh(J2, -1), K2]) / R(2.0)
        fk2 = (-self.c_sigma) * (F[I2, J2, K2] - F[I2, sh(J2, -1), K2]) * kappa_ave2 / self.c_dy
        vs = vc + dt * (
            nu[I2, J2, K2] * (v[sh(I2, -1), J2, K2] - two * vc + v[sh(I2, 1), J2, K2]) * dxi2
            + nu[I2, J2, K2] * (v[I2, sh(J2, -1), K2] - two * vc + v[I2, sh(J2, 1), K2]) * dyi2
            + nu[I2, J2, K2] * (v[I2, J2, sh(K2, -1)] - two * vc + v[I2, J2, sh(K2, 1)]) * dzi2
            - u_here * dvdx - vc * dvdy - w_here2 * dvdz
            + self.c_gy + fk2 * two / (rho[I2, J2, K2] + rho[I2, sh(J2, -1), K2]))
        # ---- w*: i in [1, nx], j in [1, ny], k in [2, nz]
        I3, J3, K3 = sl(1, nx + 1), sl(1, ny + 1), sl(2, nz + 1)
        wc = w[I3, J3, K3]
        u_here3 = R(0.25) * (u[sh(I3, 1), J3, sh(K3, -1)] + u[I3, J3, sh(K3, -1)] + u[sh(I3, 1), J3, K3] + u[I3, J3, K3])
        v_here3 = R(0.25) * (v[I3, sh(J3, 1), sh(K3, -1)] + v[I3, J3, sh(K3, -1)] + v[I3, J3, K3] + v[I3, sh(J3, 1), K3])
        dwdx = np.where(u_here3 > 0, (wc - w[sh(I3, -1), J3, K3]) * dxi, (w[sh(I3, 1), J3, K3] - wc) * dxi)
        dwdy = np.where(v_here3 > 0, (wc - w[I3, sh(J3, -1), K3]) * dyi, (w[I3, sh(J3, 1), K3] - wc) * dyi)
        dwdz = np.where(wc > 0, (wc - w[I3, J3, sh(K3, -1)]) * dzi, (w[I3, J3, sh(K3, 1)] - wc) * dzi)
        kappa_ave3 = (kap[I3, J3, K3] + kap[I3, J3, sh(K3, -1)]) / R(2.0)
        fk3 = (-self.c_sigma) * (F[I3, J3, K3] - F[I3, J3, sh(K3, -1)]) * kappa_ave3 / self.c_dz
        ws = wc + dt * (
            nu[I3, J3, K3] * (w[sh(I3, -1), J3, K3] - two * wc + w[sh(I3, 1), J3, K3]) * dxi2
            + nu[I3, J3, K3] * (w[I3, sh(J3, -1), K3] - two * wc + w[I3, sh(J3, 1), K3]) * dyi2
            + nu[I3, J3, K3] * (w[I3, J3, sh(K3, -1)] - two * wc + w[I3, J3, sh(K3, 1)]) * dzi2
            - u_here3 * dwdx - v_here3 * dwdy - wc * dwdz
            + self.c_gz + fk3 * two / (rho[I3, J3, K3] + rho[I3, J3, sh(K3, -1)]))
        self.u_star[I, J, K] = us
        self.v_star[I2, J2, K2] = vs
        self.w_star[I3, J3, K3] = ws

    # 3dvof.py:261-283 -- one sweep
    def solve_p_jacobi(self):
        P, R = self.P, self.real
        nx, ny, nz = P.nx, P.ny, P.nz
        I, J, K = slice(1, nx + 1), slice(1, ny + 1), slice(1, nz + 1)
        Ip, Im = slice(2, nx + 2), slice(0, nx)
        Jp, Jm = slice(2, ny + 2), slice(0, ny)
        Kp, Km = slice(2, nz + 2), slice(0, nz)
        us, vs, ws, p = self.u_star, self.v_star, self.w_star, self.p
        rhs = self.rho[I, J, K] / self.c_dt * (
            (us[Ip, J, K] - us[I, J, K]) * self.c_dxi + (vs[I, Jp, K] - vs[I, J, K]) * self.c_dyi
            + (ws[I, J, Kp] - ws[I, J, K]) * self.c_dzi)
        ii = np.arange(1, nx + 1)[:, None, None]
        jj = np.arange(1, ny + 1)[None, :, None]
        kk = np.arange(1, nz + 1)[None, None, :]
        z0 = R(0.0)
        ae = np.where(ii != nx, self.c_dxi2, z0).astype(R); aw = np.where(ii != 1, self.c_dxi2, z0).astype(R)
        an = np.where(jj != ny, self.c_dyi2, z0).astype(R); a_s = np.where(jj != 1, self.c_dyi2, z0).astype(R)
        af = np.where(kk != nz, self.c_dzi2, z0).astype(R); ab = np.where(kk != 1, self.c_dzi2, z0).astype(R)
        ap = R(-1.0) * (ae + aw + an + a_s + ab + af)
        self.pt[I, J, K] = (rhs - ae * p[Ip, J, K] - aw * p[Im, J, K] - an * p[I, Jp, K] - a_s * p[I, Jm, K]
                            - af * p[I, J, Kp] - ab * p[I, J, Km]) / ap
        self.p[I, J, K] = self.pt[I, J, K]

    # 3dvof.py:286-302
    def update_uv(self):
        P, R = self.P, self.real
        nx, ny, nz = P.nx, P.ny, P.nz
        rho, p = self.rho, self.p
        I, J, K = slice(2, nx + 1), slice(1, ny + 1), slice(1, nz + 1)
        r = (rho[I, J, K] + rho[slice(1, nx), J, K]) * R(0.5)
        self.u[I, J, K] = self.u_star[I, J, K] - self.c_dt / r * (p[I, J, K] - p[slice(1, nx), J, K]) * self.c_dxi
        I2, J2, K2 = slice(1, nx + 1), slice(2, ny + 1), slice(1, nz + 1)
        r = (rho[I2, J2, K2] + rho[I2, slice(1, ny), K2]) * R(0.5)
        self.v[I2, J2, K2] = self.v_star[I2, J2, K2] - self.c_dt / r * (p[I2, J2, K2] - p[I2, slice(1, ny), K2]) * self.c_dyi
        I3, J3, K3 = slice(1, nx + 1), slice(1, ny + 1), slice(2, nz + 1)
        r = (rho[I3, J3, K3] + rho[I3, J3, slice(1, nz)]) * R(0.5)
        self.w[I3, J3, K3] = self.w_star[I3, J3, K3] - self.c_dt / r * (p[I3, J3, K3] - p[I3, J3, slice(1, nz)]) * self.c_dzi
        self.courant_flags = int(np.count_nonzero(self.u[I, J, K] * self.c_dt > self.c_cfl_x)
                                 + np.count_nonzero(self.v[I2, J2, K2] * self.c_dt > self.c_cfl_y)
                                 + np.count_nonzero(self.w[I3, J3, K3] * self.c_dt > self.c_cfl_x))   # 0.25*dx, 3dvof.py:301

    def _limit(self, q, pq):
        R = self.real
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(pq > 0, np.minimum(R(1), q / pq), R(0.0)).astype(R)

    def _fct_sweep(self, axis):
        """3dvof.py:366-427 (x), 430-492 (y), 495-541 (z).  The four loops; the scratch ax/ay/az, cx/cy/cz
        are kept as temporaries (their never-written ghosts are 0, as in the reference)."""
        P, R = self.P, self.real
        n = (P.nx, P.ny, P.nz)
        F, Ftd = self.F, self.Ftd
        vel = (self.u, self.v, self.w)[axis]
        dt, dx, dy, dz, vol = self.c_dt, self.c_dx, self.c_dy, self.c_dz, self.c_vol
        core = [slice(1, n[0] + 1), slice(1, n[1] + 1), slice(1, n[2] + 1)]

        def shifted(d):
            s = list(core)
            s[axis] = slice(1 + d, n[axis] + 1 + d)
            return tuple(s)
        c, m, pl = tuple(core), shifted(-1), shifted(1)
        zero = R(0)
        vc, vp = vel[c], vel[pl]
        lo_L = np.where(vc >= 0, vc * dt * F[m], vc * dt * F[c])          # flux through the minus face
        hi_L = np.where(vp >= 0, vp * dt * F[c], vp * dt * F[pl])         # flux through the plus face
        if axis == 0:
            dv = vol - self.c_dt_yz * (vp - vc)
            t = (F[c] + (lo_L - hi_L) * dy * dz / vol) * dx * dy * dz / dv
        elif axis == 1:
            dv = vol - self.c_dt_xz * (vp - vc)
            t = (F[c] + (zero - zero + lo_L - hi_L) * dy / self.c_dxdy) * dx * dy * dz / dv
        else:
            dv = vol - self.c_dt_xy * (vp - vc)
            t = (F[c] + (lo_L - hi_L) * dy * dx / vol) * dx * dy * dz / dv
        t = np.where((t > R(1.)) | (t < 0), self._var(R(0), R(1), t), t)
        Ftd[c] = t
        fmax = np.maximum(np.maximum(Ftd[c], Ftd[m]), Ftd[pl])
        fmin = np.minimum(np.minimum(Ftd[c], Ftd[m]), Ftd[pl])
        lo_H = np.where(vc <= 0, vc * dt * F[m], vc * dt * F[c])
        hi_H = np.where(vp <= 0, vp * dt * F[c], vp * dt * F[pl])
        a = np.zeros(F.shape, dtype=R)        # antidiffusive flux on the faces along `axis` (index = face)
        a[pl] = hi_H - hi_L
        a[c] = lo_H - lo_L
        qs = dz if axis == 2 else dx          # 3dvof.py:388/452 use dx, 521/527 use dz
        if axis == 2:
            pp = np.maximum(zero, a[c]) - np.minimum(zero, a[pl])
            pm = np.maximum(zero, a[pl]) - np.minimum(zero, a[c])
        else:   # x and y sweeps carry the (zero) second pair of the 2-D code
            pp = np.maximum(zero, a[c]) - np.minimum(zero, a[pl]) + np.maximum(zero, zero) - np.minimum(zero, zero)
            pm = np.maximum(zero, a[pl]) - np.minimum(zero, a[c]) + np.maximum(zero, zero) - np.minimum(zero, zero)
        self.rp[c] = self._limit((fmax - Ftd[c]) * qs, pp)
        self.rm[c] = self._limit((Ftd[c] - fmin) * qs, pm)
        rp, rm = self.rp, self.rm
        cf = np.zeros(F.shape, dtype=R)
        cf[pl] = np.where(a[pl] >= 0, np.minimum(rp[pl], rm[c]), np.minimum(rp[c], rm[pl]))
        if axis == 2:
            Fn = Ftd[c] - ((a[pl] * cf[pl] - a[c] * cf[c]) / dz) * dx * dy * dz / dv
        else:
            Fn = Ftd[c] - ((a[pl] * cf[pl] - a[c] * cf[c] + zero - zero) / dy) * dx * dy * dz / dv
        F[c] = self._var(R(0), R(1), Fn)

    def fct_x_sweep(self):
        self._fct_sweep(0)

    def fct_y_sweep(self):
        self._fct_sweep(1)

    def fct_z_sweep(self):
        self._fct_sweep(2)

    # 3dvof.py:351-363
    def solve_VOF_rudman(self):
        r = self.istep % 3
        order = ((0, 1, 2), (1, 2, 0), (2, 0, 1))[r]
        for ax in order:
            self._fct_sweep(ax)

    # 3dvof.py:544-547
    def post_process_f(self):
        R = self.real
        self.F[...] = self._var(self.F, R(0), R(1))

    # 3dvof.py:598-623
    def step(self):
        self.istep += 1
        self.cal_nu_rho()
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

    def mass(self):
        P = self.P
        return float(np.sum(self.F[1:P.nx + 1, 1:P.ny + 1, 1:P.nz + 1], dtype=np.float64))

    def state(self):
        return {k: getattr(self, k).copy() for k in self.FIELDS}
