"""Host-side mirror of the reference's 3-D per-timestep interface (3dvof.py): same kernel names and call
order (3dvof.py:606-623), same field names (3dvof.py:70-117), executed by libvof's sm_100a kernels."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import VofParams, check


def reference_params3d(nx=200, ny=200, nz=200, Lx=0.1, Ly=0.1, Lz=0.1, rho_l=1000.0, rho_g=50.0, nu_l=1.0e-6,
                       nu_g=1.5e-5, sigma=0.007, gx=0, gy=-5, gz=0, dt=4e-6, n_jacobi=10, slab=None, halo=0,
                       device=-1) -> VofParams:
    """Constants block of 3dvof.py:20-68; dx, dy, dz derived from the fp32 node arrays as the reference does."""
    mk = lambda L, n: np.hstack((0.0, np.linspace(0, L, n + 1), L)).astype(np.float32)
    x, y, z = mk(Lx, nx), mk(Ly, ny), mk(Lz, nz)
    lo, hi = slab if slab else (0, 0)
    return VofParams(nx=nx, ny=ny, nz=nz, Lx=Lx, Ly=Ly, Lz=Lz, dx=float(x[3]) - float(x[2]), dy=float(y[3]) - float(y[2]),
                     dz=float(z[3]) - float(z[2]), dt=dt, rho_l=rho_l, rho_g=rho_g, nu_l=nu_l, nu_g=nu_g, sigma=sigma,
                     gx=float(gx), gy=float(gy), gz=float(gz), n_jacobi=n_jacobi, slab_lo=lo, slab_hi=hi, halo=halo,
                     device=device)


def scaled_params3d(n, **kw) -> VofParams:
    L = 0.1 * n / 200.0
    return reference_params3d(nx=n, ny=n, nz=n, Lx=L, Ly=L, Lz=L, **kw)


class Field3:
    def __init__(self, solver, name):
        self._s, self.name, self.fid = solver, name, _lib.FIELD_IDS[name]

    @property
    def shape(self):
        return (self._s.nrows, self._s.ny + 2, self._s.nz + 2)

    def to_numpy(self, out=None):
        if out is None:
            out = np.empty(self.shape, dtype=np.float32)
        check(self._s._L.vof3d_field_get(self._s._h, self.fid, out.ctypes.data_as(C.c_void_p)))
        return out

    def to_numpy_async(self, out):
        """Non-stalling read into a pinned array (``_lib.pinned_empty``); ``wait()`` makes it valid."""
        if out.shape != self.shape or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError(f"field {self.name}: out must be C-contiguous float32 {self.shape}")
        check(self._s._L.vof3d_field_get_async(self._s._h, self.fid, out.ctypes.data_as(C.c_void_p)))
        return self

    def wait(self):
        check(self._s._L.vof3d_field_get_wait(self._s._h))

    def from_numpy(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        if a.shape != self.shape:
            raise ValueError(f"field {self.name}: expected shape {self.shape}, got {a.shape}")
        check(self._s._L.vof3d_field_set(self._s._h, self.fid, a.ctypes.data_as(C.c_void_p)))


class VofSolver3D:
    FIELDS = ("F", "u", "v", "w", "p", "rho", "nu", "u_star", "v_star", "w_star")

    def __init__(self, params: VofParams | None = None, stream=None):
        self._L = _lib.lib()
        params = params or reference_params3d()
        self._h = C.c_void_p()
        check(self._L.vof3d_create(C.byref(params), C.byref(self._h)))
        self.P = VofParams()
        check(self._L.vof3d_get_params(self._h, C.byref(self.P)))
        self.nx, self.ny, self.nz = self.P.nx, self.P.ny, self.P.nz
        self.lo, self.hi, self.halo = self.P.slab_lo, self.P.slab_hi, self.P.halo
        self.nrows = (self.hi - self.lo + 1) + 2 * self.halo
        self.istep = 0
        if stream is not None:
            check(self._L.vof3d_set_stream(self._h, C.c_void_p(int(getattr(stream, "cuda_stream", stream)))))
        for name in self.FIELDS:
            setattr(self, name, Field3(self, name))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.vof3d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self._L.vof3d_synchronize(self._h))

    def set_init_F(self, ic):
        check(self._L.vof3d_set_init_F(self._h, int(ic)))

    def solve_p_jacobi(self, nsweeps=1):
        check(self._L.vof3d_solve_p_jacobi(self._h, int(nsweeps)))

    def solve_VOF_rudman(self):
        check(self._L.vof3d_solve_VOF_rudman(self._h, int(self.istep)))

    def step_sequence(self):
        """3dvof.py:598-623 literally, one C-ABI call per reference kernel call."""
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

    def step(self, materialize_props=False, no_fusion=False):
        self.istep += 1
        flags = (_lib.VOF_STEP_MATERIALIZE_PROPS if materialize_props else 0) | (_lib.VOF_STEP_NO_FUSION if no_fusion else 0)
        check(self._L.vof3d_step(self._h, self.istep, flags))

    def run(self, nsteps, **kw):
        for _ in range(nsteps):
            self.step(**kw)

    def diagnostics(self):
        mass, cfl, cnt = C.c_double(), C.c_float(), C.c_int64()
        check(self._L.vof3d_diagnostics(self._h, C.byref(mass), C.byref(cfl), C.byref(cnt)))
        return {"mass": mass.value, "max_cfl": cfl.value, "courant_count": cnt.value}

    def mass(self):
        return self.diagnostics()["mass"]

    def set_option(self, option: int, value: int):
        """VOF_OPT_ADAPTIVE: 1 (default) second-generation kernels, 0 first generation; identical results."""
        check(self._L.vof3d_set_option(self._h, int(option), int(value)))

    def launch_count(self):
        return int(self._L.vof3d_launch_count(self._h))

    def halo_ptr(self, name, side, send):
        dev, n = C.c_void_p(), C.c_int64()
        check(self._L.vof3d_halo_ptr(self._h, _lib.FIELD_IDS[name], side, 1 if send else 0, C.byref(dev), C.byref(n)))
        return dev.value, n.value

    def halo_push(self, name, side, peer_dst):
        check(self._L.vof3d_halo_push(self._h, _lib.FIELD_IDS[name], side, C.c_void_p(peer_dst)))

    # ---- NVLink peer-store halo exchange (same protocol as the 2-D context)
    def p2p_export(self):
        buf = (C.c_ubyte * 64)()
        n, nb = C.c_int64(), C.c_int64()
        check(self._L.vof3d_p2p_export(self._h, buf, C.byref(n), C.byref(nb)))
        return bytes(buf), n.value

    def p2p_arena(self):
        a = C.c_void_p()
        check(self._L.vof3d_p2p_arena(self._h, C.byref(a)))
        return a.value

    def p2p_connect(self, side, handle=None, arena_ptr=None, peer_nrows=0):
        hb = (C.c_ubyte * 64).from_buffer_copy(handle) if handle is not None else None
        check(self._L.vof3d_p2p_connect(self._h, side, hb, C.c_void_p(arena_ptr) if arena_ptr else None, int(peer_nrows)))

    def halo_exchange_p2p(self):
        check(self._L.vof3d_halo_exchange_p2p(self._h))

    def p2p_check(self):
        check(self._L.vof3d_p2p_check(self._h))


def _bind(name):
    def call(self):
        check(getattr(self._L, "vof3d_" + name)(self._h))
    call.__name__ = name
    return call


for _n in ("set_BC", "cal_nu_rho", "advect_upwind", "update_uv", "fct_x_sweep", "fct_y_sweep", "fct_z_sweep", "post_process_f"):
    setattr(VofSolver3D, _n, _bind(_n))
