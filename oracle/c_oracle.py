"""ctypes front-end of oracle/libvof_oracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The C twin (oracle/vof2d_oracle.c) restates /root/reference/2dvof.py with the loop
structure of the reference's ti.cpu backend; this wrapper gives it the same method
names as oracle/vof2d_oracle.py so tests can run both side by side, and bench.py can
time it as the CPU baseline ("port").  Pinned to the reference run like the NumPy form (vof2d_oracle.py header;
tests/test_reference_pin_cpu.py::test_c_oracle_equals_reference_run_2d / _3d).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .vof2d_oracle import Vof2DParams

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _OParams(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32),
                ("Lx", C.c_double), ("Ly", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
                ("dt", C.c_double), ("rho_l", C.c_double), ("rho_g", C.c_double),
                ("nu_l", C.c_double), ("nu_g", C.c_double), ("sigma", C.c_double),
                ("gx", C.c_double), ("gy", C.c_double), ("n_jacobi", C.c_int32)]


def build(force=False):
    """Compile the C oracle in place (gcc only; a few seconds)."""
    so = os.path.join(_HERE, "libvof_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("vof2d_oracle.c", "vof3d_oracle.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ovof2d_create.restype = C.c_void_p
        L.ovof2d_create.argtypes = [C.POINTER(_OParams)]
        L.ovof2d_field.restype = C.POINTER(C.c_float)
        L.ovof2d_field.argtypes = [C.c_void_p, C.c_int]
        L.ovof2d_mass.restype = C.c_double
        L.ovof2d_mass.argtypes = [C.c_void_p]
        L.ovof2d_courant_flags.restype = C.c_long
        L.ovof2d_courant_flags.argtypes = [C.c_void_p]
        L.ovof2d_istep.restype = C.c_int
        L.ovof2d_istep.argtypes = [C.c_void_p]
        L.ovof2d_set_istep.argtypes = [C.c_void_p, C.c_int]
        L.ovof2d_set_init_F.argtypes = [C.c_void_p, C.c_int]
        L.ovof2d_run.argtypes = [C.c_void_p, C.c_int]
        L.ovof2d_threads.restype = C.c_int
        L.ovof_set_threads.argtypes = [C.c_int]
        L.ovof_set_threads.restype = None
        for name in ("destroy", "set_BC", "cal_nu_rho", "get_normal_young", "advect_upwind", "solve_p_jacobi",
                     "update_uv", "fct_x_sweep", "fct_y_sweep", "post_process_f", "solve_VOF_rudman", "step"):
            getattr(L, "ovof2d_" + name).argtypes = [C.c_void_p]
            getattr(L, "ovof2d_" + name).restype = None
        _LIB = L
    return _LIB


def use_all_host_cores():
    """Give the OpenMP loops every core this process may run on (torchrun sets OMP_NUM_THREADS=1 for its workers);
    returns the thread count."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().ovof_set_threads(n)
    return n


class Vof2DCOracle:
    FIELDS = ("F", "u", "v", "p", "rho", "nu", "kappa", "u_star", "v_star")

    def __init__(self, params: Vof2DParams | None = None):
        self.P = P = params or Vof2DParams()
        self._L = lib()
        op = _OParams(P.nx, P.ny, P.Lx, P.Ly, P.dx, P.dy, P.dt, P.rho_l, P.rho_g, P.nu_l, P.nu_g,
                      P.sigma, float(P.gx), float(P.gy), P.n_jacobi)
        self._h = C.c_void_p(self._L.ovof2d_create(C.byref(op)))
        shape = (P.nx + 2, P.ny + 2)
        for k, name in enumerate(self.FIELDS):
            ptr = self._L.ovof2d_field(self._h, k)
            setattr(self, name, np.ctypeslib.as_array(ptr, shape=shape))   # zero-copy views

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ovof2d_destroy(self._h)
            self._h = None

    @property
    def istep(self):
        return self._L.ovof2d_istep(self._h)

    @istep.setter
    def istep(self, val):
        self._L.ovof2d_set_istep(self._h, int(val))

    @property
    def courant_flags(self):
        return int(self._L.ovof2d_courant_flags(self._h))

    @staticmethod
    def threads():
        return int(lib().ovof2d_threads())

    def set_init_F(self, ic):
        self._L.ovof2d_set_init_F(self._h, int(ic))

    def run(self, nsteps):
        self._L.ovof2d_run(self._h, int(nsteps))

    def mass(self):
        return float(self._L.ovof2d_mass(self._h))

    def state(self):
        return {k: getattr(self, k).copy() for k in self.FIELDS}


def _bind(name):
    def call(self):
        getattr(self._L, "ovof2d_" + name)(self._h)
    call.__name__ = name
    return call


for _n in ("set_BC", "cal_nu_rho", "get_normal_young", "advect_upwind", "solve_p_jacobi", "update_uv",
           "fct_x_sweep", "fct_y_sweep", "post_process_f", "solve_VOF_rudman", "step"):
    setattr(Vof2DCOracle, _n, _bind(_n))


# ---------------------------------------------------------------------------------------------
# 3-D twin (oracle/vof3d_oracle.c)
# ---------------------------------------------------------------------------------------------
class _O3Params(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("Lx", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double), ("dt", C.c_double),
                ("rho_l", C.c_double), ("rho_g", C.c_double), ("nu_l", C.c_double), ("nu_g", C.c_double),
                ("sigma", C.c_double), ("gx", C.c_double), ("gy", C.c_double), ("gz", C.c_double),
                ("n_jacobi", C.c_int32)]


_LIB3_READY = False


def _lib3():
    global _LIB3_READY
    L = lib()
    if not _LIB3_READY:
        L.ovof3d_create.restype = C.c_void_p
        L.ovof3d_create.argtypes = [C.POINTER(_O3Params)]
        L.ovof3d_field.restype = C.POINTER(C.c_float)
        L.ovof3d_field.argtypes = [C.c_void_p, C.c_int]
        L.ovof3d_mass.restype = C.c_double
        L.ovof3d_mass.argtypes = [C.c_void_p]
        L.ovof3d_courant_flags.restype = C.c_long
        L.ovof3d_courant_flags.argtypes = [C.c_void_p]
        L.ovof3d_istep.restype = C.c_int
        L.ovof3d_istep.argtypes = [C.c_void_p]
        L.ovof3d_set_istep.argtypes = [C.c_void_p, C.c_int]
        L.ovof3d_set_init_F.argtypes = [C.c_void_p, C.c_int]
        L.ovof3d_run.argtypes = [C.c_void_p, C.c_int]
        for name in ("destroy", "set_BC", "cal_nu_rho", "advect_upwind", "solve_p_jacobi", "update_uv", "fct_x_sweep",
                     "fct_y_sweep", "fct_z_sweep", "post_process_f", "solve_VOF_rudman", "step"):
            getattr(L, "ovof3d_" + name).argtypes = [C.c_void_p]
            getattr(L, "ovof3d_" + name).restype = None
        _LIB3_READY = True
    return L


class Vof3DCOracle:
    FIELDS = ("F", "u", "v", "w", "p", "rho", "nu", "u_star", "v_star", "w_star")

    def __init__(self, params=None):
        from .vof3d_oracle import Vof3DParams
        self.P = P = params or Vof3DParams()
        self._L = _lib3()
        op = _O3Params(P.nx, P.ny, P.nz, P.Lx, P.Ly, P.Lz, P.dx, P.dy, P.dz, P.dt, P.rho_l, P.rho_g, P.nu_l, P.nu_g,
                       P.sigma, float(P.gx), float(P.gy), float(P.gz), P.n_jacobi)
        self._h = C.c_void_p(self._L.ovof3d_create(C.byref(op)))
        shape = (P.nx + 2, P.ny + 2, P.nz + 2)
        for k, name in enumerate(self.FIELDS):
            setattr(self, name, np.ctypeslib.as_array(self._L.ovof3d_field(self._h, k), shape=shape))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ovof3d_destroy(self._h)
            self._h = None

    @property
    def istep(self):
        return self._L.ovof3d_istep(self._h)

    @istep.setter
    def istep(self, val):
        self._L.ovof3d_set_istep(self._h, int(val))

    @property
    def courant_flags(self):
        return int(self._L.ovof3d_courant_flags(self._h))

    def set_init_F(self, ic):
        self._L.ovof3d_set_init_F(self._h, int(ic))

    def run(self, nsteps):
        self._L.ovof3d_run(self._h, int(nsteps))

    def mass(self):
        return float(self._L.ovof3d_mass(self._h))


def _bind3(name):
    def call(self):
        getattr(self._L, "ovof3d_" + name)(self._h)
    call.__name__ = name
    return call


for _n in ("set_BC", "cal_nu_rho", "advect_upwind", "solve_p_jacobi", "update_uv", "fct_x_sweep", "fct_y_sweep",
           "fct_z_sweep", "post_process_f", "solve_VOF_rudman", "step"):
    setattr(Vof3DCOracle, _n, _bind3(_n))
