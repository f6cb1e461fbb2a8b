"""ctypes binding of libvof.so (the C ABI declared in include/vof.h).

There is NO CPU fallback: if the shared library is missing this module raises, and
``vof2d_create`` itself fails without an sm_100-class CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvof.so")
CSRC = os.path.join(_HERE, "csrc")

# field ids (include/vof.h)
VOF_F, VOF_U, VOF_V, VOF_P, VOF_RHO, VOF_NU, VOF_KAPPA, VOF_USTAR, VOF_VSTAR, VOF_W, VOF_WSTAR = range(11)
FIELD_IDS = {"F": VOF_F, "u": VOF_U, "v": VOF_V, "p": VOF_P, "rho": VOF_RHO, "nu": VOF_NU,
             "kappa": VOF_KAPPA, "u_star": VOF_USTAR, "v_star": VOF_VSTAR, "w": VOF_W, "w_star": VOF_WSTAR}
KERNEL_KINDS = ("props", "kappa", "advect", "bc", "rhs", "jacobi", "project", "fct_x", "fct_y", "post", "halo", "tile")
VOF_OPT_JACOBI_TB = 0
VOF_OPT_FCT_X_COLS = 1
VOF_OPT_ADVECT_COLS = 2
VOF_OPT_ADAPTIVE = 3
VOF_OPT_CHUNK_CAP = 4
VOF_OPT_JACOBI_MAXT = 5
VOF_OPT_JACOBI_PK = 6
VOF_OPT_JACOBI_ROWS = 7
VOF_OPT_JACOBI_LONG_PCT = 8
VOF_OPT_PRESSURE_SOLVER = 9
VOF_OPT_PACKED = 10
VOF_OPT_TILE = 11
VOF_OPT_FAST_MATH = 12
VOF_OPT_BARE_DIV = 13
VOF_VIEW_VOF, VOF_VIEW_U, VOF_VIEW_V, VOF_VIEW_VNORM = 0, 1, 2, 3
VOF_STEP_MATERIALIZE_PROPS = 1
VOF_STEP_NO_FUSION = 2
VOF_SLAB_MIN_HALO = 15


class VofParams(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("Lx", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
                ("dt", C.c_double),
                ("rho_l", C.c_double), ("rho_g", C.c_double),
                ("nu_l", C.c_double), ("nu_g", C.c_double),
                ("sigma", C.c_double),
                ("gx", C.c_double), ("gy", C.c_double), ("gz", C.c_double),
                ("n_jacobi", C.c_int32),
                ("slab_lo", C.c_int32), ("slab_hi", C.c_int32), ("halo", C.c_int32),
                ("device", C.c_int32)]


class VofError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libvof error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile libvof.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if os.path.isfile(os.path.join(CSRC, f))]
    srcs.append(os.path.join(_HERE, "..", "include", "vof.h"))
    stale = (not os.path.exists(LIB_PATH)
             or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs if os.path.exists(s)))
    if force or stale:
        subprocess.run(["make", "-C", CSRC, "-s", f"-j{min(8, os.cpu_count() or 1)}"] + (["-B"] if force else []), check=True)
    return LIB_PATH


_LIB = None

# every symbol include/vof.h declares: name -> (restype, argtypes)
_P = C.POINTER
_ctx = C.c_void_p
SIGNATURES = {
    "vof_last_error": (C.c_char_p, []),
    "vof_abi_version": (C.c_int, []),
    "vof_default_params": (None, [_P(VofParams)]),
    "vof2d_arena_bytes": (C.c_size_t, [_P(VofParams)]),
    "vof2d_create": (C.c_int, [_P(VofParams), _P(_ctx)]),
    "vof2d_create_in": (C.c_int, [_P(VofParams), C.c_void_p, C.c_size_t, _P(_ctx)]),
    "vof2d_destroy": (C.c_int, [_ctx]),
    "vof2d_set_stream": (C.c_int, [_ctx, C.c_void_p]),
    "vof2d_synchronize": (C.c_int, [_ctx]),
    "vof2d_get_params": (C.c_int, [_ctx, _P(VofParams)]),
    "vof2d_set_init_F": (C.c_int, [_ctx, C.c_int]),
    "vof2d_set_BC": (C.c_int, [_ctx]),
    "vof2d_cal_nu_rho": (C.c_int, [_ctx]),
    "vof2d_get_normal_young": (C.c_int, [_ctx]),
    "vof2d_advect_upwind": (C.c_int, [_ctx]),
    "vof2d_solve_p_jacobi": (C.c_int, [_ctx, C.c_int]),
    "vof2d_update_uv": (C.c_int, [_ctx]),
    "vof2d_fct_x_sweep": (C.c_int, [_ctx]),
    "vof2d_fct_y_sweep": (C.c_int, [_ctx]),
    "vof2d_solve_VOF_rudman": (C.c_int, [_ctx, C.c_int]),
    "vof2d_post_process_f": (C.c_int, [_ctx]),
    "vof2d_fct_forward": (C.c_int, [_ctx, C.c_int, C.c_float]),
    "vof2d_display_field": (C.c_int, [_ctx, C.c_int, C.c_void_p]),
    "vof2d_display_field_dev": (C.c_int, [_ctx, C.c_int, C.c_void_p]),
    "vof2d_interp_velocity": (C.c_int, [_ctx, C.c_void_p]),
    "vof2d_step": (C.c_int, [_ctx, C.c_int, C.c_uint]),
    "vof2d_run": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_uint]),
    "vof2d_step_host": (C.c_int, [_ctx, C.c_int, C.c_uint] + [C.c_void_p] * 8),
    "vof2d_streamer_create": (C.c_int, [_P(VofParams), C.c_int, _P(C.c_void_p)]),
    "vof2d_streamer_destroy": (C.c_int, [C.c_void_p]),
    "vof2d_streamer_info": (C.c_int, [C.c_void_p, _P(C.c_int), _P(C.c_int), _P(C.c_size_t)]),
    "vof2d_streamer_step_host": (C.c_int, [C.c_void_p, C.c_int, C.c_uint] + [C.c_void_p] * 8),
    "vof2d_field_ptr": (C.c_int, [_ctx, C.c_int, _P(C.c_void_p), _P(C.c_int64), _P(C.c_int64)]),
    "vof2d_field_get": (C.c_int, [_ctx, C.c_int, C.c_void_p]),
    "vof2d_field_set": (C.c_int, [_ctx, C.c_int, C.c_void_p]),
    "vof2d_field_fill": (C.c_int, [_ctx, C.c_int, C.c_float]),
    "vof2d_field_get_async": (C.c_int, [_ctx, C.c_int, C.c_void_p]),
    "vof2d_field_get_wait": (C.c_int, [_ctx]),
    "vof_pinned_alloc": (C.c_int, [C.c_size_t, _P(C.c_void_p)]),
    "vof_pinned_free": (C.c_int, [C.c_void_p]),
    "vof2d_diagnostics": (C.c_int, [_ctx, _P(C.c_double), _P(C.c_float), _P(C.c_float), _P(C.c_int64)]),
    "vof2d_launch_count": (C.c_int64, [_ctx]),
    "vof2d_profile": (C.c_int, [_ctx, C.c_int]),
    "vof2d_profile_read": (C.c_int, [_ctx, C.c_int, _P(C.c_double), _P(C.c_int64)]),
    "vof2d_set_option": (C.c_int, [_ctx, C.c_int, C.c_int]),
    "vof2d_halo_rows": (C.c_int, [_ctx, _P(C.c_int), _P(C.c_int64)]),
    "vof2d_halo_ptr": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, _P(C.c_void_p), _P(C.c_int64)]),
    "vof2d_halo_push": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_void_p]),
    "vof2d_p2p_export": (C.c_int, [_ctx, C.c_void_p, _P(C.c_int64), _P(C.c_int64)]),
    "vof2d_p2p_connect": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]),
    "vof2d_p2p_arena": (C.c_int, [_ctx, _P(C.c_void_p)]),
    "vof2d_halo_exchange_p2p": (C.c_int, [_ctx]),
    "vof2d_p2p_status": (C.c_int, [_ctx, _P(C.c_int)]),
    "vof2d_p2p_check": (C.c_int, [_ctx]),
    # ---- 3-D
    "vof3d_arena_bytes": (C.c_size_t, [_P(VofParams)]),
    "vof3d_create": (C.c_int, [_P(VofParams), _P(_ctx)]),
    "vof3d_destroy": (C.c_int, [_ctx]),
    "vof3d_set_stream": (C.c_int, [_ctx, C.c_void_p]),
    "vof3d_synchronize": (C.c_int, [_ctx]),
    "vof3d_get_params": (C.c_int, [_ctx, _P(VofParams)]),
    "vof3d_set_init_F": (C.c_int, [_ctx, C.c_int]),
    "vof3d_set_BC": (C.c_int, [_ctx]),
    "vof3d_cal_nu_rho": (C.c_int, [_ctx]),
    "vof3d_advect_upwind": (C.c_int, [_ctx]),
    "vof3d_solve_p_jacobi": (C.c_int, [_ctx, C.c_int]),
    "vof3d_update_uv": (C.c_int, [_ctx]),
    "vof3d_fct_x_sweep": (C.c_int, [_ctx]),
    "vof3d_fct_y_sweep": (C.c_int, [_ctx]),
    "vof3d_fct_z_sweep": (C.c_int, [_ctx]),
    "vof3d_solve_VOF_rudman": (C.c_int, [_ctx, C.c_int]),
    "vof3d_post_process_f": (C.c_int, [_ctx]),
    "vof3d_step": (C.c_int, [_ctx, C.c_int, C.c_uint]),
    "vof3d_run": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_uint]),
    "vof3d_field_ptr": (C.c_int, [_ctx, C.c_int, _P(C.c_void_p), _P(C.c_int64), _P(C.c_int64), _P(C.c_int64)]),
    "vof3d_field_get": (C.c_int, [_ctx, C.c_int, C.c_void_p]),
    "vof3d_field_set": (C.c_int, [_ctx, C.c_int, C.c_void_p]),
    "vof3d_field_get_async": (C.c_int, [_ctx, C.c_int, C.c_void_p]),
    "vof3d_field_get_wait": (C.c_int, [_ctx]),
    "vof3d_diagnostics": (C.c_int, [_ctx, _P(C.c_double), _P(C.c_float), _P(C.c_int64)]),
    "vof3d_launch_count": (C.c_int64, [_ctx]),
    "vof3d_set_option": (C.c_int, [_ctx, C.c_int, C.c_int]),
    "vof3d_halo_ptr": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, _P(C.c_void_p), _P(C.c_int64)]),
    "vof3d_halo_push": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_void_p]),
    "vof3d_p2p_export": (C.c_int, [_ctx, C.c_void_p, _P(C.c_int64), _P(C.c_int64)]),
    "vof3d_p2p_connect": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]),
    "vof3d_p2p_arena": (C.c_int, [_ctx, _P(C.c_void_p)]),
    "vof3d_halo_exchange_p2p": (C.c_int, [_ctx]),
    "vof3d_p2p_check": (C.c_int, [_ctx]),
}


def lib():
    """Load libvof.so.  Raises (never degrades to a CPU path) if it is absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(libvof is CUDA-only; there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)   # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def pinned_empty(shape, dtype="float32"):
    """A page-locked NumPy array (cudaHostAlloc through the C ABI): the target of the non-stalling field reads."""
    import numpy as np
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = C.c_void_p()
    check(lib().vof_pinned_alloc(n, C.byref(ptr)))
    buf = (C.c_char * n).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    _PINNED[arr.__array_interface__["data"][0]] = ptr.value
    return arr


def pinned_free(arr):
    ptr = _PINNED.pop(arr.__array_interface__["data"][0], None)
    if ptr:
        check(lib().vof_pinned_free(C.c_void_p(ptr)))


_PINNED = {}


def check(rc: int):
    if rc != 0:
        raise VofError(rc, lib().vof_last_error().decode("utf-8", "replace"))
