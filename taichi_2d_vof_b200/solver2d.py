"""Host-side mirror of the reference's per-timestep interface (2dvof.py).

``VofSolver2D`` exposes the same zero-argument kernel names the reference's main loop calls
(2dvof.py:513-528) and the same field names (2dvof.py:53-89) with ``.to_numpy()`` /
``.from_numpy()``; everything executes in hand-written sm_100a kernels behind the C ABI
(include/vof.h).  No computation happens in Python and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import VofParams, check


def reference_params(nx=200, ny=200, Lx=0.1, Ly=0.1, rho_l=1000.0, rho_g=50.0, nu_l=1.0e-6,
                     nu_g=1.5e-5, sigma=0.007, gx=0, gy=-5, dt=4e-6, n_jacobi=10,
                     slab=None, halo=0, device=-1) -> VofParams:
    """The constants block of 2dvof.py:19-50.  dx, dy are derived exactly as the reference does
    (difference of two fp32 node coordinates, kept as a Python double; 2dvof.py:41-48)."""
    x = np.hstack((0.0, np.linspace(0, Lx, nx + 1), Lx)).astype(np.float32)
    y = np.hstack((0.0, np.linspace(0, Ly, ny + 1), Ly)).astype(np.float32)
    dx = float(x[3]) - float(x[2])
    dy = float(y[3]) - float(y[2])
    lo, hi = slab if slab else (0, 0)
    return VofParams(nx=nx, ny=ny, nz=0, Lx=Lx, Ly=Ly, Lz=0.0, dx=dx, dy=dy, dz=0.0, dt=dt,
                     rho_l=rho_l, rho_g=rho_g, nu_l=nu_l, nu_g=nu_g, sigma=sigma,
                     gx=float(gx), gy=float(gy), gz=0.0, n_jacobi=n_jacobi,
                     slab_lo=lo, slab_hi=hi, halo=halo, device=device)


def scaled_params(n, **kw) -> VofParams:
    """Constant-dx scaling for the large synthetic grids: L = 0.1 * n / 200 (risk 3)."""
    L = 0.1 * n / 200.0
    return reference_params(nx=n, ny=n, Lx=L, Ly=L, **kw)


class Field:
    """A module-global ``ti.field`` of the reference, living in device memory."""

    def __init__(self, solver, name):
        self._s = solver
        self.name = name
        self.fid = _lib.FIELD_IDS[name]

    @property
    def shape(self):
        return (self._s.nrows, self._s.ny + 2)

    def to_numpy(self, out=None) -> np.ndarray:
        if out is None:
            out = np.empty(self.shape, dtype=np.float32)
        elif out.shape != self.shape or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError(f"field {self.name}: out must be C-contiguous float32 {self.shape}")
        check(self._s._L.vof2d_field_get(self._s._h, self.fid, out.ctypes.data_as(C.c_void_p)))
        return out

    def to_numpy_async(self, out) -> "Field":
        """Non-stalling read into ``out`` (use ``_lib.pinned_empty``): returns at once, the time loop may go on;
        ``wait()`` makes ``out`` valid.  One read in flight per solver."""
        if out.shape != self.shape or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError(f"field {self.name}: out must be C-contiguous float32 {self.shape}")
        check(self._s._L.vof2d_field_get_async(self._s._h, self.fid, out.ctypes.data_as(C.c_void_p)))
        return self

    def wait(self):
        check(self._s._L.vof2d_field_get_wait(self._s._h))

    def from_numpy(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        if a.shape != self.shape:
            raise ValueError(f"field {self.name}: expected shape {self.shape}, got {a.shape}")
        check(self._s._L.vof2d_field_set(self._s._h, self.fid, a.ctypes.data_as(C.c_void_p)))

    def fill(self, value):
        check(self._s._L.vof2d_field_fill(self._s._h, self.fid, float(value)))

    def device_ptr(self):
        """(address of logical element (0, 0), pitch in floats, rows).  Valid until the next
        compute call on the solver (ping-pong buffers swap)."""
        dev, pitch, rows = C.c_void_p(), C.c_int64(), C.c_int64()
        check(self._s._L.vof2d_field_ptr(self._s._h, self.fid, C.byref(dev), C.byref(pitch), C.byref(rows)))
        return dev.value, pitch.value, rows.value

    def torch(self):
        """Zero-copy torch view of the logical (rows, ny+2) array (strided over the pitch)."""
        import torch
        addr, pitch, rows = self.device_ptr()

        class _Blob:
            __cuda_array_interface__ = {"shape": (rows * pitch,), "typestr": "<f4", "data": (addr, False),
                                        "version": 3, "strides": None}
        # expose from element (0,0); the last row only reaches ny+2 columns, so clip the length
        n = (rows - 1) * pitch + self._s.ny + 2
        _Blob.__cuda_array_interface__["shape"] = (n,)
        flat = torch.as_tensor(_Blob(), device=f"cuda:{self._s.device}")
        return flat.as_strided((rows, self._s.ny + 2), (pitch, 1))

    def __getitem__(self, idx):
        return self.to_numpy()[idx]


class VofSolver2D:
    """One simulation = one CUDA context-resident state, like the reference's module globals."""

    FIELDS = ("F", "u", "v", "p", "rho", "nu", "kappa", "u_star", "v_star")

    def __init__(self, params: VofParams | None = None, stream=None, arena=None, arena_bytes=0):
        self._L = _lib.lib()
        if params is None:
            params = reference_params()
        self._h = C.c_void_p()
        if arena is not None:
            check(self._L.vof2d_create_in(C.byref(params), C.c_void_p(arena), arena_bytes, C.byref(self._h)))
        else:
            check(self._L.vof2d_create(C.byref(params), C.byref(self._h)))
        self.P = VofParams()
        check(self._L.vof2d_get_params(self._h, C.byref(self.P)))
        self.nx, self.ny = self.P.nx, self.P.ny
        self.lo, self.hi, self.halo = self.P.slab_lo, self.P.slab_hi, self.P.halo
        self.nrows = (self.hi - self.lo + 1) + 2 * self.halo
        self.gi0 = self.lo - self.halo
        self.device = self.P.device if self.P.device >= 0 else self._current_device()
        self.istep = 0
        if stream is not None:
            self.set_stream(stream)
        for name in self.FIELDS:
            setattr(self, name, Field(self, name))

    @staticmethod
    def _current_device():
        try:
            import torch
            return torch.cuda.current_device()
        except Exception:
            return 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.vof2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- plumbing
    def set_stream(self, stream):
        handle = getattr(stream, "cuda_stream", stream)
        check(self._L.vof2d_set_stream(self._h, C.c_void_p(int(handle))))

    def synchronize(self):
        check(self._L.vof2d_synchronize(self._h))

    # ---- the reference's kernels, same names (2dvof.py)
    def set_init_F(self, ic: int):
        check(self._L.vof2d_set_init_F(self._h, int(ic)))

    def set_BC(self):
        check(self._L.vof2d_set_BC(self._h))

    def cal_nu_rho(self):
        check(self._L.vof2d_cal_nu_rho(self._h))

    def get_normal_young(self):
        check(self._L.vof2d_get_normal_young(self._h))

    def advect_upwind(self):
        check(self._L.vof2d_advect_upwind(self._h))

    def solve_p_jacobi(self, nsweeps: int = 1):
        check(self._L.vof2d_solve_p_jacobi(self._h, int(nsweeps)))

    def update_uv(self):
        check(self._L.vof2d_update_uv(self._h))

    def fct_x_sweep(self):
        check(self._L.vof2d_fct_x_sweep(self._h))

    def fct_y_sweep(self):
        check(self._L.vof2d_fct_y_sweep(self._h))

    def solve_VOF_rudman(self):
        check(self._L.vof2d_solve_VOF_rudman(self._h, int(self.istep)))

    def fct_forward(self, eps: float = 1.0e-4, istep: int | None = None):
        """``solve_VOF_rudman(t, eps_value)`` of the reference's test/forward_fct.py:254-264 (the stand-alone FCT
        variant) for step ``t`` (default: the solver's step counter, which it advances)."""
        t = self.istep if istep is None else int(istep)
        check(self._L.vof2d_fct_forward(self._h, t, float(eps)))
        if istep is None:
            self.istep += 1

    def post_process_f(self):
        check(self._L.vof2d_post_process_f(self._h))

    # ---- display kernels of the GUI loop (2dvof.py:458-492): each returns what `rgb_buf.to_numpy()` / `V.to_numpy()` holds
    def _display(self, view):
        out = np.empty((2 * self.nx, 2 * self.ny), dtype=np.float32)
        check(self._L.vof2d_display_field(self._h, view, out.ctypes.data_as(C.c_void_p)))
        return out

    def get_vof_field(self):
        return self._display(_lib.VOF_VIEW_VOF)

    def get_u_field(self):
        return self._display(_lib.VOF_VIEW_U)

    def get_v_field(self):
        return self._display(_lib.VOF_VIEW_V)

    def get_vnorm_field(self):
        return self._display(_lib.VOF_VIEW_VNORM)

    def interp_velocity(self):
        out = np.empty((self.nx + 2, self.ny + 2, 2), dtype=np.float32)
        check(self._L.vof2d_interp_velocity(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- the loop body
    def step_sequence(self):
        """2dvof.py:506-528 literally: one C-ABI call per reference kernel call."""
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

    def step(self, materialize_props=False, no_fusion=False):
        """Same observable result through the fused path (vof2d_step)."""
        self.istep += 1
        flags = (_lib.VOF_STEP_MATERIALIZE_PROPS if materialize_props else 0) | (_lib.VOF_STEP_NO_FUSION if no_fusion else 0)
        check(self._L.vof2d_step(self._h, self.istep, flags))

    def run(self, nsteps, materialize_props=False, no_fusion=False):
        flags = (_lib.VOF_STEP_MATERIALIZE_PROPS if materialize_props else 0) | (_lib.VOF_STEP_NO_FUSION if no_fusion else 0)
        check(self._L.vof2d_run(self._h, self.istep + 1, int(nsteps), flags))
        self.istep += int(nsteps)

    def step_host(self, u, v, p, F, out=None, materialize_props=False):
        """Host-buffer form: H2D(u, v, p, F) -> one step -> D2H(u, v, p, F).  Arrays are the
        logical (rows, ny+2) fp32 fields; ``out`` = 4 preallocated arrays (default: in place)."""
        self.istep += 1
        out = out or (u, v, p, F)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        flags = _lib.VOF_STEP_MATERIALIZE_PROPS if materialize_props else 0
        check(self._L.vof2d_step_host(self._h, self.istep, flags, ptr(u), ptr(v), ptr(p), ptr(F),
                                      ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3])))
        return out

    # ---- diagnostics
    def diagnostics(self, residual=True):
        mass, cfl, res, cnt = C.c_double(), C.c_float(), C.c_float(), C.c_int64()
        check(self._L.vof2d_diagnostics(self._h, C.byref(mass), C.byref(cfl),
                                        C.byref(res) if residual else None, C.byref(cnt)))
        return {"mass": mass.value, "max_cfl": cfl.value, "residual": res.value if residual else None,
                "courant_count": cnt.value}

    def mass(self):
        return self.diagnostics(residual=False)["mass"]

    def state(self):
        return {k: getattr(self, k).to_numpy() for k in self.FIELDS}

    def set_option(self, option: int, value: int):
        check(self._L.vof2d_set_option(self._h, int(option), int(value)))

    # ---- measurement support
    def launch_count(self):
        return int(self._L.vof2d_launch_count(self._h))

    def profile(self, enable=True):
        """False / 0: off; True / 1: CUDA-event spans around every kernel; k > 1: around the kernels of every k-th step."""
        check(self._L.vof2d_profile(self._h, int(enable)))

    def profile_read(self):
        """{kind: (total ms, launches)} of the spans recorded since profile(True)."""
        out = {}
        for k, name in enumerate(_lib.KERNEL_KINDS):
            ms, n = C.c_double(), C.c_int64()
            check(self._L.vof2d_profile_read(self._h, k, C.byref(ms), C.byref(n)))
            if n.value:
                out[name] = (ms.value, n.value)
        return out

    # ---- slabs
    def halo_ptr(self, name, side, send):
        dev, n = C.c_void_p(), C.c_int64()
        check(self._L.vof2d_halo_ptr(self._h, _lib.FIELD_IDS[name], side, 1 if send else 0, C.byref(dev), C.byref(n)))
        return dev.value, n.value

    def p2p_export(self):
        """(64-byte CUDA IPC handle of this context's arena, local rows)."""
        buf = (C.c_ubyte * 64)()
        n, nb = C.c_int64(), C.c_int64()
        check(self._L.vof2d_p2p_export(self._h, buf, C.byref(n), C.byref(nb)))
        return bytes(buf), n.value

    def p2p_arena(self):
        a = C.c_void_p()
        check(self._L.vof2d_p2p_arena(self._h, C.byref(a)))
        return a.value

    def p2p_connect(self, side, handle=None, arena_ptr=None, peer_nrows=0):
        hb = (C.c_ubyte * 64).from_buffer_copy(handle) if handle is not None else None
        check(self._L.vof2d_p2p_connect(self._h, side, hb, C.c_void_p(arena_ptr) if arena_ptr else None, int(peer_nrows)))

    def halo_exchange_p2p(self):
        check(self._L.vof2d_halo_exchange_p2p(self._h))

    def p2p_status(self):
        t = C.c_int()
        check(self._L.vof2d_p2p_status(self._h, C.byref(t)))
        return t.value

    def p2p_check(self):
        """Raises VofError if a halo exchange timed out or the neighbours are out of lockstep (synchronises)."""
        check(self._L.vof2d_p2p_check(self._h))

    def halo_push(self, name, side, peer_dst):
        check(self._L.vof2d_halo_push(self._h, _lib.FIELD_IDS[name], side, C.c_void_p(peer_dst)))


class VofStreamer2D:
    """Host-resident state, device-streamed steps (include/vof.h: vof2d_streamer_*): the (nx+2, ny+2) arrays stay
    in host memory and every step streams them through the GPU in ``n_slabs`` row slabs, upload / step / download
    overlapped.  Results equal :meth:`VofSolver2D.step_host` bit for bit; use pinned arrays for the overlap."""

    def __init__(self, params: VofParams | None = None, n_slabs: int = 16):
        self._L = _lib.lib()
        params = params if params is not None else reference_params()
        self._h = C.c_void_p()
        check(self._L.vof2d_streamer_create(C.byref(params), int(n_slabs), C.byref(self._h)))
        n, h, b = C.c_int(), C.c_int(), C.c_size_t()
        check(self._L.vof2d_streamer_info(self._h, C.byref(n), C.byref(h), C.byref(b)))
        self.n_slabs, self.halo, self.device_bytes = n.value, h.value, b.value
        self.nx, self.ny = params.nx, params.ny
        self.istep = 0

    def step_host(self, u, v, p, F, out=None, materialize_props=False):
        """One step of the host-resident state; ``out`` = 4 preallocated arrays (default: in place)."""
        self.istep += 1
        out = out or (u, v, p, F)
        shape = (self.nx + 2, self.ny + 2)
        for a in (u, v, p, F) + tuple(out):
            if a.shape != shape or a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"streamed step needs C-contiguous float32 arrays of shape {shape}")
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        flags = _lib.VOF_STEP_MATERIALIZE_PROPS if materialize_props else 0
        check(self._L.vof2d_streamer_step_host(self._h, self.istep, flags, ptr(u), ptr(v), ptr(p), ptr(F),
                                               ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3])))
        return out

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.vof2d_streamer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
