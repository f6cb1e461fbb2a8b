"""Row-slab decomposition over the GPUs of one NVSwitch box (new capability; the reference is
single-address-space).

Design: communication-avoiding deep halos.  The whole timestep has a dependency radius of
``n_jacobi + 5`` rows along i (the x-FCT reads u at i+3, u needs p after n sweeps, whose rhs reaches
n-1 rows further and reads u* one row up, whose CSF term reads kappa, i.e. F three rows away -- see
DESIGN.md section 5), so each rank keeps H >= n_jacobi + 5 ghost rows per side, recomputes
the few halo rows redundantly, and exchanges u, v, p, F ONCE per step (4 fields x H contiguous
pitched rows per neighbour) instead of once per sweep.  Physical-wall logic applies only on the
first / last rank; interior ranks see their neighbours' rows as ordinary cells.

Transport (2-D and 3-D): ``p2p`` = peer stores over NVLink by one fused kernel per step (csrc/vof_p2p.cuh: the
neighbours' arenas are mapped with CUDA IPC, hand-shakes are device-side flags, no NCCL call and no host
synchronisation on the data path); ``nccl`` = ``torch.distributed`` batched isend / irecv (gloo on CPU for the
host-logic tests).  The only collective is the diagnostics reduction (one all-gather of a 4-vector, when asked).
``partition`` / ``exchange`` are pure host logic and device-agnostic.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

HALO_FIELDS = ("F", "u", "v", "p")


def partition(nx: int, nranks: int) -> List[Tuple[int, int]]:
    """Owned global interior rows [lo, hi] (inclusive, 1-based) of every rank; remainders go to
    the low ranks so slab heights differ by at most one."""
    if nranks < 1 or nx < nranks:
        raise ValueError(f"cannot split {nx} rows over {nranks} ranks")
    base, rem = divmod(nx, nranks)
    out, lo = [], 1
    for r in range(nranks):
        n = base + (1 if r < rem else 0)
        out.append((lo, lo + n - 1))
        lo += n
    return out


def required_halo(n_jacobi: int) -> int:
    """Dependency radius of one step along i (matches VOF_SLAB_MIN_HALO for n_jacobi = 10):
    F_new(i) <- u_new(i+3) <- p_n(i+3) <- rhs(i+3+n-1) <- u*(i+3+n) <- kappa(i+3+n) <- F(i+5+n)."""
    return n_jacobi + 5


def halo_row_blocks(nrows: int, halo: int):
    """Local row ranges [a, b) of the four halo blocks of a slab with `nrows` local rows:
    (send_lo, recv_lo, send_hi, recv_hi).  Mirrors vof2d_halo_ptr."""
    H = halo
    return ((H, 2 * H), (0, H), (nrows - 2 * H, nrows - H), (nrows - H, nrows))


def exchange(dist, rank: int, nranks: int, send_lo: Sequence, recv_lo: Sequence, send_hi: Sequence,
             recv_hi: Sequence):
    """One halo exchange: every tensor in send_lo goes to rank-1's recv_hi, send_hi to rank+1's
    recv_lo.  Tensors are contiguous 1-D views; lists are per field in the same order on all ranks."""
    ops = []
    if rank > 0:
        for t in send_lo:
            ops.append(dist.P2POp(dist.isend, t, rank - 1))
        for t in recv_lo:
            ops.append(dist.P2POp(dist.irecv, t, rank - 1))
    if rank < nranks - 1:
        for t in send_hi:
            ops.append(dist.P2POp(dist.isend, t, rank + 1))
        for t in recv_hi:
            ops.append(dist.P2POp(dist.irecv, t, rank + 1))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class SlabSolver2D:
    """One rank's slab of a global (nx, ny) domain.  ``step()`` = halo exchange + the fused step."""

    def __init__(self, global_params_fn, nx: int, rank: int, nranks: int, dist=None, halo: int | None = None,
                 n_jacobi: int = 10, device: int = -1, transport: str = "p2p", solver_cls=None, halo_fields=HALO_FIELDS):
        import torch
        if solver_cls is None:
            from .solver2d import VofSolver2D as solver_cls
        self.halo_fields = halo_fields
        self.rank, self.nranks, self.dist = rank, nranks, dist
        self.parts = partition(nx, nranks)
        self.lo, self.hi = self.parts[rank]
        H = halo if halo is not None else max(required_halo(n_jacobi), 16)
        if nranks == 1:
            params = global_params_fn(slab=None, halo=0, device=device)
        else:
            params = global_params_fn(slab=(self.lo, self.hi), halo=H, device=device)
        self.stream = torch.cuda.Stream(device=device if device >= 0 else None)
        self.solver = solver_cls(params, stream=self.stream)
        self.halo = self.solver.halo
        self._views = None
        self._steps = 0
        self.check_every = 256          # steps between p2p health checks (each one synchronises the stream); 0 = never
        import inspect
        self._diag_takes_residual = "residual" in inspect.signature(self.solver.diagnostics).parameters
        self.transport = transport if nranks > 1 else "none"
        if self.transport == "p2p":
            # map the neighbours' arenas (CUDA IPC): halo rows are then stored straight over NVLink by
            # vof2d_halo_exchange_p2p, hand-shaking through device-side flags -- no NCCL call per step
            mine = self.solver.p2p_export()
            allh = [None] * nranks
            dist.all_gather_object(allh, mine)
            if rank > 0:
                self.solver.p2p_connect(0, handle=allh[rank - 1][0], peer_nrows=allh[rank - 1][1])
            if rank < nranks - 1:
                self.solver.p2p_connect(1, handle=allh[rank + 1][0], peer_nrows=allh[rank + 1][1])
            dist.barrier()

    def _halo_views(self):
        """torch views of the 4 x 4 halo blocks (whole pitched rows, contiguous)."""
        import torch
        s = self.solver
        views = {}
        for side, has_nbr in ((0, self.rank > 0), (1, self.rank < self.nranks - 1)):
            for send in (1, 0):
                lst = []
                if has_nbr:
                    for name in self.halo_fields:
                        addr, n = s.halo_ptr(name, side, send)

                        class _Blob:
                            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (addr, False),
                                                        "version": 3, "strides": None}
                        lst.append(torch.as_tensor(_Blob(), device=f"cuda:{torch.cuda.current_device()}"))
                views[(side, send)] = lst
        return views

    def exchange_halos(self):
        if self.nranks == 1:
            return
        if self.transport == "p2p":
            self.solver.halo_exchange_p2p()
            return
        import torch
        v = self._halo_views()   # re-queried every step: F and p ping-pong between two buffers
        with torch.cuda.stream(self.stream):
            exchange(self.dist, self.rank, self.nranks, v[(0, 1)], v[(0, 0)], v[(1, 1)], v[(1, 0)])

    def set_init_F(self, ic):
        self.solver.set_init_F(ic)

    def step(self):
        self.exchange_halos()
        self.solver.step()
        self._steps += 1
        if self.check_every and self._steps % self.check_every == 0:
            self.check()

    def run(self, nsteps: int):
        if self.nranks == 1 and hasattr(self.solver, "run"):
            self.solver.run(nsteps)         # one GPU: CUDA-graph replay of whole steps
            return
        for _ in range(nsteps):
            self.step()

    @property
    def istep(self):
        return self.solver.istep

    def check(self):
        """Raises VofError if a peer-store halo exchange timed out (dead neighbour) or the ranks fell out of lockstep;
        synchronises this rank's stream.  Called by diagnostics() and every ``check_every`` steps."""
        if self.transport == "p2p":
            self.solver.p2p_check()

    def owned(self, name):
        """This rank's owned interior rows of a field, as numpy (rows lo..hi, all columns)."""
        a = getattr(self.solver, name).to_numpy()
        H = self.solver.halo if self.nranks > 1 else 1
        return a[H:a.shape[0] - H]

    def diagnostics(self, residual=True):
        """Global diagnostics of the decomposed run ("one allreduce per step of {sum F (fp64), max CFL,
        Jacobi residual, Courant count}"; the reference's only diagnostic is the Courant print, 2dvof.py:274-280).
        Every rank reduces its OWNED rows on the device (warp shuffles + one atomic per block), then one collective
        combines the per-rank 4-vectors: sum of volumes and Courant counts, max of CFL and residual."""
        self.check()
        d = self.solver.diagnostics(residual=residual) if self._diag_takes_residual else self.solver.diagnostics()
        if self.nranks == 1:
            return d
        import torch
        vec = torch.tensor([d["mass"], d["max_cfl"], d.get("residual") or 0.0, float(d["courant_count"])],
                           dtype=torch.float64, device=self._device())
        out = torch.empty((self.nranks, 4), dtype=torch.float64, device=vec.device)
        self.dist.all_gather_into_tensor(out, vec)
        out = out.cpu()
        res = {"mass": float(out[:, 0].sum()), "max_cfl": float(out[:, 1].max()), "courant_count": int(out[:, 3].sum()),
               "mass_per_rank": [float(x) for x in out[:, 0]]}
        if "residual" in d:
            res["residual"] = float(out[:, 2].max()) if residual else None
        return res

    def mass(self):
        return self.diagnostics(residual=False)["mass"]

    def _device(self):
        import torch
        if self.dist is not None and self.dist.get_backend() == "gloo":
            return torch.device("cpu")
        return torch.device(f"cuda:{self.solver.device}") if hasattr(self.solver, "device") else torch.device("cuda")

    def gather(self, name):
        """The global field on rank 0 (None elsewhere): owned rows of every rank plus the two physical ghost rows."""
        import numpy as np
        import torch
        a = getattr(self.solver, name).to_numpy()
        if self.nranks == 1:
            return a
        H = self.solver.halo
        lo = H - (1 if self.rank == 0 else 0)
        hi = a.shape[0] - H + (1 if self.rank == self.nranks - 1 else 0)
        mine = np.ascontiguousarray(a[lo:hi])
        tall = max(h - l + 1 for l, h in self.parts) + 1
        dev = self._device()
        pad = torch.zeros((tall,) + mine.shape[1:], dtype=torch.float32, device=dev)
        pad[: mine.shape[0]] = torch.from_numpy(mine).to(dev)
        outs = [torch.empty_like(pad) for _ in range(self.nranks)] if self.rank == 0 else None
        self.dist.gather(pad, outs, dst=0)
        if self.rank != 0:
            return None
        rows = []
        for r, (l, h) in enumerate(self.parts):
            n = h - l + 1 + (1 if r == 0 else 0) + (1 if r == self.nranks - 1 else 0)
            rows.append(outs[r][:n].cpu().numpy())
        return np.concatenate(rows, axis=0)

    def scatter(self, name, global_array):
        """Load this rank's local rows (owned + halo) of a field from the global array (every rank passes the same array)."""
        import numpy as np
        s = self.solver
        if self.nranks == 1:
            getattr(s, name).from_numpy(global_array)
            return
        gi0 = s.lo - s.halo
        loc = np.zeros((s.nrows,) + tuple(global_array.shape[1:]), np.float32)
        g0, g1 = max(gi0, 0), min(gi0 + s.nrows, global_array.shape[0])
        loc[g0 - gi0:g1 - gi0] = global_array[g0:g1]
        getattr(s, name).from_numpy(loc)


def slab_parity_check(dist, rank, world, device, nx=2048, ny=640, steps=6, ic=3, transport="p2p", three_d=False, n=96):
    """Correctness signal for multi-GPU runs (bench.py, tests): every rank steps its slab of a small global problem,
    rank 0 also runs the whole domain on its own GPU, and every owned row must equal it bit for bit.  Returns a dict
    on rank 0 (None elsewhere)."""
    import numpy as np
    if three_d:
        from .solver3d import VofSolver3D as cls, reference_params3d
        n = max(n, 24 * world)                 # every slab must be at least as thick as its halo (16 planes)
        L = 0.1 * n / 200

        def params_fn(slab, halo, device):
            return reference_params3d(nx=n, ny=n, nz=n, Lx=L, Ly=L, Lz=L, slab=slab, halo=halo, device=device)
        nx, fields, ic = n, ("F", "u", "v", "w", "p"), 1
    else:
        from .solver2d import VofSolver2D as cls, reference_params

        def params_fn(slab, halo, device):
            return reference_params(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200, slab=slab, halo=halo, device=device)
        fields = ("F", "u", "v", "p")
    s = SlabSolver2D(params_fn, nx, rank, world, dist=dist, device=device, transport=transport, solver_cls=cls, halo_fields=fields)
    s.set_init_F(ic)
    s.run(steps)
    d = s.diagnostics(residual=False)
    full = None
    if rank == 0:
        full = cls(params_fn(None, 0, device))
        full.set_init_F(ic)
        for _ in range(steps):
            full.step()
    bad = {}
    for name in fields:
        g = s.gather(name)
        if rank == 0:
            ref = getattr(full, name).to_numpy()
            bad[name] = int(np.count_nonzero(g != ref))
    if rank != 0:
        return None
    m_single = full.mass()
    return {"problem": (f"{n}^3 dam break" if three_d else f"{nx} x {ny} -ic {ic}") + f", {steps} steps, {world} slabs vs one GPU",
            "transport": s.transport, "fields": list(fields), "cells_differing": bad, "identical": all(v == 0 for v in bad.values()),
            "global_volume_allreduced": d["mass"], "single_gpu_volume": m_single,
            "volume_rel_diff": abs(d["mass"] - m_single) / m_single if m_single else 0.0}


class LocalSlabGroup:
    """Several slab contexts driven by ONE process (same device or peer-mapped devices): the halo
    exchange is vof2d_halo_push (a device-to-device copy into the neighbour's halo rows).  Used by
    the single-GPU parity test of the decomposition; the multi-process runner is SlabSolver2D."""

    def __init__(self, params_fn, nx: int, nslabs: int, halo: int | None = None, n_jacobi: int = 10, devices=None,
                 solver_cls=None, halo_fields=HALO_FIELDS, p2p=False):
        if solver_cls is None:
            from .solver2d import VofSolver2D as solver_cls
        self.parts = partition(nx, nslabs)
        H = halo if halo is not None else max(required_halo(n_jacobi), 16)
        devices = devices or [0] * nslabs
        self.solvers = [solver_cls(params_fn(slab=self.parts[r] if nslabs > 1 else None, halo=H if nslabs > 1 else 0,
                                             device=devices[r])) for r in range(nslabs)]
        self.nslabs = nslabs
        self.halo_fields = halo_fields
        self.p2p = p2p and nslabs > 1
        if self.p2p:
            for r in range(nslabs):
                if r > 0:
                    self.solvers[r].p2p_connect(0, arena_ptr=self.solvers[r - 1].p2p_arena(), peer_nrows=self.solvers[r - 1].nrows)
                if r < nslabs - 1:
                    self.solvers[r].p2p_connect(1, arena_ptr=self.solvers[r + 1].p2p_arena(), peer_nrows=self.solvers[r + 1].nrows)

    def exchange_halos(self):
        S = self.solvers
        if self.p2p:                   # device-side hand-shake: no host synchronisation at all
            for s in S:
                s.halo_exchange_p2p()
            return
        for s in S:
            s.synchronize()            # the producers of the rows about to be copied
        for r in range(self.nslabs - 1):
            lo, hi = S[r], S[r + 1]    # lo's upper side (1) faces hi's lower side (0)
            for name in self.halo_fields:
                dst_hi, _ = hi.halo_ptr(name, 0, send=False)
                lo.halo_push(name, 1, dst_hi)
                dst_lo, _ = lo.halo_ptr(name, 1, send=False)
                hi.halo_push(name, 0, dst_lo)
        for s in S:
            s.synchronize()

    def set_init_F(self, ic):
        for s in self.solvers:
            s.set_init_F(ic)

    def step(self):
        if self.nslabs > 1:
            self.exchange_halos()
        for s in self.solvers:
            s.step()

    def gather(self, name):
        """The global (nx+2, ny+2) field assembled from the owned rows (+ the two physical ghost rows)."""
        import numpy as np
        rows = []
        for r, s in enumerate(self.solvers):
            a = getattr(s, name).to_numpy()
            H = s.halo
            lo = H - (1 if r == 0 else 0)
            hi = a.shape[0] - H + (1 if r == self.nslabs - 1 else 0)
            rows.append(a[lo:hi])
        return np.concatenate(rows, axis=0)
