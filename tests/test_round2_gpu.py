"""Round-2 GPU tests: the BASELINE configurations at their full sizes against the C oracle, the multi-process slab
runs as driver-collected tests (spawned with torchrun when >= 2 GPUs are visible), and regressions for the
advisor's findings (halo depth n_jacobi + 5, stale CUDA graphs, P2P lockstep / timeout reporting)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle.c_oracle import Vof2DCOracle, Vof3DCOracle
from oracle.vof2d_oracle import Vof2DOracle, Vof2DParams
from oracle.vof3d_oracle import Vof3DParams

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CORE = ("F", "u", "v", "p", "kappa", "u_star", "v_star")


def _same(a, b, tag):
    bad = np.argwhere(a != b)
    assert bad.size == 0, f"{tag}: {len(bad)} cells differ, first {bad[0]}: {a[tuple(bad[0])]!r} vs {b[tuple(bad[0])]!r}"


# ----------------------------------------------------------------------------------------------
# BASELINE.json configs 2, 3 and 5 at size: default options (the Jacobi policy and item sizes the bench runs)
# ----------------------------------------------------------------------------------------------
def test_config2_rising_bubble_2048_against_c_oracle(built_lib):
    """-ic 2 at 2048^2, the reference's own constants (Lx = Ly = 0.1, dt = 4e-6), 5 steps, every element."""
    from taichi_2d_vof_b200 import VofSolver2D, reference_params
    P = Vof2DParams(nx=2048, ny=2048)
    o = Vof2DCOracle(P); o.set_init_F(2)
    s = VofSolver2D(reference_params(nx=2048, ny=2048)); s.set_init_F(2)
    _same(s.F.to_numpy(), o.F, "config 2 initial F")
    for step in (1, 2, 5):
        o.run(step - o.istep); s.run(step - s.istep)
        for k in CORE:
            _same(getattr(s, k).to_numpy(), getattr(o, k), f"config 2 (2048^2 -ic 2) step {step} field {k}")
    assert s.diagnostics()["courant_count"] == o.courant_flags


def test_config3_dropping_liquid_8192_against_c_oracle(built_lib):
    """-ic 3 at 8192^2 (constant-dx scaling, the bench workload), 3 steps through the graph-replayed default path
    (T = 5 blocked Jacobi, adaptive kernels), every element of every live field against the C oracle."""
    from taichi_2d_vof_b200 import VofSolver2D, scaled_params
    P = Vof2DParams.scaled(8192)
    o = Vof2DCOracle(P); o.set_init_F(3)
    s = VofSolver2D(scaled_params(8192)); s.set_init_F(3)
    _same(s.F.to_numpy(), o.F, "config 3 initial F")
    out = np.empty((8194, 8194), np.float32)
    for step in (1, 3):
        o.run(step - o.istep)
        while s.istep < step:
            s.step()
        for k in CORE:
            _same(getattr(s, k).to_numpy(out), getattr(o, k), f"config 3 (8192^2 -ic 3) step {step} field {k}")
    m = s.mass()
    assert abs(m - o.mass()) <= 1e-6 * o.mass()


def test_config5_dam_break_3d_256_against_c_oracle(built_lib):
    """3dvof.py at 256^3 (half of config 5's edge: the C oracle needs ~10 s), 3 steps = one rotation of the sweeps."""
    from taichi_2d_vof_b200 import VofSolver3D, scaled_params3d
    P = Vof3DParams.scaled(256)
    o = Vof3DCOracle(P); o.set_init_F(1)
    s = VofSolver3D(scaled_params3d(256)); s.set_init_F(1)
    for step in (1, 3):
        o.run(step - o.istep)
        while s.istep < step:
            s.step()
        for k in ("F", "u", "v", "w", "p"):
            _same(getattr(s, k).to_numpy(), getattr(o, k), f"config 5 (256^3) step {step} field {k}")


# ----------------------------------------------------------------------------------------------
# multi-process slabs (CUDA IPC + device flags across processes): driver-collected
# ----------------------------------------------------------------------------------------------
def _torchrun(script, nproc, port, *args, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script), *args]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, **(env or {})), capture_output=True, text=True, timeout=600)
    return r.returncode, r.stdout + r.stderr


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_two_rank_slabs_equal_single_gpu_2d(built_lib, transport):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    rc, out = _torchrun("mgpu_parity.py", 2, 29611 if transport == "p2p" else 29612, env={"VOF_TRANSPORT": transport})
    assert rc == 0 and "MGPU PARITY OK" in out, out[-3000:]


def test_all_rank_slabs_equal_single_gpu_2d(built_lib):
    n = _ngpu()
    if n < 4:
        pytest.skip("needs >= 4 GPUs")
    rc, out = _torchrun("mgpu_parity.py", n, 29613)
    assert rc == 0 and "MGPU PARITY OK" in out, out[-3000:]


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_two_rank_slabs_equal_single_gpu_3d(built_lib, transport):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    rc, out = _torchrun("mgpu_parity3d.py", 2, 29614 if transport == "p2p" else 29615, "96", "7", env={"VOF_TRANSPORT": transport})
    assert rc == 0 and "MGPU 3D PARITY OK" in out, out[-3000:]


@pytest.mark.parametrize("nslabs", [2, 3])
def test_plane_slabs_p2p_equal_full_domain_3d(built_lib, nslabs):
    """The fused peer-store exchange kernel for the 3-D context, slabs on one device (plain pointers instead of CUDA IPC)."""
    from taichi_2d_vof_b200 import VofSolver3D, reference_params3d
    from taichi_2d_vof_b200.slab import LocalSlabGroup
    nx, ny, nz = 72, 20, 36

    def params_fn(slab, halo, device):
        return reference_params3d(nx=nx, ny=ny, nz=nz, Lx=0.036, Ly=0.01, Lz=0.018, slab=slab, halo=halo, device=device)

    full = VofSolver3D(params_fn(None, 0, 0)); full.set_init_F(1)
    grp = LocalSlabGroup(params_fn, nx, nslabs, p2p=True, solver_cls=VofSolver3D, halo_fields=("F", "u", "v", "w", "p")); grp.set_init_F(1)
    for step in range(9):
        full.step(); grp.step()
    for s in grp.solvers:
        s.p2p_check()
    for k in ("F", "u", "v", "w", "p"):
        _same(grp.gather(k), getattr(full, k).to_numpy(), f"3-D {nslabs} slabs (p2p) field {k}")


# ----------------------------------------------------------------------------------------------
# advisor findings
# ----------------------------------------------------------------------------------------------
def _random_state(P, seed, cfl=0.2):
    rng = np.random.default_rng(seed)
    shape = (P.nx + 2, P.ny + 2)
    vel = cfl * P.dx / P.dt
    F = (rng.random(shape) < 0.5).astype(np.float32)
    band = rng.random(shape) < 0.5
    F[band] = rng.random(int(band.sum())).astype(np.float32)
    u = ((rng.random(shape) * 2 - 1) * vel).astype(np.float32)
    v = ((rng.random(shape) * 2 - 1) * vel).astype(np.float32)
    p = ((rng.random(shape) * 2 - 1) * 100).astype(np.float32)
    return u, v, p, F


@pytest.mark.parametrize("n_jacobi", [10, 2, 0])
def test_slabs_with_minimal_halo_on_a_live_state(built_lib, n_jacobi):
    """Halo depth exactly n_jacobi + 5 (the dependency radius), velocities at CFL ~ 0.2 and a fractional F field, so
    that the limiter and every upwind switch propagate information as far as they can: slabs == full domain."""
    from taichi_2d_vof_b200 import VofError, VofSolver2D, reference_params
    from taichi_2d_vof_b200.slab import LocalSlabGroup, required_halo
    nx, ny = 160, 96
    H = required_halo(n_jacobi)

    def params_fn(slab, halo, device):
        return reference_params(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200, n_jacobi=n_jacobi, slab=slab, halo=halo, device=device)

    with pytest.raises(VofError):
        VofSolver2D(params_fn((1, 80), H - 1, 0))                    # one row less is refused
    P = Vof2DParams(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200, n_jacobi=n_jacobi)
    u, v, p, F = _random_state(P, 7 + n_jacobi)
    full = VofSolver2D(params_fn(None, 0, 0))
    grp = LocalSlabGroup(params_fn, nx, 3, halo=H, n_jacobi=n_jacobi)
    o = Vof2DOracle(P)
    for name, a in (("u", u), ("v", v), ("p", p), ("F", F)):
        getattr(full, name).from_numpy(a)
        getattr(o, name)[...] = a
        for s in grp.solvers:                       # local row l holds global row gi0 + l (rows outside the domain: unused)
            loc = np.zeros((s.nrows, ny + 2), np.float32)
            g0, g1 = max(s.gi0, 0), min(s.gi0 + s.nrows, nx + 2)
            loc[g0 - s.gi0:g1 - s.gi0] = a[g0:g1]
            getattr(s, name).from_numpy(loc)
    for step in range(1, 4):
        full.step(); grp.step(); o.step()
        for k in ("F", "u", "v", "p"):
            _same(getattr(full, k).to_numpy(), getattr(o, k), f"full domain vs oracle, step {step} field {k}")
            _same(grp.gather(k), getattr(o, k), f"3 slabs with halo {H} vs oracle, step {step} field {k}")


@pytest.mark.parametrize("n_jacobi", [10, 2])
def test_streamer_halo_on_a_live_state(built_lib, n_jacobi):
    """The streamed host step (what bench.py's e2e runs) with its default halo on a CFL ~ 0.2 state == the oracle."""
    from taichi_2d_vof_b200 import VofStreamer2D, reference_params
    nx, ny = 240, 128
    P = Vof2DParams(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200, n_jacobi=n_jacobi)
    st = VofStreamer2D(reference_params(nx=nx, ny=ny, Lx=P.Lx, Ly=P.Ly, n_jacobi=n_jacobi), n_slabs=5)
    assert st.halo == n_jacobi + 5
    a = list(_random_state(P, 21 + n_jacobi))
    o = Vof2DOracle(P)
    for name, x in zip(("u", "v", "p", "F"), a):
        getattr(o, name)[...] = x
    for step in range(1, 4):
        o.step()
        st.step_host(*a)
        for name, x in zip(("u", "v", "p", "F"), a):
            _same(x, getattr(o, name), f"streamed step {step} field {name} (n_jacobi {n_jacobi}, halo {st.halo})")
    st.close()


def test_graph_replay_after_an_odd_number_of_buffer_flips(built_lib):
    """vof2d_run's captured graph hard-codes the F / p ping-pong buffers; public single-kernel calls between two runs
    flip them.  The replay must notice (re-capture) instead of reading the stale buffers."""
    from taichi_2d_vof_b200 import VofSolver2D, reference_params
    P = Vof2DParams(nx=96, ny=80, Lx=0.048, Ly=0.04)
    o = Vof2DOracle(P); o.set_init_F(3)
    s = VofSolver2D(reference_params(nx=96, ny=80, Lx=0.048, Ly=0.04)); s.set_init_F(3)
    # materialised rho / nu: the single-kernel entries read the arrays, as the reference's kernels do
    o.run(6); s.run(6, materialize_props=True)
    o.fct_x_sweep(); s.fct_x_sweep()              # F_cur flips once
    o.solve_p_jacobi(); s.solve_p_jacobi()        # p_cur flips once
    o.run(6); s.run(6, materialize_props=True)
    for k in ("F", "u", "v", "p"):
        _same(getattr(s, k).to_numpy(), getattr(o, k), f"graph replay after flips, field {k}")
    o.fct_y_sweep(); s.fct_y_sweep()
    o.run(4); s.run(4, materialize_props=True)
    for k in ("F", "u", "v", "p"):
        _same(getattr(s, k).to_numpy(), getattr(o, k), f"second replay after a flip, field {k}")


def test_p2p_check_reports_lockstep_violation(built_lib):
    """A slab whose neighbour holds F in the other ping-pong buffer must be reported, not silently served stale rows."""
    from taichi_2d_vof_b200 import VofError, reference_params
    from taichi_2d_vof_b200.slab import LocalSlabGroup
    nx, ny = 128, 64

    def params_fn(slab, halo, device):
        return reference_params(nx=nx, ny=ny, Lx=0.064, Ly=0.032, slab=slab, halo=halo, device=device)

    grp = LocalSlabGroup(params_fn, nx, 2, p2p=True); grp.set_init_F(3)
    for _ in range(3):
        grp.step()
    for s in grp.solvers:
        s.p2p_check()                                  # in lockstep: fine
    grp.solvers[0].fct_x_sweep()                       # rank 0 flips its F buffer, rank 1 does not
    grp.exchange_halos()
    with pytest.raises(VofError, match="lockstep"):
        for s in grp.solvers:
            s.p2p_check()


def test_create_failure_does_not_leak(built_lib):
    """A context that fails half-way through creation (arena too small) is torn down through the one cleanup path."""
    import ctypes as C
    import torch
    from taichi_2d_vof_b200 import reference_params
    free0, _ = torch.cuda.mem_get_info()
    P = reference_params(nx=512, ny=512)
    need = built_lib.vof2d_arena_bytes(C.byref(P))
    buf = torch.empty(need // 2, dtype=torch.uint8, device="cuda")
    h = C.c_void_p()
    for _ in range(20):
        assert built_lib.vof2d_create_in(C.byref(P), C.c_void_p(buf.data_ptr()), need // 2, C.byref(h)) == -1
        assert b"arena too small" in built_lib.vof_last_error()
    del buf
    torch.cuda.empty_cache()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20


# ----------------------------------------------------------------------------------------------
# the on-disk output of `2dvof.py -s` / `3dvof.py`, through the drop-in scripts at the repo root
# ----------------------------------------------------------------------------------------------
def test_2dvof_script_dash_s_writes_the_reference_steps(built_lib, tmp_path):
    """`python 2dvof.py -ic 3 -s` (the reference's default grid, 200 steps, headless): the F arrays it writes every
    100 steps (output/%06d-f.npy, 2dvof.py:563-571) equal the oracle's F at those steps, every element."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "2dvof.py"), "-ic", "3", "-s", "--steps", "200"], cwd=tmp_path,
                       env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ">>> Grid resolution: 200 x 200, dt = 4.00e-06" in r.stdout
    o = Vof2DCOracle(Vof2DParams()); o.set_init_F(3)
    for count in (0, 1):
        o.run(100)
        F = np.load(tmp_path / "output" / f"{count:06d}-f.npy")
        _same(F, o.F, f"2dvof.py -s, file {count:06d}-f.npy")


def test_3dvof_script_writes_vtr_with_the_oracle_field(built_lib, tmp_path):
    """`python 3dvof.py` on a small grid, 100 steps: output/step-00100.vtr (3dvof.py:624-627) holds the oracle's F."""
    from taichi_2d_vof_b200.vtk import read_vtr_arrays
    r = subprocess.run([sys.executable, os.path.join(ROOT, "3dvof.py"), "--nx", "24", "--ny", "20", "--nz", "28", "--scaled",
                        "--steps", "100"], cwd=tmp_path, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ">>> Exporting step-00100 result..." in r.stdout
    P = Vof3DParams(nx=24, ny=20, nz=28, Lx=0.012, Ly=0.01, Lz=0.014)
    o = Vof3DCOracle(P); o.set_init_F(1); o.run(100)
    back = read_vtr_arrays(str(tmp_path / "output" / "step-00100.vtr"))
    assert back["shape"] == (26, 22, 30)
    _same(back["VOF"], o.F, "3dvof.py export, step-00100.vtr point data VOF")


def test_async_field_read_does_not_block_and_is_a_snapshot(built_lib):
    """vof2d_field_get_async: returns at once, the loop continues on the compute stream, and the array that arrives is the
    field at the time of the call (not a later state)."""
    import time
    from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params
    s = VofSolver2D(scaled_params(4096)); s.set_init_F(3); s.run(10)
    want = s.F.to_numpy()
    out = _lib.pinned_empty(want.shape)
    s.F.to_numpy_async(out); s.F.wait()              # the first call sets up the side stream and the device snapshot buffer
    s.synchronize()
    t0 = time.perf_counter()
    s.F.to_numpy_async(out)
    t_call = time.perf_counter() - t0
    s.run(40)                                        # 40 more steps are queued behind the snapshot, none behind the copy
    s.F.wait()
    assert np.array_equal(out, want)
    assert t_call < 0.02, f"the async read took {t_call * 1e3:.1f} ms to return"
    assert not np.array_equal(s.F.to_numpy(), want)
    _lib.pinned_free(out)


def test_driver_gpus_flag_equals_single_gpu(built_lib, tmp_path):
    """`2dvof.py --gpus 2 --dump` == `2dvof.py --dump` (row slabs + NVLink peer stores behind the drop-in CLI)."""
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    common = ["-ic", "3", "--nx", "512", "--ny", "384", "--scaled", "--steps", "30", "-s", "--nstep", "10"]
    for tag, extra in (("one", []), ("two", ["--gpus", "2"])):
        d = tmp_path / tag
        d.mkdir()
        r = subprocess.run([sys.executable, os.path.join(ROOT, "2dvof.py"), *common, *extra, "--dump", "state.npz"], cwd=d,
                           env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
    a, b = np.load(tmp_path / "one" / "state.npz"), np.load(tmp_path / "two" / "state.npz")
    for k in ("u", "v", "p", "F"):
        _same(b[k], a[k], f"--gpus 2 vs one GPU, field {k}")
    for count in range(3):
        _same(np.load(tmp_path / "two" / "output" / f"{count:06d}-f.npy"), np.load(tmp_path / "one" / "output" / f"{count:06d}-f.npy"),
              f"--gpus 2 -s output {count}")


def test_chebyshev_pressure_solver_is_the_accelerated_jacobi_iteration(built_lib):
    """SURVEY 8(f) rank 4, opt-in (VOF_OPT_PRESSURE_SOLVER = 1): n sweeps of the Chebyshev semi-iterative acceleration of
    the reference's Jacobi iteration.  (1) It is that recurrence: x(k+1) = x(k-1) + w(k+1) (J x(k) - x(k-1)) with J the
    oracle's sweep (2dvof.py:236-266) -- compared with a NumPy evaluation; (2) it is stronger: on a smooth right-hand
    side (what a flow produces) 200 sweeps leave a Poisson residual more than 10 times below plain Jacobi's, whose
    smooth error components decay like 0.9997^k; (3) the default stays the reference's iteration."""
    import math
    from taichi_2d_vof_b200 import VofSolver2D, _lib
    from oracle.vof2d_oracle import Vof2DOracle, Vof2DParams
    nx, ny, n = 64, 96, 200
    P = Vof2DParams(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200)
    rng = np.random.default_rng(5)
    o = Vof2DOracle(P)
    shp = o.F.shape
    o.rho[...] = 50 + 950 * rng.random(shp, dtype=np.float32)
    ii, jj = np.arange(shp[0])[:, None] / nx, np.arange(shp[1])[None, :] / ny
    o.u_star[...] = (1e-3 * np.sin(math.pi * ii) * np.cos(math.pi * jj)).astype(np.float32)
    o.v_star[...] = (1e-3 * np.cos(2 * math.pi * ii) * np.sin(math.pi * jj)).astype(np.float32)
    o.u_star[1, :] = 0; o.u_star[nx + 1, :] = 0; o.v_star[:, 1] = 0; o.v_star[:, ny + 1] = 0    # compatible rhs
    o.p[...] = 0

    def solver(mode):
        from taichi_2d_vof_b200 import reference_params
        s = VofSolver2D(reference_params(nx=nx, ny=ny, Lx=P.Lx, Ly=P.Ly))
        s.set_option(_lib.VOF_OPT_PRESSURE_SOLVER, mode)
        for k in ("rho", "u_star", "v_star", "p"):
            getattr(s, k).from_numpy(getattr(o, k))
        s.solve_p_jacobi(n)
        return s.p.to_numpy()

    p_jac, p_cheb = solver(0), solver(1)
    # (3) default = the reference's sweeps
    ref = Vof2DOracle(P)
    for k in ("rho", "u_star", "v_star", "p"):
        getattr(ref, k)[...] = getattr(o, k)
    for _ in range(n):
        ref.solve_p_jacobi()
    assert np.array_equal(p_jac, ref.p)
    # (1) the recurrence, evaluated with the oracle's sweep as J
    rho_s = 0.5 * (1.0 + math.cos(math.pi / max(nx, ny)))
    w = 1.0
    ch = Vof2DOracle(P)
    for k in ("rho", "u_star", "v_star", "p"):
        getattr(ch, k)[...] = getattr(o, k)
    prev = ch.p.copy()
    for k in range(1, n + 1):
        cur = ch.p.copy()
        ch.solve_p_jacobi()                                   # ch.p = J cur
        if k > 1:
            w = 1.0 / (1.0 - 0.5 * rho_s ** 2) if k == 2 else 1.0 / (1.0 - 0.25 * rho_s ** 2 * w)
            ch.p[1:-1, 1:-1] = prev[1:-1, 1:-1] + np.float32(w) * (ch.p[1:-1, 1:-1] - prev[1:-1, 1:-1])
        prev = cur
    scale = np.abs(ch.p).max()
    assert np.abs(p_cheb - ch.p).max() <= 1e-5 * scale

    # (2) residual of the Poisson equation b - A p (interior, fp64), Chebyshev vs Jacobi
    def residual(p):
        b = ref.poisson_rhs().astype(np.float64)
        c = float(ref.c_dxi2)
        pp = p.astype(np.float64)
        i = np.arange(1, nx + 1)[:, None]; j = np.arange(1, ny + 1)[None, :]
        ae = np.where(i != nx, c, 0.0); aw = np.where(i != 1, c, 0.0)
        an = np.where(j != ny, float(ref.c_dyi2), 0.0); a_s = np.where(j != 1, float(ref.c_dyi2), 0.0)
        ap = -(ae + aw + an + a_s)
        r = b - ae * pp[2:, 1:-1] - aw * pp[:-2, 1:-1] - an * pp[1:-1, 2:] - a_s * pp[1:-1, :-2] - ap * pp[1:-1, 1:-1]
        return float(np.sqrt(np.mean(r * r)))
    assert residual(p_cheb) < 0.1 * residual(p_jac), (residual(p_cheb), residual(p_jac))


def test_tolerance_mode_of_the_pressure_sweeps_meets_the_north_star_tolerances(built_lib):
    """VOF_OPT_FAST_MATH = 1 (opt-in): the blocked Jacobi with fused multiply-adds and a reciprocal multiply -- 5 instead
    of 8 operations per cell-update, not bit-exact.  Against the exact path on a 2048 x 2048 dropping-liquid run (large
    enough for the blocked kernel by default): rel. L-inf (max|a - b| / max|b|) <= 1e-5 after one step and
    <= 1e-3 after 100 steps in u, v, p, F; volume within 1e-6.  The default stays bit-exact (every other test)."""
    from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params
    n = 2048

    def run(fast):
        s = VofSolver2D(scaled_params(n))
        s.set_option(_lib.VOF_OPT_FAST_MATH, fast)
        s.set_init_F(3)
        s.step()
        one = {k: getattr(s, k).to_numpy() for k in ("u", "v", "p", "F")}
        s.run(99)
        hundred = {k: getattr(s, k).to_numpy() for k in ("u", "v", "p", "F")}
        m = s.mass()
        s.close()
        return one, hundred, m

    e1, e100, em = run(0)
    f1, f100, fm = run(1)
    assert any(np.any(e1[k] != f1[k]) for k in e1), "the tolerance mode did not change a bit: it is not active"
    for tag, e, f, tol in (("1 step", e1, f1, 1e-5), ("100 steps", e100, f100, 1e-3)):
        for k in e:
            scale = float(np.abs(e[k]).max())
            if scale == 0.0:
                assert not np.any(f[k])
                continue
            rel = float(np.abs(e[k].astype(np.float64) - f[k]).max()) / scale
            assert rel <= tol, f"{tag}: {k} rel. L-inf {rel:.3e} > {tol}"
    assert abs(em - fm) <= 1e-6 * em
