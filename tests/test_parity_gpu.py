"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle.

Arithmetic in libvof is order-exact fp32 (no FMA contraction, IEEE / and sqrt), so the bar here
is stronger than the north-star tolerances: fields must be IDENTICAL to the oracle's (== on
every element, -0.0 == +0.0).  The north-star tolerances (relative L-inf <= 1e-5 after one
step, <= 1e-3 after 100 steps, volume to 1e-6) are asserted as well and are the contractual gate.
"""
import numpy as np
import pytest

from oracle.c_oracle import Vof2DCOracle
from oracle.vof2d_oracle import Vof2DOracle, Vof2DParams, rel_linf

pytestmark = pytest.mark.gpu

TOL_1STEP = 1e-5     # north_star: per-field relative L-inf after one fp32 timestep
TOL_100STEP = 1e-3   # north_star: after 100 steps
TOL_VOLUME = 1e-6    # north_star: total VOF volume


def _solver(P, **kw):
    from taichi_2d_vof_b200 import VofSolver2D, reference_params
    return VofSolver2D(reference_params(nx=P.nx, ny=P.ny, Lx=P.Lx, Ly=P.Ly, n_jacobi=P.n_jacobi, **kw))


def _params(P, **kw):
    from taichi_2d_vof_b200 import reference_params
    return reference_params(nx=P.nx, ny=P.ny, Lx=P.Lx, Ly=P.Ly, n_jacobi=P.n_jacobi, **kw)


def _compare(s, o, fields, tol, exact=True, tag=""):
    for k in fields:
        a, b = getattr(s, k).to_numpy(), getattr(o, k)
        err = rel_linf(a, b)
        assert err <= tol, f"{tag} field {k}: rel L-inf {err:.3e} > {tol}"
        if exact:
            bad = np.argwhere(a != b)
            assert bad.size == 0, f"{tag} field {k}: {len(bad)} elements differ, first at {bad[0]}: {a[tuple(bad[0])]} vs {b[tuple(bad[0])]}"


ALL = ("F", "u", "v", "p", "rho", "nu", "kappa", "u_star", "v_star")
CORE = ("F", "u", "v", "p", "kappa", "u_star", "v_star")


@pytest.mark.parametrize("ic", [1, 2, 3])
def test_init_matches_oracle(built_lib, ic):
    P = Vof2DParams()
    o = Vof2DOracle(P); o.set_init_F(ic)
    s = _solver(P); s.set_init_F(ic)
    assert np.array_equal(s.F.to_numpy(), o.F)


@pytest.mark.parametrize("ic", [1, 2, 3])
@pytest.mark.parametrize("shape", [(200, 200), (64, 96), (37, 41), (130, 260)])
def test_each_kernel_in_sequence(built_lib, ic, shape):
    """Call the C-ABI entries one reference kernel at a time and compare after EVERY call."""
    nx, ny = shape
    P = Vof2DParams(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200)
    o = Vof2DOracle(P); o.set_init_F(ic)
    s = _solver(P); s.set_init_F(ic)
    for step in range(1, 6):
        o.istep += 1; s.istep += 1
        for name in ("cal_nu_rho", "get_normal_young", "advect_upwind", "set_BC"):
            getattr(o, name)(); getattr(s, name)()
            _compare(s, o, ALL, TOL_1STEP, tag=f"step {step} after {name}")
        for sweep in range(P.n_jacobi):
            o.solve_p_jacobi(); s.solve_p_jacobi()
        _compare(s, o, ALL, TOL_1STEP, tag=f"step {step} after jacobi")
        for name in ("update_uv", "set_BC", "solve_VOF_rudman", "post_process_f", "set_BC"):
            getattr(o, name)(); getattr(s, name)()
            _compare(s, o, ALL, TOL_1STEP, tag=f"step {step} after {name}")


@pytest.mark.parametrize("ic", [1, 2, 3])
def test_one_step_fused(built_lib, ic):
    P = Vof2DParams()
    o = Vof2DOracle(P); o.set_init_F(ic); o.step()
    s = _solver(P); s.set_init_F(ic); s.step(materialize_props=True)
    _compare(s, o, ALL, TOL_1STEP, tag="fused step 1")
    s2 = _solver(P); s2.set_init_F(ic); s2.step()
    _compare(s2, o, CORE, TOL_1STEP, tag="fused step 1 (props inline)")


@pytest.mark.parametrize("ic", [1, 2, 3])
@pytest.mark.parametrize("mode", ["fused", "sequence", "no_fusion", "graph"])
def test_100_steps(built_lib, ic, mode):
    P = Vof2DParams()
    o = Vof2DCOracle(P); o.set_init_F(ic)
    m0 = o.mass()
    o.run(100)
    s = _solver(P); s.set_init_F(ic)
    if mode == "fused":
        for _ in range(100):
            s.step()
    elif mode == "sequence":
        for _ in range(100):
            s.step_sequence()
    elif mode == "no_fusion":
        for _ in range(100):
            s.step(no_fusion=True)
    else:
        s.run(100)
    _compare(s, o, CORE, TOL_100STEP, tag=f"100 steps ({mode})")
    d = s.diagnostics()
    assert abs(d["mass"] - o.mass()) / o.mass() <= TOL_VOLUME
    assert d["courant_count"] == o.courant_flags
    # the scheme's own drift (clamps + 2^-23 quantisation), reported not gated at 1e-6 by the oracle either
    assert abs(d["mass"] - m0) / m0 < 1e-4


def test_step_host_roundtrip(built_lib):
    P = Vof2DParams(nx=96, ny=80, Lx=0.048, Ly=0.04)
    o = Vof2DOracle(P); o.set_init_F(3)
    s = _solver(P)
    u, v, p, F = (getattr(o, k).copy() for k in ("u", "v", "p", "F"))
    for _ in range(4):
        o.step()
        s.step_host(u, v, p, F)
    for a, k in ((u, "u"), (v, "v"), (p, "p"), (F, "F")):
        assert np.array_equal(a, getattr(o, k)), k


@pytest.mark.parametrize("nx,ny,n_slabs,inplace", [(96, 80, 3, True), (96, 80, 6, False), (200, 200, 4, True),
                                                    (333, 130, 5, True), (64, 64, 1, True)])
def test_streamed_step_host(built_lib, nx, ny, n_slabs, inplace):
    """vof2d_streamer_step_host (row slabs streamed through the GPU, host-resident state) == oracle, bit for bit."""
    from taichi_2d_vof_b200 import VofStreamer2D
    P = Vof2DParams(nx=nx, ny=ny, Lx=0.0005 * nx, Ly=0.0005 * ny)
    o = Vof2DOracle(P); o.set_init_F(3)
    st = VofStreamer2D(_params(P), n_slabs=n_slabs)
    assert st.n_slabs == n_slabs and (n_slabs == 1 or st.halo >= P.n_jacobi + 5)
    a = [getattr(o, k).copy() for k in ("u", "v", "p", "F")]
    b = [np.empty_like(x) for x in a]
    for step in range(6):
        o.step()
        if inplace:
            st.step_host(*a)
        else:
            st.step_host(*a, out=b)
            a, b = b, a
        for x, k in zip(a, ("u", "v", "p", "F")):
            assert np.array_equal(x, getattr(o, k)), (k, step)
    st.close()


def test_streamed_step_host_random_state(built_lib):
    """Streamed vs whole-array host step on random fields (every slab boundary carries non-trivial data)."""
    from taichi_2d_vof_b200 import VofStreamer2D
    rng = np.random.default_rng(5)
    P = Vof2DParams(nx=260, ny=150, Lx=0.13, Ly=0.075)
    shape = (P.nx + 2, P.ny + 2)
    a = [(rng.random(shape, dtype=np.float32) - 0.5) * sc for sc in (2.0, 2.0, 100.0)] + [rng.random(shape, dtype=np.float32)]
    b = [x.copy() for x in a]
    s = _solver(P)
    st = VofStreamer2D(_params(P), n_slabs=7)
    for step in range(3):
        s.step_host(*a)
        st.step_host(*b)
        for x, y, k in zip(a, b, ("u", "v", "p", "F")):
            assert np.array_equal(x, y), (k, step)


def test_streamer_argument_errors(built_lib):
    from taichi_2d_vof_b200 import VofError, VofStreamer2D
    P = Vof2DParams(nx=64, ny=64)
    with pytest.raises(VofError):
        VofStreamer2D(_params(P), n_slabs=8)        # 8-row slabs are thinner than the halo
    st = VofStreamer2D(_params(P), n_slabs=2)
    bad = np.zeros((10, 10), np.float32)
    with pytest.raises(ValueError):
        st.step_host(bad, bad, bad, bad)


def test_random_state_one_step(built_lib):
    """Synthetic fields: every branch of the upwind / FCT switches gets exercised."""
    rng = np.random.default_rng(0)
    P = Vof2DParams(nx=150, ny=170, Lx=0.075, Ly=0.085)
    o = Vof2DOracle(P)
    shape = o.F.shape
    o.F[...] = rng.random(shape, dtype=np.float32)
    o.u[...] = (rng.random(shape, dtype=np.float32) - 0.5) * 2.0
    o.v[...] = (rng.random(shape, dtype=np.float32) - 0.5) * 2.0
    o.p[...] = (rng.random(shape, dtype=np.float32) - 0.5) * 100.0
    s = _solver(P)
    for k in ("F", "u", "v", "p"):
        getattr(s, k).from_numpy(getattr(o, k))
    for step in range(3):
        o.step(); s.step(materialize_props=True)
        _compare(s, o, ALL, TOL_1STEP, tag=f"random step {step + 1}")


def test_diagnostics_and_torch_view(built_lib):
    import torch
    P = Vof2DParams()
    s = _solver(P); s.set_init_F(2)
    for _ in range(3):
        s.step()
    F = s.F.to_numpy()
    d = s.diagnostics()
    assert abs(d["mass"] - float(F[1:-1, 1:-1].sum(dtype=np.float64))) < 1e-6 * d["mass"]
    t = s.F.torch()
    assert t.shape == F.shape and t.is_cuda
    assert np.array_equal(t.cpu().numpy(), F)
    u, v = s.u.to_numpy(), s.v.to_numpy()
    cfl = max(np.abs(u[1:-1, 1:-1]).max() * P.dt / P.dx, np.abs(v[1:-1, 1:-1]).max() * P.dt / P.dy)
    assert abs(d["max_cfl"] - cfl) <= 1e-5 * cfl
    assert d["residual"] >= 0


@pytest.mark.parametrize("pk", [1, 0])
@pytest.mark.parametrize("maxt", [5, 4])
@pytest.mark.parametrize("nsweeps", [1, 2, 3, 4, 5, 6, 7, 10, 11])
@pytest.mark.parametrize("shape", [(70, 300), (400, 130), (97, 131)])
def test_jacobi_temporal_blocking_equals_single_sweeps(built_lib, nsweeps, shape, maxt, pk):
    """vof2d_solve_p_jacobi(n) (<= 5 sweeps per HBM pass, register-pipelined; pk = 1: the packed fp32x2 kernel of the
    third generation, 0: the second) must equal n calls of the reference's single sweep exactly -- interior, walls
    (zeroed coefficients) and the untouched ghost frame (random, both signs: it enters as 0 * p[ghost])."""
    rng = np.random.default_rng(nsweeps)
    nx, ny = shape
    P = Vof2DParams(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200)
    o = Vof2DOracle(P)
    shp = o.F.shape
    o.rho[...] = 50 + 950 * rng.random(shp, dtype=np.float32)
    o.u_star[...] = (rng.random(shp, dtype=np.float32) - 0.5) * 1e-3
    o.v_star[...] = (rng.random(shp, dtype=np.float32) - 0.5) * 1e-3
    o.p[...] = (rng.random(shp, dtype=np.float32) - 0.5) * 100.0
    o.p[5:9, 7:11] = 0.0
    o.p[20, 20] = 1e-36          # tiny numerators take the IEEE path of the reciprocal division
    from taichi_2d_vof_b200 import _lib
    s = _solver(P)
    s.set_option(_lib.VOF_OPT_JACOBI_TB, 2)      # force the blocked kernel (small grids default to single sweeps)
    s.set_option(_lib.VOF_OPT_JACOBI_MAXT, maxt)  # 5: 10 = 5 + 5 sweeps per pass; 4: 4 + 3 + 3, narrower strip margins
    s.set_option(_lib.VOF_OPT_JACOBI_PK, pk)
    o.p[1, 1:4] = [3e-37, -2e-38, 1e-44]          # tiny numerators in a wall row / corner, subnormal included
    o.p[nx, ny - 2:ny + 1] = [-1e-40, 2e-39, -3e-36]
    for k in ("rho", "u_star", "v_star", "p"):
        getattr(s, k).from_numpy(getattr(o, k))
    for _ in range(nsweeps):
        o.solve_p_jacobi()
    s.solve_p_jacobi(nsweeps)
    a, b = s.p.to_numpy(), o.p
    bad = np.argwhere(a != b)
    assert bad.size == 0, f"{len(bad)} cells differ, first {bad[0]}: {a[tuple(bad[0])]} vs {b[tuple(bad[0])]}"


@pytest.mark.parametrize("long_pct,rows", [(75, 0), (0, 17), (100, 24), (50, 9)])
def test_packed_jacobi_item_sizes_identical(built_lib, long_pct, rows):
    """The packed kernel's work items (edge strips, wall rows, long and short chunks of the bulk) tile the grid for
    every item size: 1700 x 520 (5 strips, two of them edge strips), 10 sweeps, against the second generation."""
    from taichi_2d_vof_b200 import _lib
    rng = np.random.default_rng(7)
    P = Vof2DParams(nx=1700, ny=520, Lx=0.85, Ly=0.26)
    outs = []
    for pk in (0, 1):
        s = _solver(P)
        s.set_option(_lib.VOF_OPT_JACOBI_TB, 2)
        s.set_option(_lib.VOF_OPT_JACOBI_PK, pk)
        if pk:
            s.set_option(_lib.VOF_OPT_JACOBI_LONG_PCT, long_pct)
            s.set_option(_lib.VOF_OPT_JACOBI_ROWS, rows)
        shp = (P.nx + 2, P.ny + 2)
        rng2 = np.random.default_rng(11)
        s.rho.from_numpy(50 + 950 * rng2.random(shp, dtype=np.float32))
        s.u_star.from_numpy((rng2.random(shp, dtype=np.float32) - 0.5) * 1e-3)
        s.v_star.from_numpy((rng2.random(shp, dtype=np.float32) - 0.5) * 1e-3)
        s.p.from_numpy((rng2.random(shp, dtype=np.float32) - 0.5) * 100.0)
        s.solve_p_jacobi(10)
        outs.append(s.p.to_numpy())
    bad = np.argwhere(outs[0] != outs[1])
    assert bad.size == 0, f"{len(bad)} cells differ, first {bad[0]}"


def test_jacobi_tb_on_off_identical(built_lib):
    from taichi_2d_vof_b200 import _lib
    P = Vof2DParams(nx=256, ny=384, Lx=0.128, Ly=0.192)
    outs = []
    for tb in (2, 0):
        s = _solver(P); s.set_option(_lib.VOF_OPT_JACOBI_TB, tb); s.set_init_F(3)
        for _ in range(6):
            s.step()
        outs.append(s.state())
    for k in CORE:
        assert np.array_equal(outs[0][k], outs[1][k]), k


@pytest.mark.parametrize("ic", [2, 3])
def test_larger_grid_against_c_oracle(built_lib, ic):
    """1024 x 768 (many strips / chunks / tiles), 12 steps, fused path vs the C twin of the oracle."""
    P = Vof2DParams(nx=1024, ny=768, Lx=0.512, Ly=0.384)
    from taichi_2d_vof_b200 import _lib
    o = Vof2DCOracle(P); o.set_init_F(ic); o.run(12)
    s = _solver(P)
    if ic == 3:
        s.set_option(_lib.VOF_OPT_JACOBI_TB, 2)     # the temporally blocked path (default only beyond the L2 size)
    s.set_init_F(ic); s.run(12)
    _compare(s, o, CORE, TOL_1STEP, tag="1024x768, 12 steps")


@pytest.mark.parametrize("nslabs", [2, 3])
def test_row_slabs_p2p_flags_equal_full_domain(built_lib, nslabs):
    """The NVLink-P2P exchange path (kernel stores into the neighbour's arena + device-side flag hand-shake, no host
    synchronisation) on one device: slabs of unequal height, 30 steps, identical to the single-domain run."""
    from taichi_2d_vof_b200 import VofSolver2D, reference_params
    from taichi_2d_vof_b200.slab import LocalSlabGroup
    nx, ny = 200, 64

    def params_fn(slab, halo, device):
        return reference_params(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200, slab=slab, halo=halo, device=device)

    full = VofSolver2D(params_fn(None, 0, 0)); full.set_init_F(3)
    grp = LocalSlabGroup(params_fn, nx, nslabs, p2p=True); grp.set_init_F(3)
    for step in range(30):
        full.step(); grp.step()
    for s in grp.solvers:
        assert s.p2p_status() == 0
    for k in ("F", "u", "v", "p"):
        assert np.array_equal(grp.gather(k), getattr(full, k).to_numpy()), k


@pytest.mark.parametrize("nslabs", [2, 3])
@pytest.mark.parametrize("ic", [1, 3])
def test_row_slabs_equal_full_domain(built_lib, nslabs, ic):
    """Row-slab decomposition with deep halos (one exchange per step) must reproduce the single-domain
    run bit for bit: slab contexts on one device, halos moved with vof2d_halo_push."""
    from taichi_2d_vof_b200 import VofSolver2D, reference_params
    from taichi_2d_vof_b200.slab import LocalSlabGroup
    nx, ny = 192, 96

    def params_fn(slab, halo, device):
        return reference_params(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200, slab=slab, halo=halo, device=device)

    full = VofSolver2D(params_fn(None, 0, 0)); full.set_init_F(ic)
    grp = LocalSlabGroup(params_fn, nx, nslabs); grp.set_init_F(ic)
    for step in range(1, 25):
        full.step(); grp.step()
        if step in (1, 2, 5, 24):
            for k in ("F", "u", "v", "p"):
                a, b = grp.gather(k), getattr(full, k).to_numpy()
                bad = np.argwhere(a != b)
                assert bad.size == 0, f"step {step} field {k}: {len(bad)} cells differ, first {bad[0]}"
    # owned-row diagnostics add up to the global ones
    m = sum(s.diagnostics(residual=False)["mass"] for s in grp.solvers)
    assert abs(m - full.mass()) <= 1e-9 * full.mass()


import glob as _glob
import os as _os

_GOLD = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(_glob.glob(_os.path.join(_GOLD, "vof2d_ic*_*.npz"))))
def test_cuda_reproduces_committed_golden_vectors(built_lib, path):
    """The CUDA path against the committed fixtures (tests/golden/make_golden.py), without the oracle in the loop."""
    from taichi_2d_vof_b200 import VofSolver2D, reference_params
    g = np.load(path)
    nx, ny, Lx, Ly = int(g["params"][0]), int(g["params"][1]), float(g["params"][2]), float(g["params"][3])
    ic = int(_os.path.basename(path)[8])
    s = VofSolver2D(reference_params(nx=nx, ny=ny, Lx=Lx, Ly=Ly))
    s.set_init_F(ic)
    assert np.array_equal(s.F.to_numpy(), g["F_init"])
    for ck in (1, 10, 100):
        s.run(ck - s.istep)
        for k in ("u", "v", "p", "F", "kappa"):
            a, b = getattr(s, k).to_numpy(), g[f"{k}_{ck}"]
            assert rel_linf(a, b) <= (TOL_1STEP if ck == 1 else TOL_100STEP)
            assert np.array_equal(a, b), f"{k} after {ck} steps"
        assert abs(s.mass() - float(g[f"mass_{ck}"])) <= TOL_VOLUME * float(g[f"mass_{ck}"])


@pytest.mark.parametrize("shape", [(4, 4), (5, 7), (8, 129), (131, 6), (33, 1023)])
def test_edge_shapes(built_lib, shape):
    """Minimum sizes, odd widths (partial float4 / float2 lanes), strips of one lane, very flat and very tall grids."""
    nx, ny = shape
    rng = np.random.default_rng(nx * 1000 + ny)
    P = Vof2DParams(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200)
    o = Vof2DOracle(P)
    shp = o.F.shape
    o.F[...] = rng.random(shp, dtype=np.float32)
    o.u[...] = (rng.random(shp, dtype=np.float32) - 0.5) * 0.5
    o.v[...] = (rng.random(shp, dtype=np.float32) - 0.5) * 0.5
    o.p[...] = (rng.random(shp, dtype=np.float32) - 0.5) * 10.0
    from taichi_2d_vof_b200 import _lib
    for tb in (0, 2):
        s = _solver(P); s.set_option(_lib.VOF_OPT_JACOBI_TB, tb)
        for k in ("F", "u", "v", "p"):
            getattr(s, k).from_numpy(getattr(o, k))
        o2 = Vof2DOracle(P)
        for k in ("F", "u", "v", "p"):
            getattr(o2, k)[...] = getattr(o, k)
        for step in range(3):
            o2.step(); s.step(materialize_props=True)
            _compare(s, o2, ALL, TOL_1STEP, tag=f"{shape} tb={tb} step {step + 1}")


def test_full_size_properties_8192(built_lib):
    """At the benchmark size the oracle is too slow for a full comparison; check size-independent properties:
    bounds, finiteness, volume drift, x<->y transposition symmetry of the FCT sweeps, blocked == un-blocked Jacobi."""
    from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params
    n = 8192
    s = VofSolver2D(scaled_params(n)); s.set_init_F(3)
    m0 = s.mass()
    for _ in range(4):
        s.step()
    F = s.F.to_numpy()
    assert F.min() >= 0.0 and F.max() <= 1.0 and np.isfinite(F).all()
    d = s.diagnostics()
    assert np.isfinite(d["max_cfl"]) and d["courant_count"] == 0
    assert abs(d["mass"] - m0) / m0 < 1e-6
    # blocked (5 + 5 sweeps per HBM pass) vs one sweep per launch, same state
    p0, us, vs = s.p.to_numpy(), s.u_star.to_numpy(), s.v_star.to_numpy()
    s.cal_nu_rho()
    s.set_option(_lib.VOF_OPT_JACOBI_TB, 2); s.solve_p_jacobi(10); pa = s.p.to_numpy()
    s.p.from_numpy(p0)
    s.set_option(_lib.VOF_OPT_JACOBI_TB, 0); s.solve_p_jacobi(10); pb = s.p.to_numpy()
    assert np.array_equal(pa, pb)
    del p0, us, vs, pa, pb
    # fct_x on (F, u) == transpose of fct_y on (F^T, v = u^T)   (dx == dy)
    u = s.u.to_numpy()
    t = VofSolver2D(scaled_params(n))
    t.F.from_numpy(np.ascontiguousarray(F.T)); t.v.from_numpy(np.ascontiguousarray(u.T))
    s.fct_x_sweep(); t.fct_y_sweep()
    assert np.array_equal(s.F.to_numpy(), t.F.to_numpy().T)


def test_display_kernels_match_oracle(built_lib):
    """get_vof_field / get_u_field / get_v_field / get_vnorm_field / interp_velocity (2dvof.py:458-492)."""
    P = Vof2DParams(nx=96, ny=130, Lx=0.048, Ly=0.065)
    o = Vof2DOracle(P); o.set_init_F(3)
    s = _solver(P); s.set_init_F(3)
    for _ in range(30):
        o.step(); s.step()
    for name in ("get_vof_field", "get_u_field", "get_v_field", "get_vnorm_field", "interp_velocity"):
        a, b = getattr(s, name)(), getattr(o, name)()
        assert a.shape == b.shape and a.dtype == b.dtype
        assert np.array_equal(a, b), f"{name}: {(a != b).sum()} elements differ"
    assert s.get_vof_field().shape == (2 * P.nx, 2 * P.ny)


def test_display_kernels_refuse_slabs(built_lib):
    from taichi_2d_vof_b200 import VofSolver2D, reference_params
    s = VofSolver2D(reference_params(nx=64, ny=64, slab=(1, 32), halo=16))
    with pytest.raises(Exception):
        s.get_vof_field()


@pytest.mark.parametrize("n,adaptive", [(192, 1), (192, 0), (130, 1)])
def test_kothe_rider_fct_matches_oracle(built_lib, n, adaptive):
    """Pure FCT transport in the single-vortex field of the reference's test/forward_fct.py (Kothe-Rider), through the
    production sweeps: the disc is stretched into a filament, so bulk strips, interface strips and every transition
    between the adaptive kernels' modes occur in both sweep directions.  F identical to the oracle after every step."""
    from test_oracle_cpu import kothe_rider_state
    from taichi_2d_vof_b200 import _lib
    P, o = kothe_rider_state(n)
    s = _solver(P); s.set_option(_lib.VOF_OPT_ADAPTIVE, adaptive)
    for k in ("F", "u", "v"):
        getattr(s, k).from_numpy(getattr(o, k))
    for step in range(60):
        o.istep += 1; s.istep += 1
        for name in ("solve_VOF_rudman", "post_process_f", "set_BC"):
            getattr(o, name)(); getattr(s, name)()
        a = s.F.to_numpy()
        assert np.array_equal(a, o.F), f"step {step + 1}: F differs in {(a != o.F).sum()} cells"


def test_checkpoint_restart_is_exact(built_lib, tmp_path):
    """u, v, p, F and istep are the whole state (rho, nu, kappa, u*, v* are recomputed every step), so
    a run continued from a dump equals the uninterrupted run bit for bit -- 2-D through the driver's --dump / --resume,
    3-D through the solver."""
    import os
    from taichi_2d_vof_b200 import VofSolver3D, reference_params3d
    from taichi_2d_vof_b200.driver import main
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        common = ["-ic", "3", "--nx", "96", "--ny", "80", "--scaled"]
        assert main(common + ["--steps", "23", "--dump", "full.npz"]) == 0
        assert main(common + ["--steps", "9", "--dump", "half.npz"]) == 0
        assert main(common + ["--steps", "23", "--resume", "half.npz", "--dump", "cont.npz"]) == 0
    finally:
        os.chdir(cwd)
    a, b = np.load(tmp_path / "full.npz"), np.load(tmp_path / "cont.npz")
    assert int(a["istep"]) == int(b["istep"]) == 23
    for k in ("u", "v", "p", "F"):
        assert np.array_equal(a[k], b[k]), f"2-D restart: field {k} differs"
    P3 = dict(nx=20, ny=16, nz=40, Lx=0.01, Ly=0.008, Lz=0.02)
    s = VofSolver3D(reference_params3d(**P3)); s.set_init_F(1); s.run(11)
    t = VofSolver3D(reference_params3d(**P3)); t.set_init_F(1); t.run(4)
    r = VofSolver3D(reference_params3d(**P3))
    for k in ("u", "v", "w", "p", "F"):
        getattr(r, k).from_numpy(getattr(t, k).to_numpy())
    r.istep = t.istep
    r.run(7)
    for k in ("u", "v", "w", "p", "F"):
        assert np.array_equal(getattr(s, k).to_numpy(), getattr(r, k).to_numpy()), f"3-D restart: field {k} differs"
