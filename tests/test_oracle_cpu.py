"""CPU tests of the oracle: the two independent restatements (NumPy slices / C loops) must agree
bit for bit, reproduce the committed golden vectors, and satisfy the known answers and invariants
that follow from the reference source."""
import glob
import os

import numpy as np
import pytest

from oracle.c_oracle import Vof2DCOracle
from oracle.vof2d_oracle import Vof2DOracle, Vof2DParams, rel_linf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("F", "u", "v", "p", "rho", "nu", "kappa", "u_star", "v_star")


@pytest.mark.parametrize("ic", [1, 2, 3])
@pytest.mark.parametrize("shape", [(40, 40), (33, 57)])
def test_numpy_and_c_oracles_identical(ic, shape):
    nx, ny = shape
    P = Vof2DParams(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200)
    a, b = Vof2DOracle(P), Vof2DCOracle(P)
    a.set_init_F(ic); b.set_init_F(ic)
    for _ in range(12):
        a.step(); b.step()
    for k in FIELDS:
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert a.courant_flags == b.courant_flags


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "vof2d_ic*_*.npz"))))
def test_c_oracle_reproduces_golden(path):
    g = np.load(path)
    nx, ny, Lx, Ly = int(g["params"][0]), int(g["params"][1]), float(g["params"][2]), float(g["params"][3])
    ic = int(os.path.basename(path)[8])
    o = Vof2DCOracle(Vof2DParams(nx=nx, ny=ny, Lx=Lx, Ly=Ly))
    assert abs(o.P.dx - g["params"][4]) == 0 and abs(o.P.dy - g["params"][5]) == 0
    o.set_init_F(ic)
    assert np.array_equal(o.F, g["F_init"])
    for ck in (1, 10, 100):
        o.run(ck - o.istep)
        for k in ("u", "v", "p", "F", "kappa"):
            assert np.array_equal(getattr(o, k), g[f"{k}_{ck}"]), f"{k} after {ck} steps"
        assert abs(o.mass() - float(g[f"mass_{ck}"])) <= 1e-9 * float(g[f"mass_{ck}"])


def test_numpy_oracle_reproduces_golden_small():
    g = np.load(os.path.join(GOLD, "vof2d_ic3_64x96.npz"))
    o = Vof2DOracle(Vof2DParams(nx=64, ny=96, Lx=0.032, Ly=0.048))
    o.set_init_F(3)
    o.run(10)
    for k in ("u", "v", "p", "F", "kappa"):
        assert np.array_equal(getattr(o, k), g[f"{k}_10"]), k


def test_default_constants_match_reference_source():
    """2dvof.py:19-50: dx is a difference of two fp32 node coordinates, not Lx/nx."""
    P = Vof2DParams()
    assert (P.nx, P.ny, P.dt, P.n_jacobi) == (200, 200, 4e-6, 10)
    assert P.dx == float(np.float32(0.001)) - float(np.float32(0.0005))
    assert abs(P.dx - 5.000000237e-4) < 1e-13 and P.dx != 5e-4
    o = Vof2DOracle(P)
    assert o.c_dxi2 == np.float32((1 / P.dx) ** 2)


def test_var_is_arithmetic_not_a_clamp():
    """2dvof.py:192-195 quantises to multiples of 2^-23."""
    o = Vof2DOracle(Vof2DParams(nx=8, ny=8))
    f = np.float32
    assert o._var(f(0), f(1), f(0.3)) == f(0.29999995)
    assert o._var(f(0), f(1), f(3.3e-8)) == f(0)
    assert o._var(f(0), f(1), f(1.0000001)) == f(0.9999999)
    assert o._var(f(0), f(1), f(-1e-9)) > 0


def test_step1_dam_break_known_answers():
    """From rest: u* = 0, v* = dt*gy = -2e-5 in the bulk; F is unchanged where u = v = 0."""
    o = Vof2DOracle(Vof2DParams())
    o.set_init_F(1)
    F0 = o.F.copy()
    o.istep += 1
    o.cal_nu_rho(); o.get_normal_young(); o.advect_upwind()
    assert np.all(o.u_star[2:200, 1:201][:, 120:] == 0)                    # gas region far from the interface
    assert np.all(o.v_star[1:201, 2:201][100:, 120:] == np.float32(4e-6) * np.float32(-5))
    assert np.all(o.u_star[1, :] == 0) and np.all(o.u_star[201, :] == 0)     # walls never written
    assert np.all(o.v_star[:, 1] == 0) and np.all(o.v_star[:, 201] == 0)
    # rho, nu of the two pure phases
    assert o.rho[10, 10] == np.float32(1000.0) and o.rho[150, 150] == np.float32(50.0)
    # F only changes next to the interface during the first step
    o2 = Vof2DOracle(Vof2DParams()); o2.set_init_F(1); o2.step()
    changed = np.argwhere(o2.F != F0)
    assert len(changed) < 4 * 200


@pytest.mark.parametrize("ic", [1, 2, 3])
def test_bounds_and_volume(ic):
    o = Vof2DCOracle(Vof2DParams())
    o.set_init_F(ic)
    m0 = o.mass()
    o.run(100)
    assert o.F.min() >= 0.0 and o.F.max() <= 1.0
    # the scheme itself (clamps + 2^-23 quantisation) drifts by a few 1e-6 in 100 steps
    assert abs(o.mass() - m0) / m0 < 1e-5


def test_fct_transposition_symmetry():
    """fct_x on (F, u) is fct_y on (F^T, v = u^T) when dx == dy -- same arithmetic, transposed."""
    rng = np.random.default_rng(1)
    P = Vof2DParams(nx=48, ny=48, Lx=0.024, Ly=0.024)
    a, b = Vof2DOracle(P), Vof2DOracle(P)
    F = rng.random(a.F.shape, dtype=np.float32)
    u = (rng.random(a.F.shape, dtype=np.float32) - 0.5) * 4
    a.F[...] = F; a.u[...] = u
    b.F[...] = F.T; b.v[...] = u.T
    a.fct_x_sweep(); b.fct_y_sweep()
    assert np.array_equal(a.F, b.F.T)


def test_set_bc_corner_order():
    """Row loop then column loop: corners take the column loop's copy of the row loop's result."""
    rng = np.random.default_rng(2)
    o = Vof2DOracle(Vof2DParams(nx=6, ny=7, Lx=0.003, Ly=0.0035))
    for k in ("u", "v", "F", "p", "rho"):
        getattr(o, k)[...] = rng.random(o.F.shape, dtype=np.float32)
    F = o.F.copy()
    o.set_BC()
    assert o.F[0, 0] == F[1, 1] and o.F[7, 8] == F[6, 7] and o.F[0, 8] == F[1, 7]
    assert np.all(o.u[1, :] == 0) and np.all(o.u[7, :] == 0)
    assert np.all(o.v[:, 1] == 0) and np.all(o.v[:, 8] == 0)


def test_roundoff_perturbation_stays_within_north_star_tolerance():
    """Calibrates the 1e-3 @ 100 steps gate: a 1-ulp perturbation of v after the first step (the size
    of difference a differently-rounding compiler would introduce) must not grow beyond it in
    u, v, p, F.  (kappa is NOT stable in this sense: cells with |normal| ~ 1e-8 flip direction.)"""
    P = Vof2DParams(nx=100, ny=100, Lx=0.05, Ly=0.05)
    a, b = Vof2DCOracle(P), Vof2DCOracle(P)
    a.set_init_F(3); b.set_init_F(3)
    a.step(); b.step()
    b.v[...] = np.nextafter(b.v, np.float32(np.inf))
    a.run(99); b.run(99)
    for k in ("u", "v", "p", "F"):
        assert rel_linf(getattr(b, k), getattr(a, k)) < 1e-3, k


def test_3d_numpy_and_c_oracles_identical():
    from oracle.c_oracle import Vof3DCOracle
    from oracle.vof3d_oracle import Vof3DOracle, Vof3DParams
    P = Vof3DParams(nx=14, ny=19, nz=11, Lx=0.007, Ly=0.0095, Lz=0.0055)
    a, b = Vof3DOracle(P), Vof3DCOracle(P)
    a.set_init_F(1); b.set_init_F(1)
    for _ in range(9):      # three full rotations of the sweep order
        a.step(); b.step()
    for k in a.FIELDS:
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


def test_3d_known_answers():
    """From rest: v* = dt*gy in the bulk, u* = w* = 0; only -ic 1 writes F; sweep order rotates with istep % 3."""
    from oracle.vof3d_oracle import Vof3DOracle, Vof3DParams
    o = Vof3DOracle(Vof3DParams.scaled(20))
    o.set_init_F(2)
    assert not o.F.any()
    o.set_init_F(1)
    assert o.F[1, 1, 1] == 1 and o.F[15, 15, 15] == 0
    o.istep += 1
    o.cal_nu_rho(); o.advect_upwind()
    assert np.all(o.v_star[1:21, 2:21, 1:21][12:, 14:, 12:] == np.float32(4e-6) * np.float32(-5))
    assert not o.u_star[:, :, :][12:, 14:, 12:].any() and not o.w_star[12:, 14:, 12:].any()
    calls = []
    o._fct_sweep = lambda ax: calls.append(ax)
    for istep in (1, 2, 3):
        o.istep = istep; o.solve_VOF_rudman()
    assert calls == [1, 2, 0, 2, 0, 1, 0, 1, 2]


def test_display_kernels_semantics():
    """2dvof.py:458-492: rgb_buf[I] = field[I // 2] on a (2nx, 2ny) buffer, velocities over L / 0.2; V is cell-centred."""
    P = Vof2DParams(nx=12, ny=10, Lx=0.006, Ly=0.005)
    o = Vof2DOracle(P); o.set_init_F(1)
    rng = np.random.default_rng(5)
    o.u[...] = rng.random(o.u.shape, dtype=np.float32); o.v[...] = rng.random(o.v.shape, dtype=np.float32)
    rgb = o.get_vof_field()
    assert rgb.shape == (24, 20) and rgb.dtype == np.float32
    for I in ((0, 0), (1, 1), (5, 7), (23, 19)):
        assert rgb[I] == o.F[I[0] // 2, I[1] // 2]
    assert o.get_u_field()[7, 3] == np.float32(o.u[3, 1] / np.float32(P.Lx / 0.2))
    assert o.get_vnorm_field()[2, 9] == np.float32(np.sqrt(o.u[1, 4] * o.u[1, 4] + o.v[1, 4] * o.v[1, 4]) / np.float32(P.Ly / 0.2))
    V = o.interp_velocity()
    assert V.shape == (14, 12, 2) and V[3, 4, 0] == (o.u[3, 4] + o.u[4, 4]) / np.float32(2) and V[3, 4, 1] == (o.v[3, 4] + o.v[3, 5]) / np.float32(2)
    assert not V[0].any() and not V[:, 0].any() and not V[P.nx + 1].any()


def kothe_rider_state(n, cfl=0.2):
    """The single-vortex transport test of the reference's test/forward_fct.py:9-21, 196-204 (Kothe-Rider), on this
    solver's staggered layout: u = -sin^2(pi x) sin(2 pi y), v = sin^2(pi y) sin(2 pi x) on [0, 1]^2 (divergence-free
    on the MAC grid up to round-off), scaled to a Courant number; a disc of radius L/10 at (L/2, 3L/4)."""
    P = Vof2DParams(nx=n, ny=n, Lx=0.1 * n / 200, Ly=0.1 * n / 200)
    o = Vof2DOracle(P)
    dx = P.Lx / n
    xf = (np.arange(n + 2) - 1) * dx            # west faces of cells 0 .. n+1
    xc = xf + dx / 2                            # cell centres
    X, Y = xf[:, None] / P.Lx, xc[None, :] / P.Ly
    u = -np.sin(np.pi * X) ** 2 * np.sin(2 * np.pi * Y)
    X, Y = xc[:, None] / P.Lx, xf[None, :] / P.Ly
    v = np.sin(np.pi * Y) ** 2 * np.sin(2 * np.pi * X)
    s = cfl * dx / P.dt
    o.u[...] = (u * s).astype(np.float32); o.v[...] = (v * s).astype(np.float32)
    d = np.sqrt((xc[:, None] - P.Lx / 2) ** 2 + (xc[None, :] - 0.75 * P.Ly) ** 2)
    o.F[...] = np.clip((P.Lx / 10 - d) / dx + 0.5, 0.0, 1.0).astype(np.float32)
    o.set_BC()
    return P, o


def test_kothe_rider_transport_properties():
    """Pure FCT transport in the Kothe-Rider vortex (the reference's only test idea, test/forward_fct.py, applied to the
    production sweeps of 2dvof.py): F stays in [0, 1] and the volume is conserved while the disc is stretched."""
    P, o = kothe_rider_state(96)
    m0 = o.mass()
    for _ in range(80):
        o.istep += 1
        o.solve_VOF_rudman(); o.post_process_f(); o.set_BC()
    F = o.F[1:-1, 1:-1]
    assert np.isfinite(F).all() and F.min() >= 0.0 and F.max() <= 1.0
    assert abs(o.mass() - m0) / m0 < 5e-3            # conservative up to the clamps of var() (measured: -1.6e-3 after 80 steps)
    assert 0.002 < ((F > 0.01) & (F < 0.99)).mean() < 0.05   # a thin interface band: the disc has neither vanished nor smeared


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "vof3d_ic1_*.npz"))))
def test_c_oracle_reproduces_golden_3d(path):
    """The C twin of the 3-D oracle against the committed vectors (tests/golden/make_golden3d.py: NumPy oracle)."""
    from oracle.c_oracle import Vof3DCOracle
    from oracle.vof3d_oracle import Vof3DParams
    g = np.load(path)
    nx, ny, nz = (int(v) for v in g["params"][:3])
    o = Vof3DCOracle(Vof3DParams(nx=nx, ny=ny, nz=nz, Lx=float(g["params"][3]), Ly=float(g["params"][4]), Lz=float(g["params"][5])))
    o.set_init_F(1)
    assert np.array_equal(o.F, g["F_init"])
    for ck in (1, 3, 12):
        o.run(ck - o.istep)
        for k in ("u", "v", "w", "p", "F"):
            assert np.array_equal(getattr(o, k), g[f"{k}_{ck}"]), f"{k} after {ck} steps"
        assert abs(o.mass() - float(g[f"mass_{ck}"])) <= 1e-9 * float(g[f"mass_{ck}"])


# ----------------------------------------------------------------------------------------------
# dependency radius of one step along i: what the slab halo depth must cover
# ----------------------------------------------------------------------------------------------
def _band_after_step(n_jacobi, R, seed, three_d=False):
    """Two fp64 states that agree only within R rows of the band [a, b]; returns whether the band agrees after one step."""
    if three_d:
        from oracle.vof3d_oracle import Vof3DOracle, Vof3DParams
        P = Vof3DParams(nx=72, ny=5, nz=6, Lx=0.036, Ly=0.0025, Lz=0.003, n_jacobi=n_jacobi)
        mk, vels = (lambda: Vof3DOracle(P, real=np.float64)), ("u", "v", "w")
    else:
        P = Vof2DParams(nx=72, ny=20, Lx=0.036, Ly=0.01, n_jacobi=n_jacobi)
        mk, vels = (lambda: Vof2DOracle(P, real=np.float64)), ("u", "v")
    a, b = 34, 38
    out = []
    for variant in (0, 1):
        o = mk()
        rng, far = np.random.default_rng(seed), np.random.default_rng(seed + 1000 + variant)
        vel = 0.2 * P.dx / P.dt                                   # CFL ~ 0.2: upwind switches and the limiter are live
        for k in ("F",) + vels + ("p",):
            arr = getattr(o, k)
            scale = 1.0 if k in ("F", "p") else vel
            arr[...] = (rng.random(arr.shape) * (1 if k == "F" else 2) - (0 if k == "F" else 1)) * scale
            other = (far.random(arr.shape) * (1 if k == "F" else 2) - (0 if k == "F" else 1)) * scale
            arr[: a - R] = other[: a - R]
            arr[b + R + 1:] = other[b + R + 1:]
        o.step()
        out.append({k: getattr(o, k)[a:b + 1].copy() for k in ("F",) + vels + ("p",)})
    return all(np.array_equal(out[0][k], out[1][k]) for k in out[0])


@pytest.mark.parametrize("n_jacobi", [10, 4, 1])
def test_step_dependency_radius_2d(n_jacobi):
    """One 2-D step reads n_jacobi + 5 rows either side (x-FCT <- u(i+3) <- p after n sweeps <- rhs <- u* <- kappa <- F):
    rows further away cannot influence a row, rows n_jacobi + 5 away do."""
    from taichi_2d_vof_b200.slab import required_halo
    R = required_halo(n_jacobi)
    assert R == n_jacobi + 5
    for seed in range(4):
        assert _band_after_step(n_jacobi, R, seed)
    if n_jacobi <= 4:     # tightness (after 9 sweeps the farthest influence, ~4^-9 of an ulp-level term, drops out of fp64)
        assert not all(_band_after_step(n_jacobi, R - 1, seed) for seed in range(4)), "n_jacobi + 4 rows would do?"


def test_step_dependency_radius_3d():
    """3-D (no curvature): n_jacobi + 4 planes."""
    for seed in range(2):
        assert _band_after_step(10, 14, seed, three_d=True)
        assert _band_after_step(3, 7, seed, three_d=True)
