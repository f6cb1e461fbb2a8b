"""The whole-step tile kernel (csrc/vof2d_tile.cuh: one launch per step, the step's dependency radius as a shared-
memory halo) against the oracle and against the streaming kernels -- every element of every live field, bit for bit.
Cases: the reference's own 200 x 200 configuration for -ic 1/2/3 (tiles of 9 x 34 owned cells, 138 blocks), odd sizes
whose last tiles are ragged, live random states with CFL ~ 0.3 velocities (the halo depth n_jacobi + 5 is exactly the
dependency radius: one cell less and owned cells next to a tile edge differ), other sweep counts, graph replay."""
import numpy as np
import pytest

from oracle.vof2d_oracle import Vof2DOracle, Vof2DParams

pytestmark = pytest.mark.gpu

CORE = ("F", "u", "v", "p", "kappa", "u_star", "v_star")


def _solver(P, tile):
    from taichi_2d_vof_b200 import VofSolver2D, _lib, reference_params
    s = VofSolver2D(reference_params(nx=P.nx, ny=P.ny, Lx=P.Lx, Ly=P.Ly, n_jacobi=P.n_jacobi))
    s.set_option(_lib.VOF_OPT_TILE, tile)
    return s


def _same(s, o, tag):
    for k in CORE:
        a, b = getattr(s, k).to_numpy(), getattr(o, k)
        bad = np.argwhere(a.view(np.uint32) != np.ascontiguousarray(b, np.float32).view(np.uint32))
        assert bad.size == 0, f"{tag}: {k} differs in {len(bad)} cells, first {tuple(bad[0])}: {a[tuple(bad[0])]!r} vs {b[tuple(bad[0])]!r}"


@pytest.mark.parametrize("ic", [1, 2, 3])
def test_tile_step_equals_oracle_reference_configuration(built_lib, ic):
    P = Vof2DParams()                       # 200 x 200, the reference's constants
    o = Vof2DOracle(P)
    s = _solver(P, 2)
    o.set_init_F(ic); s.set_init_F(ic)
    for step in range(1, 31):
        o.step(); s.step()
        if step in (1, 2, 3, 10, 30):
            _same(s, o, f"-ic {ic} step {step}")
    assert s.launch_count() < 30 * 4         # one kernel (+ one memset) per step
    assert s.diagnostics(residual=False)["courant_count"] == o.courant_flags


def _live_state(nx, ny, seed, vel_scale, P):
    rng = np.random.default_rng(seed)
    shp = (nx + 2, ny + 2)
    ii, jj = np.meshgrid(np.arange(shp[0]), np.arange(shp[1]), indexing="ij")
    F = ((np.sin(ii / 3.1) + np.cos(jj / 2.3) + 0.6 * rng.standard_normal(shp)) > 0.2).astype(np.float32)
    F = np.clip(F + 0.3 * (rng.random(shp, dtype=np.float32) - 0.5) * (rng.random(shp) < 0.3), 0, 1).astype(np.float32)
    vel = vel_scale * P.dx / P.dt
    u = ((rng.random(shp, dtype=np.float32) - 0.5) * 2 * vel).astype(np.float32)
    v = ((rng.random(shp, dtype=np.float32) - 0.5) * 2 * vel).astype(np.float32)
    p = ((rng.random(shp, dtype=np.float32) - 0.5) * 200).astype(np.float32)
    return F, u, v, p


@pytest.mark.parametrize("shape,n_jacobi,seed", [((97, 131), 10, 1), ((64, 35), 10, 2), ((150, 70), 2, 3), ((33, 140), 13, 4),
                                                 ((200, 200), 7, 5),
                                                 # the far ghost row / column would be a block's only owned line (65 = 16 * 4 + 1 rows,
                                                 # 69 = 2 * 34 + 1 and 103 = 3 * 34 + 1 columns): it goes with row nx / column ny
                                                 ((63, 67), 10, 6), ((50, 101), 10, 7)])
def test_tile_step_on_live_states_equals_oracle(built_lib, shape, n_jacobi, seed):
    nx, ny = shape
    P = Vof2DParams(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200, n_jacobi=n_jacobi)
    F, u, v, p = _live_state(nx, ny, seed, 0.3, P)
    o = Vof2DOracle(P)
    s = _solver(P, 2)
    for k, a in (("F", F), ("u", u), ("v", v), ("p", p)):
        getattr(o, k)[...] = a; getattr(s, k).from_numpy(a)
    for step in range(1, 5):
        o.step(); s.step()
        _same(s, o, f"{shape} n_jacobi {n_jacobi} step {step}")


@pytest.mark.parametrize("nx,ny,Lx,Ly,n_jacobi", [(96, 64, 0.1, 0.1, 10), (50, 120, 0.05, 0.03, 4), (70, 70, 0.035, 0.035, 0), (70, 70, 0.035, 0.035, 1)])
def test_tile_step_non_square_cells_and_few_sweeps(built_lib, nx, ny, Lx, Ly, n_jacobi):
    """dx != dy (the Poisson coefficients differ by direction) and the shallow halos of n_jacobi = 0, 1."""
    P = Vof2DParams(nx=nx, ny=ny, Lx=Lx, Ly=Ly, n_jacobi=n_jacobi)
    F, u, v, p = _live_state(nx, ny, 9, 0.2, P)
    o = Vof2DOracle(P)
    s = _solver(P, 2)
    for k, a in (("F", F), ("u", u), ("v", v), ("p", p)):
        getattr(o, k)[...] = a; getattr(s, k).from_numpy(a)
    for step in range(1, 4):
        o.step(); s.step()
        _same(s, o, f"{nx}x{ny} L=({Lx},{Ly}) n_jacobi {n_jacobi} step {step}")


def test_tile_and_streaming_paths_interleave(built_lib):
    """Tile steps exchange the u / v buffers with the (dead) rho / nu buffers and flip F / p; streaming steps, single
    kernel entries and graph replay in between must keep working on the live buffers."""
    from taichi_2d_vof_b200 import _lib
    P = Vof2DParams(nx=120, ny=90, Lx=0.06, Ly=0.045)
    o = Vof2DOracle(P)
    s = _solver(P, 2)
    o.set_init_F(3); s.set_init_F(3)
    for _ in range(3):
        o.step(); s.step()                    # tile
    s.set_option(_lib.VOF_OPT_TILE, 0)
    for _ in range(3):
        o.step(); s.step()                    # streaming kernels
    _same(s, o, "after tile + streaming")
    s.set_option(_lib.VOF_OPT_TILE, 2)
    o.step(); s.step_sequence()               # one C-ABI entry per reference kernel (materialises rho / nu)
    o.run(9); s.run(9)                        # graph replay of tile steps (odd count: buffers end up swapped)
    _same(s, o, "after sequence + graph replay")
    o.run(4); s.run(4)
    _same(s, o, "after a second replay")


def test_tile_default_policy(built_lib, monkeypatch):
    """Default (VOF_OPT_TILE = 1): the tile kernel at the reference's 200 x 200, the streaming kernels at 2048 x 2048."""
    from taichi_2d_vof_b200 import VofSolver2D, scaled_params
    monkeypatch.delenv("VOF_TILE", raising=False)          # conftest.py pins the rest of the suite to the streaming kernels
    small = VofSolver2D(scaled_params(200)); small.set_init_F(1)
    l0 = small.launch_count(); small.step(); small.step()
    assert small.launch_count() - l0 <= 4
    big = VofSolver2D(scaled_params(2048)); big.set_init_F(1)
    l0 = big.launch_count(); big.step()
    assert big.launch_count() - l0 > 8
