"""Interface-adaptive kernels (VOF_OPT_ADAPTIVE, default on): the warp-uniform bulk short-cuts of the FCT sweeps and
of the curvature kernel, and the cp.async row ring of the x-sweep, must leave every bit unchanged.

The short-cuts fire on rows where a whole warp strip (64 / 120 / 128 columns) holds one value, and switch back to
the general pipeline where that stops, so the states here are built from large uniform blocks (0, 1 and a third
constant) with ragged edges, noise patches and zero / tiny / random velocities: every transition between the modes
happens many times per sweep, at walls and away from them.  Checked against the oracle and against the
first-generation kernels (adaptive off).
"""
import numpy as np
import pytest

from oracle.vof2d_oracle import Vof2DOracle, Vof2DParams

pytestmark = pytest.mark.gpu


def _solver(P, adaptive, cols=None):
    from taichi_2d_vof_b200 import VofSolver2D, _lib, reference_params
    s = VofSolver2D(reference_params(nx=P.nx, ny=P.ny, Lx=P.Lx, Ly=P.Ly, n_jacobi=P.n_jacobi))
    s.set_option(_lib.VOF_OPT_ADAPTIVE, adaptive)
    if cols:
        s.set_option(_lib.VOF_OPT_FCT_X_COLS, cols)
    return s


def blocky_state(nx, ny, seed, vel_scale=0.5):
    """F made of big uniform blocks with ragged borders and noise patches; u, v random / zero / tiny by block."""
    rng = np.random.default_rng(seed)
    shape = (nx + 2, ny + 2)
    F = np.zeros(shape, np.float32)
    u = ((rng.random(shape, dtype=np.float32) - 0.5) * 2 * vel_scale).astype(np.float32)
    v = ((rng.random(shape, dtype=np.float32) - 0.5) * 2 * vel_scale).astype(np.float32)
    vals = [0.0, 1.0, 0.5, 1.0, 0.0]
    for _ in range(14):
        i0, i1 = sorted(rng.integers(0, nx + 2, 2)); j0, j1 = sorted(rng.integers(0, ny + 2, 2))
        F[i0:i1 + 1, j0:j1 + 1] = vals[rng.integers(len(vals))]
    for _ in range(6):       # full-width bands: whole rows uniform across every strip
        i0 = rng.integers(0, nx + 2); h = rng.integers(1, 24)
        F[i0:i0 + h, :] = vals[rng.integers(len(vals))]
    for _ in range(6):       # noise patches and single odd cells
        i0, j0 = rng.integers(0, nx), rng.integers(0, ny)
        h, w = rng.integers(1, 12), rng.integers(1, 40)
        F[i0:i0 + h, j0:j0 + w] = rng.random((min(h, nx + 2 - i0), min(w, ny + 2 - j0)), dtype=np.float32)
    for _ in range(10):
        F[rng.integers(0, nx + 2), rng.integers(0, ny + 2)] = rng.random(dtype=np.float32)
    for _ in range(8):       # velocity: dead zones, tiny values, one-signed zones
        i0, i1 = sorted(rng.integers(0, nx + 2, 2)); j0, j1 = sorted(rng.integers(0, ny + 2, 2))
        kind = rng.integers(4)
        for a in (u, v):
            blk = a[i0:i1 + 1, j0:j1 + 1]
            if kind == 0: blk[...] = 0.0
            elif kind == 1: blk *= np.float32(1e-9)
            elif kind == 2: blk[...] = np.abs(blk)
            else: blk[...] = np.float32(1e-38) * np.sign(blk)
    p = ((rng.random(shape, dtype=np.float32) - 0.5) * 10).astype(np.float32)
    return F, u, v, p


@pytest.mark.parametrize("cols", [2, 4])
@pytest.mark.parametrize("shape,seed", [((300, 700), 1), ((97, 1030), 2), ((520, 260), 3), ((64, 129), 4)])
def test_blocky_fct_and_kappa_entries_match_oracle(built_lib, shape, seed, cols):
    """Each entry on its own, many mode transitions: fct_x, fct_y (both POST variants via the rudman order), kappa."""
    nx, ny = shape
    P = Vof2DParams(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200)
    F, u, v, p = blocky_state(nx, ny, seed, vel_scale=20.0)   # |u| dt / dx up to 0.16: fluxes matter
    for adaptive in (1, 0):
        o = Vof2DOracle(P)
        s = _solver(P, adaptive, cols)
        for k, a in (("F", F), ("u", u), ("v", v), ("p", p)):
            getattr(o, k)[...] = a; getattr(s, k).from_numpy(a)
        for name in ("get_normal_young", "fct_x_sweep", "fct_y_sweep", "get_normal_young", "fct_y_sweep", "fct_x_sweep",
                     "post_process_f", "get_normal_young"):
            getattr(o, name)(); getattr(s, name)()
            for fld in ("F", "kappa"):
                a, b = getattr(s, fld).to_numpy(), getattr(o, fld)
                bad = np.argwhere(a != b)
                assert bad.size == 0, (f"adaptive={adaptive} cols={cols} after {name}: {fld} differs in {len(bad)} cells, "
                                       f"first {bad[0]}: {a[tuple(bad[0])]} vs {b[tuple(bad[0])]}")


@pytest.mark.parametrize("shape,seed", [((300, 700), 11), ((200, 390), 12)])
def test_blocky_whole_steps_match_oracle(built_lib, shape, seed):
    nx, ny = shape
    P = Vof2DParams(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200)
    F, u, v, p = blocky_state(nx, ny, seed, vel_scale=0.5)
    o = Vof2DOracle(P)
    s = _solver(P, 1)
    for k, a in (("F", F), ("u", u), ("v", v), ("p", p)):
        getattr(o, k)[...] = a; getattr(s, k).from_numpy(a)
    for step in range(4):
        o.step(); s.step()
        for fld in ("F", "u", "v", "p", "kappa"):
            a, b = getattr(s, fld).to_numpy(), getattr(o, fld)
            assert np.array_equal(a, b), f"step {step + 1}: {fld} differs in {(a != b).sum()} cells"


@pytest.mark.parametrize("ic", [1, 2, 3])
def test_adaptive_on_off_identical_1024(built_lib, ic):
    """The three initial conditions at a size where most strips are bulk: 30 steps, every field bit-identical."""
    from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params
    out = []
    for adaptive in (1, 0):
        s = VofSolver2D(scaled_params(1024)); s.set_option(_lib.VOF_OPT_ADAPTIVE, adaptive); s.set_init_F(ic)
        for _ in range(30):
            s.step()
        out.append({k: getattr(s, k).to_numpy() for k in ("F", "u", "v", "p", "kappa")})
    for k in out[0]:
        assert out[0][k].tobytes() == out[1][k].tobytes(), f"ic {ic}: field {k} differs between adaptive on and off"


def test_adaptive_on_off_identical_8192(built_lib):
    """Benchmark size and workload: 6 steps with the adaptive kernels == 6 steps with the first-generation kernels."""
    from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params
    ref = None
    for adaptive in (1, 0):
        s = VofSolver2D(scaled_params(8192)); s.set_option(_lib.VOF_OPT_ADAPTIVE, adaptive); s.set_init_F(3)
        for _ in range(6):
            s.step()
        cur = {k: getattr(s, k).to_numpy() for k in ("F", "u", "v", "p")}
        del s
        if ref is None:
            ref = cur
        else:
            for k in ref:
                assert np.array_equal(ref[k], cur[k]), f"field {k} differs between adaptive on and off at 8192^2"
