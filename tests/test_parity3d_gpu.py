"""GPU parity of the 3-D path (3dvof.py) against the oracle: identical fields, call by call and over many steps."""
import numpy as np
import pytest

from oracle.c_oracle import Vof3DCOracle
from oracle.vof3d_oracle import Vof3DOracle, Vof3DParams

pytestmark = pytest.mark.gpu
ALL = ("F", "u", "v", "w", "p", "rho", "nu", "u_star", "v_star", "w_star")
CORE = ("F", "u", "v", "w", "p", "u_star", "v_star", "w_star")


def _solver(P, **kw):
    from taichi_2d_vof_b200 import VofSolver3D, reference_params3d
    return VofSolver3D(reference_params3d(nx=P.nx, ny=P.ny, nz=P.nz, Lx=P.Lx, Ly=P.Ly, Lz=P.Lz, n_jacobi=P.n_jacobi, **kw))


def _same(s, o, fields, tag):
    for k in fields:
        a, b = getattr(s, k).to_numpy(), getattr(o, k)
        bad = np.argwhere(a != b)
        assert bad.size == 0, f"{tag} field {k}: {len(bad)} cells differ, first {bad[0]}: {a[tuple(bad[0])]} vs {b[tuple(bad[0])]}"


@pytest.mark.parametrize("shape", [(24, 24, 24), (17, 30, 150), (40, 9, 21), (8, 12, 128), (6, 5, 256)])
def test_each_kernel_in_sequence_3d(built_lib, shape):
    nx, ny, nz = shape
    P = Vof3DParams(nx=nx, ny=ny, nz=nz, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200, Lz=0.1 * nz / 200)
    o = Vof3DOracle(P); o.set_init_F(1)
    s = _solver(P); s.set_init_F(1)
    _same(s, o, ("F",), "init")
    for step in range(1, 8):
        o.istep += 1; s.istep += 1
        for name in ("cal_nu_rho", "advect_upwind", "set_BC"):
            getattr(o, name)(); getattr(s, name)()
            _same(s, o, ALL, f"step {step} after {name}")
        for _ in range(P.n_jacobi):
            o.solve_p_jacobi(); s.solve_p_jacobi()
        _same(s, o, ALL, f"step {step} after jacobi")
        for name in ("update_uv", "set_BC", "solve_VOF_rudman", "post_process_f", "set_BC"):
            getattr(o, name)(); getattr(s, name)()
            _same(s, o, ALL, f"step {step} after {name}")


def test_random_state_3d(built_lib):
    rng = np.random.default_rng(3)
    P = Vof3DParams(nx=20, ny=26, nz=140, Lx=0.01, Ly=0.013, Lz=0.07)
    o = Vof3DOracle(P)
    shp = o.F.shape
    o.F[...] = rng.random(shp, dtype=np.float32)
    for k in ("u", "v", "w"):
        getattr(o, k)[...] = (rng.random(shp, dtype=np.float32) - 0.5) * 2.0
    o.p[...] = (rng.random(shp, dtype=np.float32) - 0.5) * 100.0
    s = _solver(P)
    for k in ("F", "u", "v", "w", "p"):
        getattr(s, k).from_numpy(getattr(o, k))
    for step in range(4):      # one full rotation of the sweep order + 1
        o.step(); s.step(materialize_props=True)
        _same(s, o, ALL, f"random step {step + 1}")


@pytest.mark.parametrize("mode", ["fused", "sequence", "no_fusion"])
def test_dam_break_3d_60_steps(built_lib, mode):
    P = Vof3DParams.scaled(48)
    o = Vof3DCOracle(P); o.set_init_F(1); o.run(60)
    s = _solver(P); s.set_init_F(1)
    for _ in range(60):
        if mode == "fused":
            s.step()
        elif mode == "sequence":
            s.step_sequence()
        else:
            s.step(no_fusion=True)
    _same(s, o, CORE, f"60 steps ({mode})")
    d = s.diagnostics()
    assert abs(d["mass"] - o.mass()) <= 1e-9 * o.mass()
    assert d["courant_count"] == o.courant_flags


def test_ic_2_and_3_leave_F_zero(built_lib):
    """3dvof.py:126-138 accepts -ic 2/3 but only -ic 1 writes F."""
    P = Vof3DParams.scaled(16)
    s = _solver(P)
    s.set_init_F(2); s.set_init_F(3)
    assert not s.F.to_numpy().any()


@pytest.mark.parametrize("nslabs", [2, 3])
def test_plane_slabs_equal_full_domain_3d(built_lib, nslabs):
    """3-D slabs along i (deep halo of whole planes, one exchange per step) reproduce the single-domain run."""
    from taichi_2d_vof_b200 import VofSolver3D, reference_params3d
    from taichi_2d_vof_b200.slab import LocalSlabGroup
    nx, ny, nz = 96, 20, 40

    def params_fn(slab, halo, device):
        return reference_params3d(nx=nx, ny=ny, nz=nz, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200, Lz=0.1 * nz / 200,
                                  slab=slab, halo=halo, device=device)

    full = VofSolver3D(params_fn(None, 0, 0)); full.set_init_F(1)
    grp = LocalSlabGroup(params_fn, nx, nslabs, solver_cls=VofSolver3D, halo_fields=("F", "u", "v", "w", "p"))
    grp.set_init_F(1)
    for step in range(1, 14):
        full.step(); grp.step()
        if step in (1, 3, 13):
            for k in ("F", "u", "v", "w", "p"):
                a, b = grp.gather(k), getattr(full, k).to_numpy()
                bad = np.argwhere(a != b)
                assert bad.size == 0, f"step {step} field {k}: {len(bad)} cells differ, first {bad[0]}"


def test_generations_identical_3d(built_lib):
    """Second-generation 3-D kernels (default) against the first generation: 12 steps at 128 x 96 x 160, every field."""
    from taichi_2d_vof_b200 import VofSolver3D, _lib, reference_params3d
    out = []
    for gen2 in (1, 0):
        s = VofSolver3D(reference_params3d(nx=128, ny=96, nz=160, Lx=0.064, Ly=0.048, Lz=0.08))
        s.set_option(_lib.VOF_OPT_ADAPTIVE, gen2)
        s.set_init_F(1)
        for _ in range(12):
            s.step()
        out.append({k: getattr(s, k).to_numpy() for k in CORE})
    for k in CORE:
        assert out[0][k].tobytes() == out[1][k].tobytes(), f"field {k} differs between the kernel generations"


@pytest.mark.parametrize("shape", [(40, 12, 128), (21, 9, 150), (36, 7, 260), (5, 6, 8), (4, 4, 8), (9, 11, 130)])
def test_multi_sweep_call_on_a_random_state_3d(built_lib, shape):
    """`solve_p_jacobi(7)` -- hoisted rhs + seven launches of the second-generation 7-point sweep (k3_jacobi5; a single-sweep
    call runs the first-generation kernel) -- on a random state with exact zeros in p, against 7 sweeps of the oracle
    (3dvof.py:261-283): p with ghosts, every bit.  Shapes: full strips, a float4 straddling nz, three strips with a
    4-column tail, tiny grids, and several plane chunks."""
    nx, ny, nz = shape
    rng = np.random.default_rng(11)
    P = Vof3DParams(nx=nx, ny=ny, nz=nz, Lx=5e-4 * nx, Ly=5e-4 * ny, Lz=5e-4 * nz)
    o = Vof3DOracle(P)
    shp = o.F.shape
    o.F[...] = rng.random(shp, dtype=np.float32)
    for k in ("u_star", "v_star", "w_star"):
        getattr(o, k)[...] = (rng.random(shp, dtype=np.float32) - 0.5) * 2.0
    o.p[...] = (rng.random(shp, dtype=np.float32) - 0.5) * 100.0
    o.p[1:-1, 1:-1, 1:-1][rng.random((nx, ny, nz)) < 0.3] = 0.0        # zero numerators and signed zeros on the way
    init = {k: getattr(o, k).copy() for k in ("F", "u_star", "v_star", "w_star", "p")}
    o.cal_nu_rho()
    for _ in range(7):
        o.solve_p_jacobi()
    s = _solver(P)
    for k, v in init.items():
        getattr(s, k).from_numpy(v)
    s.cal_nu_rho()
    s.solve_p_jacobi(7)
    got = s.p.to_numpy()
    bad = np.argwhere(got.view(np.uint32) != o.p.view(np.uint32))
    assert bad.size == 0, f"{len(bad)} cells of p differ from the oracle, first {bad[0]}: {got[tuple(bad[0])]} vs {o.p[tuple(bad[0])]}"


def test_lean_bc_tracks_outside_writers_3d(built_lib):
    """The fused 3-D step skips set_BC on fields whose interior has not changed since their ghosts were filled; any writer
    other than a whole step (field_set, single entries) must bring the full calls back.  Interleave both with the oracle."""
    rng = np.random.default_rng(11)
    P = Vof3DParams(nx=16, ny=14, nz=36, Lx=0.008, Ly=0.007, Lz=0.018)
    o = Vof3DOracle(P); o.set_init_F(1)
    s = _solver(P); s.set_init_F(1)
    for step in range(1, 10):
        if step == 4:      # overwrite u (ghosts included) from outside: the next step's first set_BC matters again
            un = (rng.random(o.u.shape, dtype=np.float32) - 0.5) * 0.1
            o.u[...] = un; s.u.from_numpy(un)
        if step == 7:      # a single entry between steps
            o.post_process_f(); s.post_process_f()
            Fn = o.F.copy(); Fn[0, :, :] = 0.25; Fn[:, :, 0] = 0.5   # and ghosts that set_BC must repair
            o.F[...] = Fn; s.F.from_numpy(Fn)
        o.step(); s.step()
        _same(s, o, CORE, f"lean-bc step {step}")


import glob as _glob
import os as _os


@pytest.mark.parametrize("path", sorted(_glob.glob(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "vof3d_ic1_*.npz"))))
def test_cuda_reproduces_committed_golden_vectors_3d(built_lib, path):
    """The 3-D CUDA path against the committed fixtures (tests/golden/make_golden3d.py), without the oracle in the loop."""
    from taichi_2d_vof_b200 import VofSolver3D, reference_params3d
    g = np.load(path)
    nx, ny, nz = (int(v) for v in g["params"][:3])
    s = VofSolver3D(reference_params3d(nx=nx, ny=ny, nz=nz, Lx=float(g["params"][3]), Ly=float(g["params"][4]), Lz=float(g["params"][5])))
    s.set_init_F(1)
    assert np.array_equal(s.F.to_numpy(), g["F_init"])
    for ck in (1, 3, 12):
        s.run(ck - s.istep)
        for k in ("u", "v", "w", "p", "F"):
            assert np.array_equal(getattr(s, k).to_numpy(), g[f"{k}_{ck}"]), f"{k} after {ck} steps"
        m = float(g[f"mass_{ck}"])
        assert abs(s.mass() - m) <= 1e-6 * m
