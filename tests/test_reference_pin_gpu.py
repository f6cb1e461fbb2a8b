"""The CUDA path against the REFERENCE ITSELF: ``tests/golden/ref_*.npz`` hold what the unmodified text of
/root/reference/2dvof.py and 3dvof.py computed under ``oracle/refshim/taichi`` (oracle/run_reference.py).
No oracle is involved here: libvof's fields are compared with the reference run directly, bit for bit
(sign of zero included), after every kernel call of the first steps (C-ABI entries named like the
reference kernels) and, through the fused ``vof2d_step`` / graph-replayed ``vof2d_run``, at the step
snapshots -- 100 steps at the reference's own 200 x 200 configuration."""
import numpy as np
import pytest

import refpin

pytestmark = pytest.mark.gpu

FIX2D = refpin.fixtures("2d_")
FIX3D = refpin.fixtures("3d_")
LIVE2D = ("F", "u", "v", "p", "rho", "nu", "kappa", "u_star", "v_star")
CORE2D = ("F", "u", "v", "p", "kappa", "u_star", "v_star")      # rho / nu are inlined by the fused step
LIVE3D = ("F", "u", "v", "w", "p", "rho", "nu", "u_star", "v_star", "w_star")
CORE3D = ("F", "u", "v", "w", "p", "u_star", "v_star", "w_star")


def _solver2d(z, meta):
    from taichi_2d_vof_b200 import VofSolver2D, reference_params
    n = refpin.sizes(meta)
    P = reference_params(nx=n["nx"], ny=n["ny"])
    assert (P.dx, P.dy, P.dt) == (meta["dx"], meta["dy"], meta["dt"])
    s = VofSolver2D(P)
    s.set_init_F(meta["ic"])
    refpin.assert_same(s.F.to_numpy(), z["F_init"], "set_init_F")
    for k in ("u", "v", "p", "F"):
        if k + "_in" in z.files:
            getattr(s, k).from_numpy(z[k + "_in"])
    return s


def _solver3d(z, meta):
    from taichi_2d_vof_b200 import VofSolver3D, reference_params3d
    n = refpin.sizes(meta)
    P = reference_params3d(nx=n["nx"], ny=n["ny"], nz=n["nz"])
    assert (P.dx, P.dy, P.dz) == (meta["dx"], meta["dy"], meta["dz"])
    s = VofSolver3D(P)
    s.set_init_F(meta["ic"])
    refpin.assert_same(s.F.to_numpy(), z["F_init"], "set_init_F")
    for k in ("u", "v", "w", "p", "F"):
        if k + "_in" in z.files:
            getattr(s, k).from_numpy(z[k + "_in"])
    return s


@pytest.mark.parametrize("name", [n for n in FIX2D if "200x200" not in n])
def test_cuda_kernels_equal_reference_run_call_by_call_2d(built_lib, name):
    z, meta = refpin.load(name)
    s = _solver2d(z, meta)
    for c, (kname, ref) in enumerate(refpin.calls(z, meta)):
        if kname == "cal_nu_rho":
            s.istep += 1
        getattr(s, kname)()
        for k in LIVE2D:
            refpin.assert_same(getattr(s, k).to_numpy(), ref[k], f"{name} call {c} ({kname}) field {k}")


@pytest.mark.parametrize("mode", ["fused", "graph", "sequence", "tile", "tile_graph"])
@pytest.mark.parametrize("name", FIX2D)
def test_cuda_steps_equal_reference_run_2d(built_lib, name, mode):
    """fused / graph: the streaming kernels; tile / tile_graph: the whole-step tile kernel (one launch per step)."""
    from taichi_2d_vof_b200 import _lib
    z, meta = refpin.load(name)
    s = _solver2d(z, meta)
    s.set_option(_lib.VOF_OPT_TILE, 2 if mode.startswith("tile") else 0)
    for t in meta["steps"]:
        if mode.endswith("graph"):
            s.run(t - s.istep)
        else:
            while s.istep < t:
                s.step_sequence() if mode == "sequence" else s.step()
        for k in (LIVE2D if mode == "sequence" else CORE2D):
            refpin.assert_same(getattr(s, k).to_numpy(), z[f"{k}_{t}"], f"{name} step {t} ({mode}) field {k}")


@pytest.mark.parametrize("name", FIX3D)
def test_cuda_kernels_equal_reference_run_call_by_call_3d(built_lib, name):
    z, meta = refpin.load(name)
    s = _solver3d(z, meta)
    for c, (kname, ref) in enumerate(refpin.calls(z, meta)):
        if kname == "cal_nu_rho":
            s.istep += 1
        getattr(s, kname)()
        for k in LIVE3D:
            refpin.assert_same(getattr(s, k).to_numpy(), ref[k], f"{name} call {c} ({kname}) field {k}")


@pytest.mark.parametrize("mode", ["fused", "sequence"])
@pytest.mark.parametrize("name", FIX3D)
def test_cuda_steps_equal_reference_run_3d(built_lib, name, mode):
    z, meta = refpin.load(name)
    s = _solver3d(z, meta)
    for t in meta["steps"]:
        while s.istep < t:
            s.step() if mode == "fused" else s.step_sequence()
        for k in (LIVE3D if mode == "sequence" else CORE3D):
            refpin.assert_same(getattr(s, k).to_numpy(), z[f"{k}_{t}"], f"{name} step {t} ({mode}) field {k}")


@pytest.mark.parametrize("name", refpin.fixtures("fct_"))
def test_cuda_forward_fct_equals_reference_run(built_lib, name):
    """vof2d_fct_forward (the stand-alone FCT variant, test/forward_fct.py:254-351) on the Kothe-Rider vortex of the
    reference run: every stored half-step level, bit for bit.  Odd levels (between the two sweeps of a step) are
    checked through a second solver that stops after the first sweep's level by replaying the reference order."""
    from taichi_2d_vof_b200 import VofSolver2D, reference_params
    z, meta = refpin.load(name)
    n = refpin.sizes(meta)
    P = reference_params(nx=n["nx"], ny=n["ny"], Lx=float(np.pi), Ly=float(np.pi), dt=meta["dt"])
    P.dx, P.dy = meta["dx"], meta["dy"]
    s = VofSolver2D(P)
    assert (s.P.dx, s.P.dy, s.P.dt) == (meta["dx"], meta["dy"], meta["dt"])
    levels = [int(v) for v in z["levels"]]
    want = dict(zip(levels, z["F_levels"]))
    s.F.from_numpy(want[0])
    s.u.from_numpy(z["u"])
    s.v.from_numpy(z["v"])
    checked = 0
    for t in range(n["tmax"]):
        s.fct_forward(meta["eps"])
        if 2 * t + 2 in want:
            refpin.assert_same(s.F.to_numpy(), want[2 * t + 2], f"{name} level {2 * t + 2}")
            checked += 1
    assert checked == sum(1 for v in levels if v and v % 2 == 0)
