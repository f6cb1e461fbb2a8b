"""Pins the oracles to the REFERENCE ITSELF (CPU, no GPU needed).

``tests/golden/ref_*.npz`` hold what the unmodified source text of /root/reference/2dvof.py, 3dvof.py and
test/forward_fct.py computed when executed under ``oracle/refshim/taichi`` (``oracle/run_reference.py``; the
real taichi==1.4.1 cannot be installed here).  These tests demand that the hand-written oracles --
the NumPy restatement and the C/OpenMP one that doubles as the CPU baseline -- reproduce those runs BIT FOR
BIT: every element of every field (live state and the reference's scratch arrays), sign of zero included,
after every single kernel call of the first steps and at the step snapshots.  2dvof.py at its own 200 x 200
default is executed with no substitution at all; the other grids substitute only nx/ny(/nz).
"""
import os

import numpy as np
import pytest

import refpin
from oracle.c_oracle import Vof2DCOracle, Vof3DCOracle
from oracle.vof2d_oracle import Vof2DOracle, Vof2DParams
from oracle.vof3d_oracle import Vof3DOracle, Vof3DParams

FIX2D = refpin.fixtures("2d_")
FIX3D = refpin.fixtures("3d_")
SCRATCH2D = ("Ftd", "ax", "ay", "cx", "cy", "rp", "rm", "pt", "mx", "my")
LIVE2D = ("F", "u", "v", "p", "rho", "nu", "kappa", "u_star", "v_star")
LIVE3D = ("F", "u", "v", "w", "p", "rho", "nu", "u_star", "v_star", "w_star")


def _oracle2d(z, meta, cls=Vof2DOracle):
    n = refpin.sizes(meta)
    P = Vof2DParams(nx=n["nx"], ny=n["ny"])          # Lx = Ly = 0.1 as in the text; dx != dy when nx != ny
    assert (P.dx, P.dy, P.dt) == (meta["dx"], meta["dy"], meta["dt"])
    o = cls(P)
    o.set_init_F(meta["ic"])
    return o


def _inject(o, z, names):
    for k in names:
        if k + "_in" in z.files:
            getattr(o, k)[...] = z[k + "_in"]        # NumPy arrays, or zero-copy views of the C oracle's state


def test_fixtures_present():
    assert len(FIX2D) >= 9 and len(FIX3D) >= 3 and len(refpin.fixtures("fct_")) >= 2
    for ic in (1, 2, 3):                              # the reference's own configuration, text untouched
        z, meta = refpin.load(f"2d_ic{ic}_200x200")
        assert meta["substitutions"] == [] and meta["script"] == "2dvof.py" and 100 in meta["steps"]


@pytest.mark.parametrize("name", FIX2D)
def test_numpy_oracle_equals_reference_run_2d(name):
    z, meta = refpin.load(name)
    o = _oracle2d(z, meta)
    refpin.assert_same(o.F, z["F_init"], f"{name} set_init_F")
    _inject(o, z, ("u", "v", "p", "F"))
    # (1) kernel by kernel: drive the oracle with the reference's recorded call sequence
    ncheck = 0
    for c, (kname, ref) in enumerate(refpin.calls(z, meta)):
        if kname == "cal_nu_rho":
            o.istep += 1
        getattr(o, kname)()
        for k, b in ref.items():
            if hasattr(o, k):
                refpin.assert_same(getattr(o, k), b, f"{name} call {c} ({kname}) field {k}")
                ncheck += 1
    assert ncheck or meta["kernel_steps"] == 0
    # (2) the oracle's own step sequencing, snapshots at the recorded steps
    o = _oracle2d(z, meta)
    _inject(o, z, ("u", "v", "p", "F"))
    for s in meta["steps"]:
        o.run(s - o.istep)
        for k in LIVE2D + SCRATCH2D:
            if f"{k}_{s}" in z.files:
                refpin.assert_same(getattr(o, k), z[f"{k}_{s}"], f"{name} step {s} field {k}")


@pytest.mark.parametrize("name", FIX2D)
def test_c_oracle_equals_reference_run_2d(name):
    """The C/OpenMP oracle (bench.py's cpu_baseline and --impl reference arm) against the same runs."""
    z, meta = refpin.load(name)
    live = ("F", "u", "v", "p", "rho", "nu", "kappa", "u_star", "v_star")
    o = _oracle2d(z, meta, Vof2DCOracle)
    refpin.assert_same(o.F, z["F_init"], f"{name} set_init_F (C oracle)")
    _inject(o, z, ("u", "v", "p", "F"))
    for c, (kname, ref) in enumerate(refpin.calls(z, meta)):
        if kname == "cal_nu_rho":
            o.istep = o.istep + 1
        getattr(o, kname)()
        for k in live:
            refpin.assert_same(getattr(o, k), ref[k], f"{name} call {c} ({kname}) field {k} (C oracle)")
    o = _oracle2d(z, meta, Vof2DCOracle)
    _inject(o, z, ("u", "v", "p", "F"))
    for s in meta["steps"]:
        o.run(s - o.istep)
        for k in live:
            refpin.assert_same(getattr(o, k), z[f"{k}_{s}"], f"{name} step {s} field {k} (C oracle)")


@pytest.mark.parametrize("name", FIX3D)
def test_numpy_oracle_equals_reference_run_3d(name):
    z, meta = refpin.load(name)
    n = refpin.sizes(meta)
    P = Vof3DParams(nx=n["nx"], ny=n["ny"], nz=n["nz"])
    assert (P.dx, P.dy, P.dz) == (meta["dx"], meta["dy"], meta["dz"])

    def fresh():
        o = Vof3DOracle(P)
        o.set_init_F(meta["ic"])
        refpin.assert_same(o.F, z["F_init"], f"{name} set_init_F")
        _inject(o, z, ("u", "v", "w", "p", "F"))
        return o

    o = fresh()
    for c, (kname, ref) in enumerate(refpin.calls(z, meta)):
        if kname == "cal_nu_rho":
            o.istep += 1
        getattr(o, kname)()
        for k, b in ref.items():
            if hasattr(o, k) and isinstance(getattr(o, k), np.ndarray):
                refpin.assert_same(getattr(o, k), b, f"{name} call {c} ({kname}) field {k}")
    o = fresh()
    for s in meta["steps"]:
        o.run(s - o.istep)
        for k in LIVE3D:
            refpin.assert_same(getattr(o, k), z[f"{k}_{s}"], f"{name} step {s} field {k}")


@pytest.mark.parametrize("name", FIX3D)
def test_c_oracle_equals_reference_run_3d(name):
    z, meta = refpin.load(name)
    n = refpin.sizes(meta)
    o = Vof3DCOracle(Vof3DParams(nx=n["nx"], ny=n["ny"], nz=n["nz"]))
    o.set_init_F(meta["ic"])
    _inject(o, z, ("u", "v", "w", "p", "F"))
    for s in meta["steps"]:
        o.run(s - o.istep)
        for k in LIVE3D:
            refpin.assert_same(getattr(o, k), z[f"{k}_{s}"], f"{name} step {s} field {k} (3-D C oracle)")


@pytest.mark.skipif(not os.path.exists("/root/reference/2dvof.py"), reason="needs the reference checkout")
def test_fixture_is_what_the_reference_text_computes_today():
    """Re-executes the reference text (build container only) and compares with the committed fixture,
    with ti.max/min returning either operand on ties (the result must not depend on it)."""
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for mode in ("first", "second"):
        with tempfile.TemporaryDirectory() as d:
            env = dict(os.environ, REF_OUT_DIR=d, TI_SHIM_MAXMIN=mode)
            subprocess.run([sys.executable, os.path.join(root, "oracle", "run_reference.py"), "2d_synth_26x18_b",
                            "3d_synth_12x8x9"], check=True, env=env, stdout=subprocess.DEVNULL, cwd=root)
            for name in ("2d_synth_26x18_b", "3d_synth_12x8x9"):
                new = np.load(os.path.join(d, f"ref_{name}.npz"))
                old, _ = refpin.load(name)
                assert set(new.files) == set(old.files)
                for k in old.files:
                    if old[k].dtype == np.float32:
                        refpin.assert_same(new[k], old[k], f"{name} {k} (max/min ties: {mode})")


@pytest.mark.parametrize("name", refpin.fixtures("fct_"))
def test_forward_fct_oracle_equals_reference_run(name):
    """The stand-alone FCT variant (test/forward_fct.py:254-351, Kothe-Rider vortex): every stored half-step level of
    the reference run, from its level 0 and its u, v, against oracle/fct_forward_oracle.py -- bit for bit."""
    from oracle.fct_forward_oracle import FctForwardOracle
    z, meta = refpin.load(name)
    n = refpin.sizes(meta)
    o = FctForwardOracle(n["nx"], n["ny"], meta["dx"], meta["dy"], meta["dt"], meta["eps"])
    levels = [int(v) for v in z["levels"]]
    o.F[...] = z["F_levels"][levels.index(0)]
    o.u[...] = z["u"]
    o.v[...] = z["v"]
    want = dict(zip(levels, z["F_levels"]))
    checked = 0
    for t in range(n["tmax"]):
        order = (1, 0) if t % 2 == 0 else (0, 1)
        for half, axis in enumerate(order):
            o._sweep(axis)
            o.set_BC()
            lvl = 2 * t + half + 1
            if lvl in want:
                refpin.assert_same(o.F, want[lvl], f"{name} level {lvl}")
                checked += 1
        o.t += 1
    assert checked == len(levels) - 1
