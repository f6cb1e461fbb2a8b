"""Runs the reference's own scripts under the taichi stand-in -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

``python oracle/run_reference.py <case> [...]`` executes the UNMODIFIED text of
``/root/reference/{2dvof.py,3dvof.py,test/forward_fct.py}`` with ``oracle/refshim/taichi`` standing in for
``taichi==1.4.1`` (not installable here) and writes the fields it produced to ``tests/golden/ref_*.npz``.
Those fixtures are what pins the hand-written oracles and the CUDA path to the reference itself
(tests/test_reference_pin_cpu.py, tests/test_reference_pin_gpu.py).

The only edits ever made to the text are the size constants (``nx = 200`` -> ``nx = N``, likewise
``ny``, ``nz`` and, for test/forward_fct.py, ``tmax``) -- the mechanical substitution SURVEY.md 8c allows;
each is a whole-line regex that must match exactly once, and the fixture records which were applied.
The default-size cases (``--nx 200 --ny 200``) run the text with no substitution at all.

Inputs other than the built-in initial conditions are injected as DATA, never as code: with
``--inject SEED`` the arrays u, v, p, F are overwritten with a seeded synthetic state right after the
script's own ``set_init_F`` returned (so that every branch of the upwind / limiter logic is exercised
in a few steps; the natural initial conditions start from rest).

This needs ``/root/reference`` and therefore only runs in the build container; the GPU box uses the
committed fixtures.
"""
from __future__ import annotations

import argparse
import linecache
import os
import re
import sys
import tempfile
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
GOLDEN = os.environ.get("REF_OUT_DIR") or os.path.join(os.path.dirname(HERE), "tests", "golden")


def _install_stubs():
    sys.path.insert(0, os.path.join(HERE, "refshim"))
    import taichi  # noqa: F401  (the stand-in)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Any:
        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **kw):
            return _Any()

    plt = mod("matplotlib.pyplot", __getattr__=lambda k: _Any())
    cm = mod("matplotlib.cm", __getattr__=lambda k: _Any())
    mod("matplotlib", pyplot=plt, cm=cm)
    vtk_calls = []
    hl = mod("pyevtk.hl", gridToVTK=lambda path, *a, **kw: vtk_calls.append((path, kw)))
    mod("pyevtk", hl=hl)
    mod("flow_visualization", plot_arrow_field=lambda **kw: None, plot_vector_field=lambda **kw: None)
    return taichi, vtk_calls


def synthetic_state(shape, seed, vel):
    """Seeded input state for --inject (also used by the tests to feed the oracles / the CUDA path).

    F: blocky two-phase field with a smeared band (exact 0/1 bulk + fractional cells), u/v/w: uniform in
    (-vel, vel) with exact zeros sprinkled in, p: uniform in (-1, 1).  All fp32, ghosts included.
    """
    rng = np.random.default_rng(seed)
    nd = len(shape)
    F = (rng.random(shape) < 0.5).astype(np.float32)
    frac = rng.random(shape) < 0.4
    F[frac] = rng.random(int(frac.sum())).astype(np.float32)
    out = {"F": F}
    for name in ("u", "v", "w")[:nd]:
        a = ((rng.random(shape) * 2 - 1) * vel).astype(np.float32)
        a[rng.random(shape) < 0.1] = 0.0
        out[name] = a
    out["p"] = (rng.random(shape) * 2 - 1).astype(np.float32)
    return out


def run(script, ic, sizes, steps, kernel_steps, inject=None, vel=None, verbose=True):
    """Returns (dict of arrays, meta dict)."""
    ti, vtk_calls = _install_stubs()
    path = os.path.join(REF, script)
    text = open(path).read()
    subs = []
    for name, val in sizes.items():
        pat = re.compile(rf"^{name} = \d+\b", re.M)
        hits = pat.findall(text)
        assert len(hits) == 1, (name, hits)
        if hits[0] != f"{name} = {val}":
            text = pat.sub(f"{name} = {val}", text)
            subs.append(f"{hits[0]} -> {name} = {val}")
    fname = path if not subs else f"<reference {script} " + " ".join(f"{k}={v}" for k, v in sizes.items()) + ">"
    if subs:
        lines = text.splitlines(True)
        linecache.cache[fname] = (len(text), None, lines, fname)

    out = {}
    g = {"__name__": "__main__", "__file__": path}
    state = {"step": 0, "call": 0}
    is3d = script == "3dvof.py"

    big = max(sizes.values()) >= 100      # large grids: only the live state, not the 19 scratch arrays
    keep = ("F", "u", "v", "w", "p", "rho", "nu", "kappa", "u_star", "v_star", "w_star")

    def grid_fields():
        full = tuple(sizes[k] + 2 for k in (("nx", "ny", "nz") if is3d else ("nx", "ny")))
        return {k: v for k, v in g.items() if isinstance(v, ti.Field) and not v.vector and v.a.shape == full
                and (not big or k in keep)}

    last = {}

    def after_kernel(name):
        if name == "set_init_F":
            out["F_init"] = g["F"].a.copy()
            if inject is not None:
                st = synthetic_state(g["F"].a.shape, inject, vel)
                for k, a in st.items():
                    g[k].a[...] = a
                    out[f"{k}_in"] = a
            last.update({k: f.a.copy() for k, f in grid_fields().items()})
            return
        step = state["step"] + 1
        if step <= kernel_steps:                # per-call record: the fields this call changed (bitwise)
            c = state["call"]
            out[f"call{c:04d}_name"] = np.array(name)
            for k, f in grid_fields().items():
                prev = last.get(k)
                if prev is None or not np.array_equal(prev.view(np.uint32), f.a.view(np.uint32)):
                    last[k] = out[f"call{c:04d}_{k}"] = f.a.copy()
            state["call"] = c + 1

    t0 = time.time()

    def on_step(done):
        state["step"] = done
        if done in steps:
            for k, f in grid_fields().items():
                out[f"{k}_{done}"] = f.a.copy()
        if verbose:
            print(f"  step {done} done, {time.time() - t0:.1f} s", flush=True)

    ti._hooks["after_kernel"] = after_kernel
    ti.GUI.budget = max(steps) if steps else 0
    ti.GUI.on_step = on_step
    argv, cwd = sys.argv, os.getcwd()
    sys.argv = [script, "-ic", str(ic)]
    tmp = tempfile.mkdtemp(prefix="refrun_")
    os.chdir(tmp)
    old = np.seterr(all="ignore")
    try:
        exec(compile(text, fname, "exec"), g)
    finally:
        np.seterr(**old)
        os.chdir(cwd)
        sys.argv = argv
        ti._hooks["after_kernel"] = None
        ti.GUI.on_step = None
    meta = {"script": script, "ic": ic, "substitutions": subs, "steps": sorted(steps),
            "kernel_steps": kernel_steps, "inject": inject, "vel": vel,
            "maxmin": os.environ.get("TI_SHIM_MAXMIN", "first"),
            "dx": g.get("dx"), "dy": g.get("dy"), "dz": g.get("dz"), "dt": g.get("dt"),
            "n_calls": state["call"], "vtk_calls": [p for p, _ in vtk_calls]}
    if script.endswith("forward_fct.py"):          # F is a history of 2 tmax + 1 half-step levels
        hist = g["F"].a
        levels = sorted(set(list(range(0, 9)) + list(range(0, hist.shape[0], 32)) + [hist.shape[0] - 1]))
        out["levels"] = np.array(levels)
        out["F_levels"] = hist[levels].copy()
        for k in ("u", "v", "Ftarget"):
            out[k] = g[k].a.copy()
        meta["eps"] = 1.0e-4                        # forward(eps_value=1.0e-4), test/forward_fct.py:382
    return out, meta


CASES = {
    # name: (script, ic, sizes, steps, kernel_steps, inject seed, velocity scale)
    # 2-D, the reference's own configuration: NO substitution at all
    "2d_ic1_200x200": ("2dvof.py", 1, dict(nx=200, ny=200), (1, 2, 3, 10, 100), 0, None, None),
    "2d_ic2_200x200": ("2dvof.py", 2, dict(nx=200, ny=200), (1, 2, 3, 10, 100), 0, None, None),
    "2d_ic3_200x200": ("2dvof.py", 3, dict(nx=200, ny=200), (1, 2, 3, 10, 100), 0, None, None),
    # small grids (nx != ny, so dx != dy), every kernel call of the first steps kept
    "2d_ic1_24x20": ("2dvof.py", 1, dict(nx=24, ny=20), (1, 2, 3, 10), 2, None, None),
    "2d_ic2_20x28": ("2dvof.py", 2, dict(nx=20, ny=28), (1, 2, 3, 10), 2, None, None),
    "2d_ic3_16x36": ("2dvof.py", 3, dict(nx=16, ny=36), (1, 2, 3, 10), 2, None, None),
    # injected synthetic states: CFL ~ 0.1 and ~ 0.3 (dx/dt ~ 1e3 m/s on these grids)
    "2d_synth_22x26_a": ("2dvof.py", 1, dict(nx=22, ny=26), (1, 2, 3, 4), 2, 11, 100.0),
    "2d_synth_26x18_b": ("2dvof.py", 3, dict(nx=26, ny=18), (1, 2, 3, 4), 2, 12, 300.0),
    "2d_synth_40x40_c": ("2dvof.py", 2, dict(nx=40, ny=40), (1, 2, 6), 1, 13, 50.0),
    # 3-D
    "3d_ic1_10x12x8": ("3dvof.py", 1, dict(nx=10, ny=12, nz=8), (1, 2, 3, 6), 2, None, None),
    "3d_synth_8x10x12": ("3dvof.py", 1, dict(nx=8, ny=10, nz=12), (1, 2, 3, 4), 2, 21, 100.0),
    "3d_synth_12x8x9": ("3dvof.py", 1, dict(nx=12, ny=8, nz=9), (1, 2, 3), 1, 22, 300.0),
    # the stand-alone FCT variant, Kothe-Rider vortex; tmax is substituted too
    # (max CFL = 2 nx / tmax, so tmax >= 4 nx keeps it <= 0.5 as in the original 500^2 / 1000 steps ~ 1)
    "fct_40x40_t160": ("test/forward_fct.py", 0, dict(nx=40, ny=40, tmax=160), (), 0, None, None),
    "fct_32x48_t200": ("test/forward_fct.py", 0, dict(nx=32, ny=48, tmax=200), (), 0, None, None),
}


def save(name, out, meta):
    path = os.path.join(GOLDEN, f"ref_{name}.npz")
    out = dict(out)
    out["meta"] = np.array(repr(meta))
    np.savez_compressed(path, **out)
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", help="case names (default: all but the 200x200 ones)")
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--steps", type=str, default=None, help="override the snapshot steps, e.g. 1,2,3,10")
    ap.add_argument("--no-save", action="store_true")
    a = ap.parse_args()
    if a.list:
        for k, v in CASES.items():
            print(k, v)
        return
    names = a.cases or [k for k in CASES if "200x200" not in k]
    for name in names:
        script, ic, sizes, steps, ksteps, inject, vel = CASES[name]
        if a.steps:
            steps = tuple(int(s) for s in a.steps.split(","))
        t0 = time.time()
        print(f"[{name}] {script} -ic {ic} {sizes} steps {steps}", flush=True)
        out, meta = run(script, ic, sizes, set(steps), ksteps, inject, vel)
        if not a.no_save:
            path = save(name, out, meta)
            print(f"[{name}] {time.time() - t0:.1f} s -> {path} ({os.path.getsize(path) // 1024} KiB)", flush=True)


if __name__ == "__main__":
    main()
