"""Helpers for the reference-pin tests: read the ``tests/golden/ref_*.npz`` fixtures that
``oracle/run_reference.py`` produced by EXECUTING the reference's own source text (2dvof.py, 3dvof.py,
test/forward_fct.py) under the taichi stand-in, and replay them against an implementation."""
import ast
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixtures(prefix):
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, f"ref_{prefix}*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    meta = ast.literal_eval(str(z["meta"]))
    return z, meta


def sizes(meta):
    """Grid sizes of a fixture: substituted values, else the reference's defaults."""
    out = {"nx": 200, "ny": 200, "nz": 200, "tmax": 1000}
    if meta["script"].endswith("forward_fct.py"):
        out.update(nx=500, ny=500)
    for s in meta["substitutions"]:
        k, v = s.split("->")[1].split("=")
        out[k.strip()] = int(v)
    return out


def bits_differ(a, b):
    """Indices where two fp32 arrays differ bitwise (sign of zero and NaN payloads included)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert a.shape == b.shape, (a.shape, b.shape)
    return np.argwhere(a.view(np.uint32) != b.view(np.uint32))


def assert_same(a, b, tag):
    bad = bits_differ(a, b)
    if bad.size:
        i = tuple(bad[0])
        ua, ub = int(np.asarray(a, np.float32)[i].view(np.uint32)), int(np.asarray(b, np.float32)[i].view(np.uint32))
        raise AssertionError(f"{tag}: {len(bad)} elements differ from the reference run, first at {i}: "
                             f"{a[i]!r} vs {b[i]!r} ({abs(ua - ub)} ulp)")


def calls(z, meta):
    """Yields (kernel name, {field: array after the call}) with the delta encoding undone."""
    names = {k.split("_", 1)[1] for k in z.files if k.startswith("call") and not k.endswith("_name")}
    state = {n: np.zeros_like(z["F_init"]) for n in names}       # every field starts at zero (2dvof.py:53-89)
    state["F"] = z["F_init"]
    for k in ("F", "u", "v", "w", "p"):
        if k + "_in" in z.files:                                  # injected synthetic input state
            state[k] = z[k + "_in"]
    for c in range(meta["n_calls"]):
        pre = f"call{c:04d}_"
        for k in z.files:
            if k.startswith(pre) and k != pre + "name":
                state[k[len(pre):]] = z[k]
        yield str(z[pre + "name"]), dict(state)
