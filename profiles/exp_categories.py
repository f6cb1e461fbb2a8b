"""Where do the FCT / kappa kernels spend their time?  `python profiles/exp_categories.py [n] [steps]`
Runs -ic 3 for `steps` steps, classifies F (exactly 0 / exactly 1 / other) and times the entries on that state
and on synthetic all-gas / all-liquid / random states."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, scaled_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 23
stream = torch.cuda.Stream()
s = VofSolver2D(scaled_params(n), stream=stream)
s.set_init_F(3)


def timeit(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"  {name:24s} {min(ts):7.3f} ms")


def classify(tag):
    F = s.F.to_numpy()[1:-1, 1:-1]
    z, o = (F == 0).mean(), (F == 1).mean()
    rows0 = (F == 0).all(axis=1).mean()
    # 64-column x 1-row segments (the unit a warp of the x sweep sees) that are uniform
    seg = F[:, : (F.shape[1] // 64) * 64].reshape(F.shape[0], -1, 64)
    seg0 = (seg == 0).all(axis=2).mean(); seg1 = (seg == 1).all(axis=2).mean()
    near1 = ((F != 1) & (np.abs(F - 1) < 1e-5)).mean()
    print(f"{tag}: F==0 {z:.4f}  F==1 {o:.4f}  other {1 - z - o:.4f} (of which |F-1|<1e-5: {near1:.4f});"
          f" 64-col segments all-0 {seg0:.4f} all-1 {seg1:.4f}; rows all-0 {rows0:.4f}")
    u = s.u.to_numpy(); v = s.v.to_numpy()
    print(f"   u==0 {(u == 0).mean():.4f}  v==0 {(v == 0).mean():.4f}  max|u| {np.abs(u).max():.3e} max|v| {np.abs(v).max():.3e}")


classify("step 0")
for k in range(steps):
    s.step()
    if k + 1 in (1, 3, 10, steps):
        classify(f"step {k + 1}")
from taichi_2d_vof_b200 import _lib
state = {k: getattr(s, k).to_numpy() for k in ("F", "u", "v")}
MODES = ((1, 2), (0, 2))
for adaptive, cols in MODES:
    s.set_option(_lib.VOF_OPT_ADAPTIVE, adaptive); s.set_option(_lib.VOF_OPT_FCT_X_COLS, cols)
    print(f"timings on the -ic 3 state, adaptive={adaptive} fct_x cols={cols}:")
    for nm in ("get_normal_young", "fct_x_sweep", "fct_y_sweep"):
        for k, a in state.items():
            getattr(s, k).from_numpy(a)
        timeit(nm, getattr(s, nm), reps=1)
rng = np.random.default_rng(0)
shape = (n + 2, n + 2)
uv = (rng.random(shape, dtype=np.float32) - 0.5) * 2e-2
for tag, Fv in (("all gas F=0", np.zeros(shape, np.float32)), ("all liquid F=1", np.ones(shape, np.float32)),
                ("random F", rng.random(shape, dtype=np.float32))):
    for adaptive, cols in MODES:
        s.set_option(_lib.VOF_OPT_ADAPTIVE, adaptive); s.set_option(_lib.VOF_OPT_FCT_X_COLS, cols)
        print(f"timings on {tag}, random u, v (|u| < 1e-2), adaptive={adaptive} fct_x cols={cols}:")
        for nm in ("get_normal_young", "fct_x_sweep", "fct_y_sweep"):
            s.F.from_numpy(Fv); s.u.from_numpy(uv); s.v.from_numpy(uv)
            timeit(nm, getattr(s, nm), reps=1)
