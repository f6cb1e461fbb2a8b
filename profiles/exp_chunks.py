"""Load-balance experiment: kappa / fct_x / fct_y on the -ic 3 state at 8192^2 versus the rows-per-item cap."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params

n = 8192
stream = torch.cuda.Stream()
s = VofSolver2D(scaled_params(n), stream=stream)
s.set_init_F(3)
for _ in range(23):
    s.step()
state = {k: getattr(s, k).to_numpy() for k in ("F", "u", "v")}


def t1(fn, reps=3):
    best = 1e9
    for _ in range(reps + 1):
        for k, a in state.items():
            getattr(s, k).from_numpy(a)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for adaptive, cols, caps in ((0, 2, (0,)), (1, 2, (0, 64, 48, 32, 24, 16)), (1, 4, (0, 48))):
    s.set_option(_lib.VOF_OPT_ADAPTIVE, adaptive); s.set_option(_lib.VOF_OPT_FCT_X_COLS, cols)
    for cap in caps:
        s.set_option(_lib.VOF_OPT_CHUNK_CAP, cap)
        print(f"adaptive {adaptive} cols {cols} cap {cap:3d}: kappa {t1(s.get_normal_young):.3f}  fct_x {t1(s.fct_x_sweep):.3f}  "
              f"fct_y {t1(s.fct_y_sweep):.3f} ms", flush=True)
