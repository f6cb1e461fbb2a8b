"""A/B of the whole fused step at 8192^2 (-ic 3): per-kernel event times for adaptive on / off.  `python profiles/exp_step.py [n] [steps]`"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
caps = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
for adaptive in (1, 0):
    for cap in (caps if adaptive else [0]):
        s = VofSolver2D(scaled_params(n)); s.set_option(_lib.VOF_OPT_ADAPTIVE, adaptive); s.set_option(_lib.VOF_OPT_CHUNK_CAP, cap)
        s.set_init_F(3)
        for _ in range(3):
            s.step()
        s.synchronize()
        s.profile(True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        for _ in range(steps):
            s.step()
        s.synchronize()
        dt = (time.perf_counter() - t0) / steps * 1e3
        pr = s.profile_read()
        print(f"adaptive {adaptive} cap {cap}: {dt:.3f} ms/step (with profiling events)", {k: round(v[0] / max(1, v[1]), 3) for k, v in pr.items()}, flush=True)
        del s
