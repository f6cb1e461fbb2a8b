"""Whole-step tile kernel (VOF_OPT_TILE = 2) against the streaming kernels (0) by grid size: graph-replayed steps/s, -ic 1.
`python profiles/exp_tile.py [n ...]`"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params

sizes = [int(a) for a in sys.argv[1:]] or [200, 256, 384, 512, 640, 768, 1024]
for n in sizes:
    out = []
    for tile in (0, 2):
        s = VofSolver2D(scaled_params(n))
        s.set_option(_lib.VOF_OPT_TILE, tile)
        s.set_init_F(1)
        s.run(200); s.synchronize()
        k = 2000 if n <= 512 else 800
        t0 = time.perf_counter(); s.run(k); s.synchronize(); dt = time.perf_counter() - t0
        out.append(k / dt)
        s.close()
    print(f"{n}^2: streaming {out[0]:9.0f} steps/s   tile {out[1]:9.0f} steps/s   ({out[1] / out[0]:.2f}x)")
