"""A developed-flow state (PRE steps after -ic 3, default 1200) followed by a few steps: the target of the per-kernel ncu captures.
`PRE=1200 ncu --set full --import-source on --clock-control none -k regex:<kernel> -s <PRE * launches per step> -c 1 -o out python profiles/exp_one_kernel.py [n]`"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
s = VofSolver2D(scaled_params(n))
for a in sys.argv[2:]:
    k, v = a.split("=")
    s.set_option(getattr(_lib, k), int(v))
s.set_init_F(3)
s.run(int(os.environ.get("PRE", "1200")))
s.synchronize()
for _ in range(4):
    s.step()
s.synchronize()
print("mass", s.mass())
