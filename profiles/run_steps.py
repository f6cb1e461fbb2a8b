"""Tiny driver for ncu captures: `python profiles/run_steps.py [n] [steps] [ic]` runs the fused step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, scaled_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ic = int(sys.argv[3]) if len(sys.argv) > 3 else 3
s = VofSolver2D(scaled_params(n))
s.set_init_F(ic)
for _ in range(steps):
    s.step()
s.synchronize()
print("mass", s.mass())
