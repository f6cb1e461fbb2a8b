"""`python profiles/run_steps3d.py [n] [steps]`: fused 3-D steps (dam break) for ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver3D, scaled_params3d

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
s = VofSolver3D(scaled_params3d(n))
s.set_init_F(1)
for _ in range(steps):
    s.step()
s.synchronize()
print("mass", s.mass())
