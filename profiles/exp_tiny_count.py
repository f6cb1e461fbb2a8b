import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, scaled_params
s = VofSolver2D(scaled_params(8192)); s.set_init_F(3); s.run(1200); s.synchronize()
p = s.p.torch()
a = p.abs()
tiny = ((a > 0) & (a < 2.0 ** -100))
print("p: zeros", int((p == 0).sum()), "tiny", int(tiny.sum()), "min nonzero", float(a[a > 0].min()), "rows with tiny", int(tiny.any(dim=1).sum()))
c = 1.0 / (s.P.dx ** 2)
print("c", c)
# numerators after one sweep-ish: b - c*(4 neighbours)
