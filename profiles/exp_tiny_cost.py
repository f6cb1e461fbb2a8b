"""What the sub-normal fix-up of the exact division costs the packed Jacobi: 10 sweeps on the developed state as it is and on the
same state with the pressure shifted by a constant (no tiny numerators any more; a constant shift changes no pressure gradient)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, scaled_params

stream = torch.cuda.Stream()
s = VofSolver2D(scaled_params(8192), stream=stream)
s.set_init_F(3); s.run(int(os.environ.get("PRE", "1200"))); s.synchronize()
p0 = s.p.torch().clone()
for label, shift in (("as is", 0.0), ("p + 1000", 1000.0)):
    s.p.torch().copy_(p0 + shift)
    s.solve_p_jacobi(10); torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        s.p.torch().copy_(p0 + shift)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); s.solve_p_jacobi(10); b.record(stream); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(f"{label}: 10 sweeps (rhs + 2 passes + frame copy) {min(ts):.4f} ms")
