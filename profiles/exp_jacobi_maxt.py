"""Blocked Jacobi: 5 + 5 sweeps per pass against 4 + 3 + 3 (narrower strip margins).  `python profiles/exp_jacobi_maxt.py`"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params
stream = torch.cuda.Stream()
s = VofSolver2D(scaled_params(8192), stream=stream); s.set_init_F(3)
for _ in range(10):
    s.step()
s.synchronize()
for maxt in (5, 4, 3, 5):
    s.set_option(_lib.VOF_OPT_JACOBI_MAXT, maxt)
    s.solve_p_jacobi(10); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); s.solve_p_jacobi(10); b.record(stream); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"max sweeps per pass {maxt}: solve_p_jacobi(10) incl. rhs {best:.3f} ms", flush=True)
