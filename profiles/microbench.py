"""Per-entry microbenchmark on synthetic random fields (p,u*,v* ~ U(-1,1), F ~ U(0,1)):
`python profiles/microbench.py [n] [reps]`.  CUDA-event time per C-ABI entry, data-independent of any IC."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
stream = torch.cuda.Stream()
s = VofSolver2D(scaled_params(n), stream=stream)
rng = np.random.default_rng(0)
shape = (n + 2, n + 2)


def fill():
    s.F.from_numpy(rng.random(shape, dtype=np.float32))
    for k, sc in (("u", 1e-2), ("v", 1e-2), ("u_star", 1e-2), ("v_star", 1e-2), ("p", 100.0)):
        getattr(s, k).from_numpy((rng.random(shape, dtype=np.float32) - 0.5) * 2 * sc)
    s.cal_nu_rho(); s.get_normal_young()


def timeit(name, fn, cells_bytes, updates=1):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = min(ts)
    gb = cells_bytes * n * n * updates / 1e9
    print(f"{name:34s} {ms:8.3f} ms   {gb / (ms * 1e-3):8.0f} GB/s algorithmic   {n * n * updates / ms / 1e6:8.1f} Gcell/s")


fill()
timeit("get_normal_young", s.get_normal_young, 8)
s.set_option(_lib.VOF_OPT_ADVECT_COLS, 4)
timeit("advect_upwind (4 cols/lane)", s.advect_upwind, 24)
s.set_option(_lib.VOF_OPT_ADVECT_COLS, 2)
timeit("advect_upwind (2 cols/lane)", s.advect_upwind, 24)
timeit("solve_p_jacobi(1)", lambda: s.solve_p_jacobi(1), 20)
s.set_option(_lib.VOF_OPT_JACOBI_TB, 0)
timeit("solve_p_jacobi(10) no TB (+rhs)", lambda: s.solve_p_jacobi(10), 12, 10)
s.set_option(_lib.VOF_OPT_JACOBI_TB, 2)
timeit("solve_p_jacobi(10) TB 5+5 (+rhs)", lambda: s.solve_p_jacobi(10), 12, 10)
timeit("solve_p_jacobi(5) TB (+rhs)", lambda: s.solve_p_jacobi(5), 12, 5)
timeit("update_uv", s.update_uv, 24)
fill()
s.set_option(_lib.VOF_OPT_FCT_X_COLS, 4)
timeit("fct_x_sweep (4 cols/lane)", s.fct_x_sweep, 12)
s.set_option(_lib.VOF_OPT_FCT_X_COLS, 2)
timeit("fct_x_sweep (2 cols/lane)", s.fct_x_sweep, 12)
timeit("fct_y_sweep", s.fct_y_sweep, 12)
timeit("set_BC", s.set_BC, 0)
timeit("step (fused)", s.step, 216)
