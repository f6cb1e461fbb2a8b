"""compute-sanitizer target for the round-2 kernels: the packed blocked Jacobi (bulk, edge-strip and wall-row items, exact and
tolerance mode), the packed predictor, the whole-step tile kernel (ragged last tiles), the forward-FCT variant and the
Chebyshev sweep -- small odd-sized grids, a few steps each.  `compute-sanitizer --tool memcheck|racecheck python profiles/sanitize_round2.py`"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from taichi_2d_vof_b200 import VofSolver2D, _lib, reference_params
from test_adaptive_gpu import blocky_state

os.environ.pop("VOF_TILE", None)
for (nx, ny), tile, fast in (((97, 203), 0, 0), ((300, 520), 0, 0), ((300, 520), 0, 1), ((97, 131), 2, 0), ((64, 35), 2, 0), ((200, 200), 2, 0)):
    F, u, v, p = blocky_state(nx, ny, 7)
    s = VofSolver2D(reference_params(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200))
    s.set_option(_lib.VOF_OPT_TILE, tile)
    s.set_option(_lib.VOF_OPT_JACOBI_TB, 2)            # the blocked (packed) Jacobi also on these small grids
    s.set_option(_lib.VOF_OPT_FAST_MATH, fast)
    for k, a in (("F", F), ("u", u), ("v", v), ("p", p)):
        getattr(s, k).from_numpy(a)
    for _ in range(3):
        s.step()
    s.solve_p_jacobi(7)
    s.synchronize()
    print("ok", (nx, ny), "tile" if tile else "streaming", "fast" if fast else "exact", s.mass())
s = VofSolver2D(reference_params(nx=80, ny=70, Lx=3.14, Ly=3.14, dt=1e-4))
rng = np.random.default_rng(1)
s.F.from_numpy(rng.random((82, 72), dtype=np.float32)); s.u.from_numpy(rng.random((82, 72), dtype=np.float32) * 100); s.v.from_numpy(rng.random((82, 72), dtype=np.float32) * 100)
for _ in range(4):
    s.fct_forward(1e-4)
s.set_option(_lib.VOF_OPT_PRESSURE_SOLVER, 1)
s.solve_p_jacobi(9); s.step()
s.synchronize(); print("ok extras", s.mass())
