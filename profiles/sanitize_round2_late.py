"""compute-sanitizer target for the last changes of round 2: the tile kernel with dependency-cone margins (geometries where the
far ghost row / column is absorbed, ragged tiles, odd n_jacobi) and the streamer's staging kernel (in place and out of place).
`compute-sanitizer --tool memcheck|racecheck python profiles/sanitize_round2_late.py`"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from taichi_2d_vof_b200 import VofSolver2D, VofStreamer2D, _lib, reference_params
from test_adaptive_gpu import blocky_state

for (nx, ny), nj in (((63, 67), 10), ((50, 101), 10), ((97, 131), 10), ((70, 70), 1), ((33, 140), 13), ((200, 200), 10)):
    F, u, v, p = blocky_state(nx, ny, 7)
    s = VofSolver2D(reference_params(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200, n_jacobi=nj))
    s.set_option(_lib.VOF_OPT_TILE, 2)
    for k, a in (("F", F), ("u", u), ("v", v), ("p", p)):
        getattr(s, k).from_numpy(a)
    for _ in range(3):
        s.step()
    s.synchronize()
    print("ok tile", (nx, ny), nj, s.mass(), flush=True)
nx, ny = 130, 75
F, u, v, p = blocky_state(nx, ny, 3)
st = VofStreamer2D(reference_params(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200), n_slabs=5)
a = [u.copy(), v.copy(), p.copy(), F.copy()]
b = [np.empty_like(x) for x in a]
st.step_host(*a)                 # in place
st.step_host(*a, out=b)          # out of place
st.step_host(*b)
st.close()
print("ok streamer", float(b[3].sum()), flush=True)
