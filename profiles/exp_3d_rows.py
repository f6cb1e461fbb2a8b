"""k3_jacobi5: planes per block sweep.  `python profiles/exp_3d_rows.py [n]`"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver3D, _lib, scaled_params3d
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
s = VofSolver3D(scaled_params3d(n)); s.set_init_F(1); s.run(5); s.synchronize()
for rows in (16, 32, 64, 128, 8):
    s.set_option(_lib.VOF_OPT_CHUNK_CAP, rows)
    s.solve_p_jacobi(10); s.synchronize()
    t0 = time.perf_counter(); s.solve_p_jacobi(10); s.solve_p_jacobi(10); s.synchronize(); t = (time.perf_counter() - t0) / 2
    print(f"rows {rows:4d}: 10 sweeps {t * 1e3:.3f} ms", flush=True)
