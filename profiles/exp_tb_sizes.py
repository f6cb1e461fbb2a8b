"""Where does the blocked Jacobi start to pay, and with how many sweeps per pass?  Graph-replayed steps/s with
VOF_OPT_JACOBI_TB = 0 (single-sweep kernel) / 2 (blocked kernel, VOF_OPT_JACOBI_MAXT sweeps per pass at most)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, _lib, reference_params
for n, ic in ((512, 1), (1024, 1), (2048, 2), (3072, 2), (4096, 3)):
    for tb, maxt in ((0, 5), (2, 1), (2, 2), (2, 3), (2, 5)):
        s = VofSolver2D(reference_params(nx=n, ny=n, Lx=0.1 * n / 200, Ly=0.1 * n / 200))
        s.set_option(_lib.VOF_OPT_JACOBI_TB, tb); s.set_option(_lib.VOF_OPT_JACOBI_MAXT, maxt); s.set_init_F(ic)
        k = 1000 if n <= 1024 else 200
        s.run(50); s.synchronize()
        t0 = time.perf_counter(); s.run(k); s.synchronize(); t = (time.perf_counter() - t0) / k
        print(f"n {n} jacobi_tb {tb} max sweeps/pass {maxt}: {1 / t:.0f} steps/s ({t * 1e6:.1f} us/step)", flush=True)
        del s
