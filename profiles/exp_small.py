"""Small grids (config 1, 200^2 ...): graph-replayed steps/s against the rows-per-item cap (VOF_OPT_CHUNK_CAP) and the
kernel generation.  Small grids are latency bound: ~19 dependent kernels per step, each a few serial row marches."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, _lib, reference_params
for n in (200, 512, 1024):
    for ad, cap in ((1, 0), (1, 16), (1, 12), (1, 8), (1, 4), (0, 0), (0, 8)):
        s = VofSolver2D(reference_params(nx=n, ny=n, Lx=0.1 * n / 200, Ly=0.1 * n / 200))
        s.set_option(_lib.VOF_OPT_ADAPTIVE, ad); s.set_option(_lib.VOF_OPT_CHUNK_CAP, cap); s.set_init_F(1)
        s.run(200); s.synchronize()
        t0 = time.perf_counter(); s.run(2000); s.synchronize(); t = (time.perf_counter() - t0) / 2000
        print(f"n {n} adaptive {ad} cap {cap:2d}: {1 / t:.0f} steps/s ({t * 1e6:.1f} us/step)", flush=True)
        del s
