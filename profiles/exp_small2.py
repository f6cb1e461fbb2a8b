"""Grid-size sweep: graph-replayed steps/s with the default item sizing and with caps on the rows per item."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, _lib, reference_params
for n, ic in ((200, 1), (512, 1), (1024, 1), (2048, 2), (4096, 3), (8192, 3)):
    for cap in ((0, 8) if n < 8192 else (0, 32, 16)):
        s = VofSolver2D(reference_params(nx=n, ny=n, Lx=0.1 * n / 200, Ly=0.1 * n / 200)); s.set_option(_lib.VOF_OPT_CHUNK_CAP, cap); s.set_init_F(ic)
        k = 1000 if n <= 1024 else (200 if n <= 4096 else 40)
        s.run(50 if n < 8192 else 6); s.synchronize()
        t0 = time.perf_counter(); s.run(k); s.synchronize(); t = (time.perf_counter() - t0) / k
        print(f"n {n} cap {cap:2d}: {1 / t:.0f} steps/s ({t * 1e6:.1f} us/step)", flush=True)
        del s
