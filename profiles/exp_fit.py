"""Item sizes fitted to whole rounds of the resident warps (VOF_OPT_FIT_ROUNDS) on / off: per-kernel event times of the
fused step at 8192^2 (-ic 3) and graph-replayed steps/s at a few sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, _lib, reference_params, scaled_params
for fit in (1, 0, 1, 0):
    s = VofSolver2D(scaled_params(8192)); s.set_option(_lib.VOF_OPT_FIT_ROUNDS, fit); s.set_init_F(3)
    for _ in range(3):
        s.step()
    s.synchronize(); s.profile(True)
    for _ in range(20):
        s.step()
    s.synchronize(); pr = s.profile_read(); s.profile(False)
    t0 = time.perf_counter(); s.run(40); s.synchronize(); t = (time.perf_counter() - t0) / 40
    print(f"fit {fit}: graph replay {t * 1e3:.4f} ms/step", {k: round(v[0] / max(1, v[1]), 4) for k, v in pr.items()}, flush=True)
    del s
for n, ic in ((2048, 2), (4096, 3)):
    for fit in (1, 0):
        s = VofSolver2D(reference_params(nx=n, ny=n, Lx=0.1 * n / 200, Ly=0.1 * n / 200)); s.set_option(_lib.VOF_OPT_FIT_ROUNDS, fit); s.set_init_F(ic)
        s.run(50); s.synchronize()
        t0 = time.perf_counter(); s.run(200); s.synchronize(); t = (time.perf_counter() - t0) / 200
        print(f"n {n} fit {fit}: {1 / t:.0f} steps/s", flush=True)
        del s
