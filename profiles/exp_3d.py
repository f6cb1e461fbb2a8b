"""3-D A/B: `python profiles/exp_3d.py [n] [steps]` -- ms/step of the fused 3-D step, kernel generations 1 and 0, plus 10 Jacobi sweeps alone."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver3D, _lib, scaled_params3d

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
for gen2, smem in ((1, 1), (1, 0), (0, 0)):
    s = VofSolver3D(scaled_params3d(n)); s.set_option(_lib.VOF_OPT_ADAPTIVE, gen2); s.set_option(_lib.VOF_OPT_JACOBI_TB, smem); s.set_init_F(1)
    s.run(3); s.synchronize()
    t0 = time.perf_counter(); s.run(steps); s.synchronize(); t = (time.perf_counter() - t0) / steps
    t0 = time.perf_counter(); s.solve_p_jacobi(10); s.synchronize(); tj = time.perf_counter() - t0
    tt = {}
    for nm in ("fct_x_sweep", "fct_y_sweep", "fct_z_sweep", "advect_upwind", "update_uv"):
        t0 = time.perf_counter(); getattr(s, nm)(); s.synchronize(); tt[nm] = round((time.perf_counter() - t0) * 1e3, 3)
    print(f"gen2={gen2} jacobi-smem={smem}: {t * 1e3:.3f} ms/step = {1 / t:.1f} steps/s, {10 * n ** 3 / t / 1e9:.1f} Gcell-upd/s; 10 sweeps {tj * 1e3:.3f} ms; {tt}", flush=True)
    del s
