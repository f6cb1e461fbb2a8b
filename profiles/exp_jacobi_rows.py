"""Blocked Jacobi at 8192^2: rows per work item (VOF_OPT_JACOBI_ROWS).  Event time of the two 5-sweep launches inside fused steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params
s = VofSolver2D(scaled_params(8192)); s.set_init_F(3)
for _ in range(4):
    s.step()
for rows in (0, 48, 64, 96, 128, 160, 224, 320, 0):
    s.set_option(_lib.VOF_OPT_JACOBI_ROWS, rows)
    s.step(); s.synchronize(); s.profile(True)
    for _ in range(8):
        s.step()
    s.synchronize(); pr = s.profile_read(); s.profile(False)
    print(f"rows {rows:3d}: jacobi {pr['jacobi'][0] / pr['jacobi'][1]:.4f} ms per 5-sweep launch", flush=True)
