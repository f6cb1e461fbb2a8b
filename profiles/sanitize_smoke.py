import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from taichi_2d_vof_b200 import VofSolver2D, VofSolver3D, reference_params, reference_params3d, _lib
from test_adaptive_gpu import blocky_state
# 2-D: blocky state, odd sizes, whole steps + single entries
for (nx, ny) in ((97, 203), (64, 129), (40, 61)):
    F, u, v, p = blocky_state(nx, ny, 7)
    s = VofSolver2D(reference_params(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200))
    for k, a in (("F", F), ("u", u), ("v", v), ("p", p)):
        getattr(s, k).from_numpy(a)
    for _ in range(3):
        s.step()
    s.get_normal_young(); s.fct_x_sweep(); s.fct_y_sweep(); s.advect_upwind(); s.get_vof_field(); s.interp_velocity()
    s.synchronize(); print("2d ok", nx, ny, s.mass())
# slab context (halo rows)
s = VofSolver2D(reference_params(nx=128, ny=96, slab=(33, 64), halo=16)); s.set_init_F(3)
for _ in range(2):
    s.step()
s.synchronize(); print("slab ok")
# 3-D
for shape in ((12, 9, 21), (8, 12, 128), (10, 10, 70)):
    nx, ny, nz = shape
    s3 = VofSolver3D(reference_params3d(nx=nx, ny=ny, nz=nz, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200, Lz=0.1 * nz / 200)); s3.set_init_F(1)
    for _ in range(4):
        s3.step()
    s3.synchronize(); print("3d ok", shape, s3.mass())
