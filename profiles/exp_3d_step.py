import sys, time
sys.path.insert(0, "/root/repo")
from taichi_2d_vof_b200 import VofSolver3D, scaled_params3d
s = VofSolver3D(scaled_params3d(512)); s.set_init_F(1); s.run(30); s.synchronize()
t0 = time.perf_counter(); s.run(40); s.synchronize(); dt = time.perf_counter() - t0
print("3-D 512^3 ms/step", dt / 40 * 1e3)
