import sys, os
sys.path.insert(0, os.getcwd())
from taichi_2d_vof_b200 import VofSolver2D, _lib, reference_params, scaled_params
for name, P in (("reference constants", reference_params(nx=2048, ny=2048)), ("scaled", scaled_params(2048))):
    for ic in (2, 1):
        s = VofSolver2D(P); s.set_init_F(ic); s.run(30); s.synchronize()
        s.profile(True)
        for _ in range(10): s.step()
        s.synchronize(); r = s.profile_read()
        print(name, "ic", ic, {k: round(ms / 10, 4) for k, (ms, cnt) in r.items()}, "sum", round(sum(ms for ms, _ in r.values()) / 10, 4))
        s.close()
