"""Per-kernel times of the fused step on a developed flow (PRE steps after -ic 3), CUDA events around every kernel of 12 steps.
`[PRE=1200] [env knobs] python profiles/exp_step_kernels.py [n] [VOF_OPT_NAME=value ...]`"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
s = VofSolver2D(scaled_params(n))
for a in sys.argv[2:]:
    k, v = a.split("=")
    s.set_option(getattr(_lib, k), int(v))
s.set_init_F(3)
s.run(int(os.environ.get("PRE", "1200")))
for _ in range(3):
    s.step()
s.synchronize()
s.profile(True)
for _ in range(12):
    s.step()
s.synchronize()
r = s.profile_read()
tot = sum(ms for ms, _ in r.values()) / 12
print(" ".join(sys.argv[2:]) or "defaults", {k: round(ms / 12, 4) for k, (ms, cnt) in r.items()}, "sum", round(tot, 4))
