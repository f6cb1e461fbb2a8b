"""Blocked Jacobi, third generation (packed fp32x2) against the second: 10 sweeps at n^2.
`[PRE=steps] python profiles/exp_jacobi_pk.py [n] [long_pct:short_rows ...]`
VOF_OPT_JACOBI_PK: 0 second generation, 1 third (default); VOF_OPT_JACOBI_LONG_PCT / VOF_OPT_JACOBI_ROWS: item sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, _lib, scaled_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cfgs = [tuple(int(v) for v in a.split(":")) for a in sys.argv[2:]] or [(75, 0)]
stream = torch.cuda.Stream()
s = VofSolver2D(scaled_params(n), stream=stream)
s.set_init_F(3)
s.run(int(os.environ.get("PRE", "200")))
p0 = s.p.torch().clone()


def timed(mode, pct, r):
    s.set_option(_lib.VOF_OPT_JACOBI_PK, mode)
    s.set_option(_lib.VOF_OPT_JACOBI_LONG_PCT, pct)
    s.set_option(_lib.VOF_OPT_JACOBI_ROWS, r)
    s.p.torch().copy_(p0)
    s.solve_p_jacobi(10); torch.cuda.synchronize()
    out = s.p.torch().clone()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); s.solve_p_jacobi(10); b.record(stream); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), out


t0, ref = timed(0, 75, 0)
print(f"second generation: 10 sweeps (rhs + passes + frame copy) {t0:.4f} ms")
for pct, r in cfgs:
    t1, out = timed(1, pct, r)
    same = torch.equal(ref.view(torch.int32), out.view(torch.int32))
    print(f"third generation, {pct} % long items, short items of {r or 'default'} rows: {t1:.4f} ms   bit-identical to the second: {same}")
