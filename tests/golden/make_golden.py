"""Generates the golden vectors under tests/golden/ from the NumPy oracle (oracle/vof2d_oracle.py).

These vectors are the ORACLE's own output (regression fixtures at sizes / step counts the reference-run fixtures do not
cover); what pins the oracle to the reference is tests/golden/ref_*.npz, written by oracle/run_reference.py from the
reference's own source text executed under the taichi stand-in (tests/test_reference_pin_cpu.py).  Re-run: `python tests/golden/make_golden.py`.
Files: vof2d_ic{1,2,3}_{nx}x{ny}.npz with u,v,p,F,kappa after 1, 10 and 100 steps + interior volumes.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.vof2d_oracle import Vof2DOracle, Vof2DParams  # noqa: E402

CHECKPOINTS = (1, 10, 100)
FIELDS = ("u", "v", "p", "F", "kappa")
CASES = [(200, 200, 0.1, 0.1), (64, 96, 0.032, 0.048)]   # the reference default + a small non-square grid


def main():
    for nx, ny, Lx, Ly in CASES:
        for ic in (1, 2, 3):
            o = Vof2DOracle(Vof2DParams(nx=nx, ny=ny, Lx=Lx, Ly=Ly))
            o.set_init_F(ic)
            out = {"F_init": o.F.copy(), "mass_init": o.mass(), "params": np.array([nx, ny, Lx, Ly, o.P.dx, o.P.dy, o.P.dt])}
            for ck in CHECKPOINTS:
                o.run(ck - o.istep)
                for k in FIELDS:
                    out[f"{k}_{ck}"] = getattr(o, k).copy()
                out[f"mass_{ck}"] = o.mass()
            path = os.path.join(HERE, f"vof2d_ic{ic}_{nx}x{ny}.npz")
            np.savez_compressed(path, **out)
            print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
