"""Generates the 3-D golden vectors under tests/golden/ from the NumPy oracle (oracle/vof3d_oracle.py).

As for the 2-D vectors these are the ORACLE's own output (regression fixtures); the pin to the reference is
tests/golden/ref_3d_*.npz (oracle/run_reference.py: the reference's own text under the taichi stand-in).  Re-run:
`python tests/golden/make_golden3d.py`.  Files: vof3d_ic1_{nx}x{ny}x{nz}.npz with u,v,w,p,F after 1, 3 and 12
steps (3 = one rotation of the x/y/z sweep order, 3dvof.py:351-363) + interior volumes.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.vof3d_oracle import Vof3DOracle, Vof3DParams  # noqa: E402

CHECKPOINTS = (1, 3, 12)
FIELDS = ("u", "v", "w", "p", "F")
CASES = [(24, 20, 28), (12, 34, 18)]


def main():
    for nx, ny, nz in CASES:
        P = Vof3DParams(nx=nx, ny=ny, nz=nz, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200, Lz=0.1 * nz / 200)
        o = Vof3DOracle(P)
        o.set_init_F(1)
        out = {"F_init": o.F.copy(), "mass_init": o.mass(), "params": np.array([nx, ny, nz, P.Lx, P.Ly, P.Lz])}
        for ck in CHECKPOINTS:
            o.run(ck - o.istep)
            for k in FIELDS:
                out[f"{k}_{ck}"] = getattr(o, k).copy()
            out[f"mass_{ck}"] = o.mass()
        path = os.path.join(HERE, f"vof3d_ic1_{nx}x{ny}x{nz}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
