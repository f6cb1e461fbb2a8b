"""Drop-in for the main loop of the reference's 3-D script (`python 3dvof.py [-ic {1,2,3}] [-s]`).

Keeps the command line (3dvof.py:12-18; `-s` is parsed and ignored there too), the start-up banner (119-123), the
constants (20-38), the kernel call sequence (606-623) and the export cadence (every nstep = 100 steps: a
RectilinearGrid `output/step-%05d.vtr` with the point data "VOF" on the unit-cube coordinates, 60-62, 624-627), and
runs the kernels on a B200 through libvof (`vof3d_*`).  Extensions whose defaults reproduce the reference: `--steps`
(the GUI loop never ends by itself), `--nx/--ny/--nz`, `--scaled`, `--no-export`, `--sequence`, `--dump`.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="B200-native 3-D VOF solver (drop-in for 3dvof.py's loop)")
    p.add_argument('-ic', type=int, choices=[1, 2, 3], default=1)      # 3dvof.py:14 (only 1 sets F: 126-138)
    p.add_argument('-s', action='store_true')                          # 3dvof.py:15 (unused by the reference's loop)
    p.add_argument('--nx', type=int, default=200)
    p.add_argument('--ny', type=int, default=200)
    p.add_argument('--nz', type=int, default=200)
    p.add_argument('--scaled', action='store_true', help="constant-dx scaling: L = 0.1 * n / 200 per axis")
    p.add_argument('--dt', type=float, default=4e-6)
    p.add_argument('--jacobi', type=int, default=10)
    p.add_argument('--steps', type=int, default=0, help="stop after this many steps (0 = run until interrupted)")
    p.add_argument('--nstep', type=int, default=100, help="export interval (3dvof.py:590)")
    p.add_argument('--no-export', action='store_true', help="skip the .vtr files")
    p.add_argument('--sequence', action='store_true', help="one C-ABI entry per reference kernel instead of the fused vof3d_step")
    p.add_argument('--dump', type=str, default=None, help="write u,v,w,p,F (+istep) to this .npz at the end")
    p.add_argument('--resume', type=str, default=None, help="continue from a --dump file, bit-identical to an uninterrupted run")
    p.add_argument('--device', type=int, default=0)
    return p


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    from . import VofSolver3D, reference_params3d
    from .vtk import grid_to_vtk

    nx, ny, nz = args.nx, args.ny, args.nz
    L = [0.1 * n / 200.0 for n in (nx, ny, nz)] if args.scaled else [0.1, 0.1, 0.1]      # 3dvof.py:24-26
    P = reference_params3d(nx=nx, ny=ny, nz=nz, Lx=L[0], Ly=L[1], Lz=L[2], dt=args.dt, n_jacobi=args.jacobi, device=args.device)
    # banner, 3dvof.py:119-123
    print(f'>>> A 3D VOF solver on B200 (libvof, sm_100a); Ctrl-C to exit.')
    print(f'>>> Grid resolution: {nx} x {ny} x {nz}, dt = {P.dt:4.2e}')
    print(f'>>> Density ratio: {P.rho_l / P.rho_g : 4.2f}, gravity : {P.gy : 4.2f}, sigma : {P.sigma : 4.2f}')
    print(f'>>> Viscosity ratio: {P.nu_l / P.nu_g : 4.2f}')
    # for vtk file export, 3dvof.py:60-62
    xcor = np.linspace(0.0, 1.0, nx + 2).astype(np.float32)
    ycor = np.linspace(0.0, 1.0, ny + 2).astype(np.float32)
    zcor = np.linspace(0.0, 1.0, nz + 2).astype(np.float32)

    s = VofSolver3D(P)
    nstep = args.nstep
    if args.resume:
        st = np.load(args.resume)
        for k in ("u", "v", "w", "p", "F"):
            getattr(s, k).from_numpy(st[k])
        s.istep = int(st["istep"])
        print(f'>>> Resumed from {args.resume} at step {s.istep}')
    else:
        s.set_init_F(args.ic)               # 3dvof.py:591
    os.makedirs('output', exist_ok=True)    # 3dvof.py:593
    t0 = time.perf_counter()
    try:
        while args.steps == 0 or s.istep < args.steps:
            todo = nstep - (s.istep % nstep)
            if args.steps:
                todo = min(todo, args.steps - s.istep)
            if args.sequence:
                for _ in range(todo):
                    s.step_sequence()       # 3dvof.py:606-623, one call per reference kernel
            else:
                s.run(todo)
            istep = s.istep
            if (istep % nstep) == 0:        # 3dvof.py:624
                d = s.diagnostics()
                print(f'>>> Exporting step-{istep:05d} result...  VOF volume {d["mass"]:.6e}, max CFL {d["max_cfl"]:.3e}, '
                      f'{istep / (time.perf_counter() - t0):.1f} steps/s')
                if not args.no_export:
                    grid_to_vtk(f'./output/step-{istep:05d}', xcor, ycor, zcor, pointData={"VOF": np.ascontiguousarray(s.F.to_numpy())})
    except KeyboardInterrupt:
        pass
    s.synchronize()
    if args.dump:
        np.savez_compressed(args.dump, istep=s.istep, **{k: getattr(s, k).to_numpy() for k in ("u", "v", "w", "p", "F")})
    s.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
