"""Drop-in for the main loop of the reference script (`python 2dvof.py [-ic {1,2,3}] [-s]`).

Keeps the reference's command line (2dvof.py:11-17), start-up banner (95-99), constants
(19-34), kernel call sequence (513-528) and output cadence (every nstep = 100 steps, 497/531),
and runs the kernels on a B200 through libvof.  Differences, all opt-in extensions whose
defaults reproduce the reference: the loop can end (`--steps`), needs no display (`ti.GUI` is
replaced by a text progress line), grid/domain are flags instead of edited constants, and `-s`
writes `output/%06d-f.npy` (+ the reference's `output/%06d-f.png` when matplotlib exists).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="B200-native VOF solver (drop-in for 2dvof.py's loop)")
    # the reference's two flags, verbatim (2dvof.py:13-14)
    p.add_argument('-ic', type=int, choices=[1, 2, 3], default=1)
    p.add_argument('-s', action='store_true')
    # extensions (defaults = reference behaviour)
    p.add_argument('--nx', type=int, default=200)
    p.add_argument('--ny', type=int, default=200)
    p.add_argument('--L', type=float, default=None, help="domain edge Lx = Ly (default 0.1; 'scaled' runs use 0.1*n/200)")
    p.add_argument('--scaled', action='store_true', help="constant-dx scaling: L = 0.1 * nx / 200")
    p.add_argument('--dt', type=float, default=4e-6)
    p.add_argument('--jacobi', type=int, default=10)
    p.add_argument('--steps', type=int, default=0, help="stop after this many steps (0 = run until interrupted, like the GUI loop)")
    p.add_argument('--nstep', type=int, default=100, help="output interval (2dvof.py:497)")
    p.add_argument('--sequence', action='store_true', help="call one C-ABI entry per reference kernel instead of the fused vof2d_step")
    p.add_argument('--view', choices=['vof', 'u', 'v', 'vnorm', 'vectors'], default=None,
                   help="every nstep steps also run the GUI loop's display kernel of that view (2dvof.py:530-561: the 'v' key "
                        "cycles through these) and write output/%%06d-<view>.npy instead of painting a window")
    p.add_argument('--dump', type=str, default=None, help="write u,v,p,F (+istep) to this .npz at the end")
    p.add_argument('--resume', type=str, default=None, help="continue from a --dump file (u, v, p, F and istep are the whole state: "
                   "rho, nu, kappa, u*, v* are recomputed every step), bit-identical to an uninterrupted run")
    p.add_argument('--device', type=int, default=0)
    return p


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    from . import VofSolver2D, reference_params

    nx, ny = args.nx, args.ny
    if args.scaled:
        Lx, Ly = 0.1 * nx / 200.0, 0.1 * ny / 200.0
    elif args.L is not None:
        Lx, Ly = args.L, args.L * ny / nx
    else:
        Lx = Ly = 0.1                       # 2dvof.py:22-23
    P = reference_params(nx=nx, ny=ny, Lx=Lx, Ly=Ly, dt=args.dt, n_jacobi=args.jacobi, device=args.device)
    initial_condition = args.ic
    SAVE_FIG = args.s

    # banner, 2dvof.py:95-99
    print(f'>>> A VOF solver on B200 (libvof, sm_100a); Ctrl-C to exit.')
    print(f'>>> Grid resolution: {nx} x {ny}, dt = {P.dt:4.2e}')
    print(f'>>> Density ratio: {P.rho_l / P.rho_g : 4.2f}, gravity : {P.gy : 4.2f}, sigma : {P.sigma : 4.2f}')
    print(f'>>> Viscosity ratio: {P.nu_l / P.nu_g : 4.2f}')

    s = VofSolver2D(P)
    nstep = args.nstep                      # 2dvof.py:497
    if args.resume:
        st = np.load(args.resume)
        for k in ("u", "v", "p", "F"):
            getattr(s, k).from_numpy(st[k])
        s.istep = int(st["istep"])
        print(f'>>> Resumed from {args.resume} at step {s.istep}')
    else:
        s.set_init_F(initial_condition)     # 2dvof.py:498 (no set_BC before the first step)
    os.makedirs('output', exist_ok=True)    # 2dvof.py:500
    t0 = time.perf_counter()
    try:
        while args.steps == 0 or s.istep < args.steps:
            todo = nstep - (s.istep % nstep)
            if args.steps:
                todo = min(todo, args.steps - s.istep)
            if args.sequence:
                for _ in range(todo):
                    s.step_sequence()       # 2dvof.py:513-528, one call per reference kernel
            else:
                s.run(todo)                 # same result, fused kernels / CUDA-graph replay
            istep = s.istep
            if (istep % nstep) == 0:        # 2dvof.py:531
                d = s.diagnostics(residual=False)
                rate = istep / (time.perf_counter() - t0)
                print(f'>>> Number of steps:{istep:<5d}, Time:{istep * P.dt:5.2e} sec. '
                      f'VOF volume {d["mass"]:.6e}, max CFL {d["max_cfl"]:.3e}, {rate:.1f} steps/s')
                if d["courant_count"]:
                    print(f'U/V velocity courant number > 1 on {d["courant_count"]} faces')   # 2dvof.py:274-280
                if args.view:               # 2dvof.py:533-561, headless: rgb_buf / V to disk
                    img = {'vof': s.get_vof_field, 'u': s.get_u_field, 'v': s.get_v_field, 'vnorm': s.get_vnorm_field,
                           'vectors': s.interp_velocity}[args.view]()
                    np.save(f'output/{istep // nstep - 1:06d}-{args.view}.npy', img)
                if SAVE_FIG:                # 2dvof.py:563-571
                    count = istep // nstep - 1
                    Fnp = s.F.to_numpy()
                    np.save(f'output/{count:06d}-f.npy', Fnp)
                    try:
                        import matplotlib
                        matplotlib.use("Agg")
                        import matplotlib.pyplot as plt
                        fx, fy = 5, Ly / Lx * 5
                        plt.figure(figsize=(fx, fy))
                        plt.axis('off')
                        plt.contourf(Fnp.T, cmap=plt.cm.Blues)
                        plt.savefig(f'output/{count:06d}-f.png')
                        plt.close()
                    except ImportError:
                        pass
    except KeyboardInterrupt:
        pass
    s.synchronize()
    if args.dump:
        st = {k: getattr(s, k).to_numpy() for k in ("u", "v", "p", "F")}
        np.savez_compressed(args.dump, istep=s.istep, **st)
    s.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
