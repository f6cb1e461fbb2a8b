"""Drop-in for the main loop of the reference script (`python 2dvof.py [-ic {1,2,3}] [-s]`).

Keeps the reference's command line (2dvof.py:11-17), start-up banner (95-99), constants
(19-34), kernel call sequence (513-528) and output cadence (every nstep = 100 steps, 497/531),
and runs the kernels on a B200 through libvof.  Differences, all opt-in extensions whose
defaults reproduce the reference: the loop can end (`--steps`), needs no display (`ti.GUI` is
replaced by a text progress line), grid/domain are flags instead of edited constants, `-s`
writes `output/%06d-f.npy` (+ the reference's `output/%06d-f.png` when matplotlib exists) through a
non-stalling read (device snapshot -> pinned host buffer on a side stream -> writer thread, see output.py;
the reference blocks on F.to_numpy() and plt.savefig), and `--gpus N` runs the same loop on N GPUs of one box
(row slabs along i, one NVLink peer-store halo exchange per step, all-reduced diagnostics; fields bit-identical
to the single-GPU run).
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
    p.add_argument('--pressure-solver', choices=['jacobi', 'chebyshev'], default='jacobi',
                   help="'chebyshev': the same number of sweeps of the Chebyshev-accelerated Jacobi iteration (stronger projection; "
                        "not the reference's arithmetic)")
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
    p.add_argument('--gpus', type=int, default=1, help="row-slab decomposition over this many GPUs of one box (one process per GPU)")
    p.add_argument('--transport', choices=['p2p', 'nccl'], default='p2p', help="halo exchange with --gpus > 1")
    return p


def _save_png(Fnp, path, Lx, Ly):
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        return
    fx, fy = 5, Ly / Lx * 5                  # 2dvof.py:566-570
    plt.figure(figsize=(fx, fy))
    plt.axis('off')
    plt.contourf(Fnp.T, cmap=plt.cm.Blues)
    plt.savefig(path)
    plt.close()


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else list(argv)
    args = build_parser().parse_args(argv)
    from . import launch
    if args.gpus > 1 and not launch.under_torchrun():
        return launch.respawn(args.gpus, "taichi_2d_vof_b200.driver", argv)
    rank, world, local, dist = launch.init()
    from . import reference_params
    from .output import FieldDumper
    from .slab import SlabSolver2D

    nx, ny = args.nx, args.ny
    if args.scaled:
        Lx, Ly = 0.1 * nx / 200.0, 0.1 * ny / 200.0
    elif args.L is not None:
        Lx, Ly = args.L, args.L * ny / nx
    else:
        Lx = Ly = 0.1                       # 2dvof.py:22-23
    device = local if local is not None else args.device

    def params_fn(slab, halo, device):
        return reference_params(nx=nx, ny=ny, Lx=Lx, Ly=Ly, dt=args.dt, n_jacobi=args.jacobi, slab=slab, halo=halo, device=device)

    P = params_fn(None, 0, device)
    initial_condition = args.ic
    SAVE_FIG = args.s
    say = print if rank == 0 else (lambda *a, **k: None)

    # banner, 2dvof.py:95-99
    say(f'>>> A VOF solver on B200 (libvof, sm_100a); Ctrl-C to exit.')
    say(f'>>> Grid resolution: {nx} x {ny}, dt = {P.dt:4.2e}' + (f' ({world} GPUs, row slabs, halo exchange: {args.transport})' if world > 1 else ''))
    say(f'>>> Density ratio: {P.rho_l / P.rho_g : 4.2f}, gravity : {P.gy : 4.2f}, sigma : {P.sigma : 4.2f}')
    say(f'>>> Viscosity ratio: {P.nu_l / P.nu_g : 4.2f}')

    slab = SlabSolver2D(params_fn, nx, rank, world, dist=dist, n_jacobi=args.jacobi, device=device, transport=args.transport)
    s = slab.solver
    if args.pressure_solver == 'chebyshev':
        from . import _lib
        s.set_option(_lib.VOF_OPT_PRESSURE_SOLVER, 1)
        say('>>> Pressure solver: Chebyshev-accelerated Jacobi (opt-in; results differ from the reference)')
    nstep = args.nstep                      # 2dvof.py:497
    if args.resume:
        st = np.load(args.resume)
        for k in ("u", "v", "p", "F"):
            slab.scatter(k, st[k])
        s.istep = int(st["istep"])
        say(f'>>> Resumed from {args.resume} at step {s.istep}')
    else:
        slab.set_init_F(initial_condition)  # 2dvof.py:498 (no set_BC before the first step)
    if rank == 0:
        os.makedirs('output', exist_ok=True)    # 2dvof.py:500
    dumper = None
    if SAVE_FIG and rank == 0:
        def write(Fnp, count):
            np.save(f'output/{count:06d}-f.npy', Fnp)
            _save_png(Fnp, f'output/{count:06d}-f.png', Lx, Ly)
        dumper = FieldDumper((nx + 2, ny + 2), write)
    t0 = time.perf_counter()
    try:
        while args.steps == 0 or s.istep < args.steps:
            todo = nstep - (s.istep % nstep)
            if args.steps:
                todo = min(todo, args.steps - s.istep)
            if args.sequence and world == 1:
                for _ in range(todo):
                    s.step_sequence()       # 2dvof.py:513-528, one call per reference kernel
            else:
                slab.run(todo)              # same result: fused kernels; CUDA-graph replay on one GPU
            istep = s.istep
            if (istep % nstep) == 0:        # 2dvof.py:531
                d = slab.diagnostics(residual=False)     # one all-gather of {volume, max CFL, Courant count} when world > 1
                rate = istep / (time.perf_counter() - t0)
                say(f'>>> Number of steps:{istep:<5d}, Time:{istep * P.dt:5.2e} sec. '
                    f'VOF volume {d["mass"]:.6e}, max CFL {d["max_cfl"]:.3e}, {rate:.1f} steps/s')
                if d["courant_count"]:
                    say(f'U/V velocity courant number > 1 on {d["courant_count"]} faces')   # 2dvof.py:274-280
                if args.view and world == 1:   # 2dvof.py:533-561, headless: rgb_buf / V to disk
                    img = {'vof': s.get_vof_field, 'u': s.get_u_field, 'v': s.get_v_field, 'vnorm': s.get_vnorm_field,
                           'vectors': s.interp_velocity}[args.view]()
                    np.save(f'output/{istep // nstep - 1:06d}-{args.view}.npy', img)
                if SAVE_FIG:                # 2dvof.py:563-571, without blocking the loop
                    count = istep // nstep - 1
                    if world == 1:
                        dumper.dump(s.F, count)
                    else:
                        Fnp = slab.gather("F")
                        if rank == 0:
                            dumper.dump_array(Fnp, count)
    except KeyboardInterrupt:
        pass
    s.synchronize()
    if dumper is not None:
        dumper.close()
    if args.dump:
        st = {k: slab.gather(k) for k in ("u", "v", "p", "F")}
        if rank == 0:
            np.savez_compressed(args.dump, istep=s.istep, **st)
    s.close()
    launch.finish(dist)
    return 0


if __name__ == "__main__":
    sys.exit(main())
