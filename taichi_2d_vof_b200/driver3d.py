"""Drop-in for the main loop of the reference's 3-D script (`python 3dvof.py [-ic {1,2,3}] [-s]`).

Keeps the command line (3dvof.py:12-18; `-s` is parsed and ignored there too), the start-up banner (119-123), the
constants (20-38), the kernel call sequence (606-623) and the export cadence (every nstep = 100 steps: a
RectilinearGrid `output/step-%05d.vtr` with the point data "VOF" on the unit-cube coordinates, 60-62, 624-627), and
runs the kernels on a B200 through libvof (`vof3d_*`).  Extensions whose defaults reproduce the reference: `--steps`
(the GUI loop never ends by itself), `--nx/--ny/--nz`, `--scaled`, `--no-export`, `--sequence`, `--dump`, and
`--gpus N` (plane slabs along i on N GPUs of one box, one NVLink peer-store exchange of the halo planes per step;
BASELINE config 5 is `--nx 512 --ny 512 --nz 512 --scaled --gpus 8`).  The export does not stall the loop: device
snapshot -> pinned host buffer on a side stream -> writer thread (output.py); the reference blocks on F.to_numpy().
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
    p.add_argument('--gpus', type=int, default=1, help="plane-slab decomposition along i over this many GPUs of one box")
    p.add_argument('--transport', choices=['p2p', 'nccl'], default='p2p', help="halo exchange with --gpus > 1")
    return p


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else list(argv)
    args = build_parser().parse_args(argv)
    from . import launch
    if args.gpus > 1 and not launch.under_torchrun():
        return launch.respawn(args.gpus, "taichi_2d_vof_b200.driver3d", argv)
    rank, world, local, dist = launch.init()
    from . import VofSolver3D, reference_params3d
    from .output import FieldDumper
    from .slab import SlabSolver2D
    from .vtk import grid_to_vtk

    nx, ny, nz = args.nx, args.ny, args.nz
    L = [0.1 * n / 200.0 for n in (nx, ny, nz)] if args.scaled else [0.1, 0.1, 0.1]      # 3dvof.py:24-26
    device = local if local is not None else args.device

    def params_fn(slab, halo, device):
        return reference_params3d(nx=nx, ny=ny, nz=nz, Lx=L[0], Ly=L[1], Lz=L[2], dt=args.dt, n_jacobi=args.jacobi,
                                  slab=slab, halo=halo, device=device)

    P = params_fn(None, 0, device)
    say = print if rank == 0 else (lambda *a, **k: None)
    # banner, 3dvof.py:119-123
    say(f'>>> A 3D VOF solver on B200 (libvof, sm_100a); Ctrl-C to exit.')
    say(f'>>> Grid resolution: {nx} x {ny} x {nz}, dt = {P.dt:4.2e}' + (f' ({world} GPUs, plane slabs, halo exchange: {args.transport})' if world > 1 else ''))
    say(f'>>> Density ratio: {P.rho_l / P.rho_g : 4.2f}, gravity : {P.gy : 4.2f}, sigma : {P.sigma : 4.2f}')
    say(f'>>> Viscosity ratio: {P.nu_l / P.nu_g : 4.2f}')
    # for vtk file export, 3dvof.py:60-62
    xcor = np.linspace(0.0, 1.0, nx + 2).astype(np.float32)
    ycor = np.linspace(0.0, 1.0, ny + 2).astype(np.float32)
    zcor = np.linspace(0.0, 1.0, nz + 2).astype(np.float32)

    slab = SlabSolver2D(params_fn, nx, rank, world, dist=dist, n_jacobi=args.jacobi, device=device, transport=args.transport,
                        solver_cls=VofSolver3D, halo_fields=("F", "u", "v", "w", "p"))
    s = slab.solver
    nstep = args.nstep
    if args.resume:
        st = np.load(args.resume)
        for k in ("u", "v", "w", "p", "F"):
            slab.scatter(k, st[k])
        s.istep = int(st["istep"])
        say(f'>>> Resumed from {args.resume} at step {s.istep}')
    else:
        slab.set_init_F(args.ic)            # 3dvof.py:591
    dumper = None
    if rank == 0:
        os.makedirs('output', exist_ok=True)    # 3dvof.py:593
        if not args.no_export:
            dumper = FieldDumper((nx + 2, ny + 2, nz + 2),
                                 lambda Fnp, istep: grid_to_vtk(f'./output/step-{istep:05d}', xcor, ycor, zcor, pointData={"VOF": Fnp}))
    t0 = time.perf_counter()
    try:
        while args.steps == 0 or s.istep < args.steps:
            todo = nstep - (s.istep % nstep)
            if args.steps:
                todo = min(todo, args.steps - s.istep)
            if args.sequence and world == 1:
                for _ in range(todo):
                    s.step_sequence()       # 3dvof.py:606-623, one call per reference kernel
            else:
                slab.run(todo)
            istep = s.istep
            if (istep % nstep) == 0:        # 3dvof.py:624
                d = slab.diagnostics()
                say(f'>>> Exporting step-{istep:05d} result...  VOF volume {d["mass"]:.6e}, max CFL {d["max_cfl"]:.3e}, '
                    f'{istep / (time.perf_counter() - t0):.1f} steps/s')
                if not args.no_export:      # 3dvof.py:626-627, without blocking the loop
                    if world == 1:
                        dumper.dump(s.F, istep)
                    else:
                        Fnp = slab.gather("F")
                        if rank == 0:
                            dumper.dump_array(Fnp, istep)
    except KeyboardInterrupt:
        pass
    s.synchronize()
    if dumper is not None:
        dumper.close()
    if args.dump:
        st = {k: slab.gather(k) for k in ("u", "v", "w", "p", "F")}
        if rank == 0:
            np.savez_compressed(args.dump, istep=s.istep, **st)
    s.close()
    launch.finish(dist)
    return 0


if __name__ == "__main__":
    sys.exit(main())
