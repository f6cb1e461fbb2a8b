"""3-D multi-process slab parity (torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/mgpu_parity3d.py [n]

Plane slabs along i, deep halo of 16 planes, one NCCL exchange of u, v, w, p, F per step; rank 0 also runs the
whole domain and every owned plane must match it bit for bit.  With n = 512 this is BASELINE config 5."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver3D, reference_params3d  # noqa: E402
from taichi_2d_vof_b200.slab import SlabSolver2D  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 9
nx = ny = nz = n
L = 0.1 * n / 200


def params_fn(slab, halo, device):
    return reference_params3d(nx=nx, ny=ny, nz=nz, Lx=L, Ly=L, Lz=L, slab=slab, halo=halo, device=device)


s = SlabSolver2D(params_fn, nx, rank, world, dist=dist, device=local, transport="nccl", solver_cls=VofSolver3D,
                 halo_fields=("F", "u", "v", "w", "p"))
s.set_init_F(1)
for _ in range(2):
    s.step()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(steps - 2):
    s.step()
s.solver.synchronize(); dist.barrier()
dt = (time.perf_counter() - t0) / (steps - 2)
ok = True
full = None
if rank == 0:
    full = VofSolver3D(params_fn(None, 0, local)); full.set_init_F(1)
    for _ in range(steps):
        full.step()
for name in ("F", "u", "v", "w", "p"):
    mine = torch.from_numpy(s.owned(name)).cuda()
    h = max(hi - lo + 1 for lo, hi in s.parts)
    pad = torch.zeros((h, ny + 2, nz + 2), dtype=torch.float32, device="cuda"); pad[: mine.shape[0]] = mine
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    if rank == 0:
        glob = np.concatenate([o[: hi - lo + 1].cpu().numpy() for o, (lo, hi) in zip(out, s.parts)], axis=0)
        ref = getattr(full, name).to_numpy()[1:-1]
        same = np.array_equal(glob, ref)
        print(f"{name}: {'identical' if same else 'DIFFERS'}")
        ok = ok and same
if rank == 0:
    print("MGPU 3D PARITY", "OK" if ok else "FAILED", f"({world} ranks, {n}^3, {steps} steps); slab run {1 / dt:.1f} steps/s, "
          f"{10 * n ** 3 / dt / 1e9:.1f} Jacobi Gcell-updates/s")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
