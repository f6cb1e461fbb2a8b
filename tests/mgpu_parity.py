"""Multi-process slab parity (run under torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py

Every rank steps its slab (NCCL halo exchange over NVLink, one exchange per step); rank 0 also runs the
whole domain on its own GPU and all owned rows must match it bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200 import VofSolver2D, reference_params  # noqa: E402
from taichi_2d_vof_b200.slab import SlabSolver2D  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
nx, ny, steps, ic = 1024, 640, 12, 3


def params_fn(slab, halo, device):
    return reference_params(nx=nx, ny=ny, Lx=0.1 * nx / 200, Ly=0.1 * ny / 200, slab=slab, halo=halo, device=device)


transport = os.environ.get("VOF_TRANSPORT", "p2p")
s = SlabSolver2D(params_fn, nx, rank, world, dist=dist, device=local, transport=transport)
s.set_init_F(ic)
for _ in range(steps):
    s.step()
ok = True
full = None
if rank == 0:
    full = VofSolver2D(params_fn(None, 0, local)); full.set_init_F(ic)
    for _ in range(steps):
        full.step()
for name in ("F", "u", "v", "p"):
    mine = torch.from_numpy(s.owned(name)).cuda()
    parts = [torch.empty((hi - lo + 1, ny + 2), dtype=torch.float32, device="cuda") for lo, hi in s.parts]
    # all_gather needs equal shapes; pad to the tallest slab
    h = max(p.shape[0] for p in parts)
    pad = torch.zeros((h, ny + 2), dtype=torch.float32, device="cuda"); pad[: mine.shape[0]] = mine
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    if rank == 0:
        glob = np.concatenate([o[: p.shape[0]].cpu().numpy() for o, p in zip(out, parts)], axis=0)
        ref = getattr(full, name).to_numpy()[1:-1]
        bad = np.argwhere(glob != ref)
        print(f"{name}: {'identical' if bad.size == 0 else str(len(bad)) + ' cells differ, first ' + str(bad[0])}")
        ok = ok and bad.size == 0
m = torch.tensor([s.solver.mass()], dtype=torch.float64, device="cuda")
dist.all_reduce(m)
if rank == 0:
    print("global volume", float(m), "single-GPU", full.mass())
    ok = ok and abs(float(m) - full.mass()) <= 1e-9 * full.mass()
    print("MGPU PARITY", "OK" if ok else "FAILED", f"({world} ranks, {nx}x{ny}, {steps} steps, transport {transport})")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
