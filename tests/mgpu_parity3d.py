"""3-D multi-process slab parity (spawned by tests/test_round2_gpu.py under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/mgpu_parity3d.py [n] [steps]

Plane slabs along i, deep halo of 16 planes, one exchange of u, v, w, p, F per step (VOF_TRANSPORT=p2p: fused NVLink
peer-store kernel; nccl: send/recv); rank 0 also runs the whole domain and every owned plane must match it bit for bit."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200.slab import slab_parity_check  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 9
transport = os.environ.get("VOF_TRANSPORT", "p2p")
r = slab_parity_check(dist, rank, world, local, steps=steps, transport=transport, three_d=True, n=n)
ok = True
if rank == 0:
    print(r)
    ok = r["identical"] and r["volume_rel_diff"] <= 1e-9
    print("MGPU 3D PARITY", "OK" if ok else "FAILED", f"({world} ranks, {n}^3, {steps} steps, transport {r['transport']})")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
