"""Multi-process slab parity (spawned by tests/test_round2_gpu.py under torchrun on >= 2 GPUs; also runnable by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py

Every rank steps its slab (VOF_TRANSPORT=p2p: fused NVLink peer-store exchange; nccl: send/recv; one exchange per
step); rank 0 also runs the whole domain on its own GPU and all owned rows must match it bit for bit; the all-reduced
global volume must equal the single-GPU one."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_2d_vof_b200.slab import slab_parity_check  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
transport = os.environ.get("VOF_TRANSPORT", "p2p")
r = slab_parity_check(dist, rank, world, local, nx=1024, ny=640, steps=12, ic=3, transport=transport)
ok = True
if rank == 0:
    print(r)
    ok = r["identical"] and r["volume_rel_diff"] <= 1e-9
    print("MGPU PARITY", "OK" if ok else "FAILED", f"({world} ranks, transport {r['transport']})")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
