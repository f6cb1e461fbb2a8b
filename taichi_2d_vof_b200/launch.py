"""`--gpus N` for the drop-in drivers: one process per GPU (torch.distributed over NCCL for the plumbing, NVLink
peer stores for the halo rows).  A driver started with `--gpus N > 1` outside torchrun re-launches itself as
`python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ...` and returns its exit code."""
from __future__ import annotations

import os
import socket
import subprocess
import sys


def under_torchrun() -> bool:
    return "RANK" in os.environ and "WORLD_SIZE" in os.environ


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def respawn(gpus: int, module: str, argv) -> int:
    """Run `python -m <module> <argv>` once per GPU under torchrun; returns the exit code."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={gpus}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), "-m", module, *argv]
    return subprocess.call(cmd, env=env)


def init():
    """(rank, world, local_rank, dist or None); sets the CUDA device of this rank."""
    if not under_torchrun():
        return 0, 1, None, None
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    return rank, world, local, dist


def finish(dist):
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
