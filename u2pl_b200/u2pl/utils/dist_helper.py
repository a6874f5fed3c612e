"""`setup_distributed` for the drop-in `u2pl.utils.dist_helper` (reference :13-46): one process per GPU,
rank / world size from SLURM or from the torchrun environment, NCCL process group over NVLink."""
import os
import subprocess

import torch
import torch.distributed as dist


def _slurm_env(port):
    """Translate SLURM variables into the env:// rendezvous torch.distributed expects."""
    env = os.environ
    rank, world = int(env["SLURM_PROCID"]), int(env["SLURM_NTASKS"])
    head = subprocess.getoutput(f"scontrol show hostname {env['SLURM_NODELIST']} | head -n1")
    if port is not None:
        env["MASTER_PORT"] = str(port)
    env.setdefault("MASTER_PORT", "10685")
    env.setdefault("MASTER_ADDR", head)
    env.update(WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank % torch.cuda.device_count()))
    return rank, world


def setup_distributed(backend="nccl", port=None):
    if "SLURM_JOB_ID" in os.environ:
        rank, world = _slurm_env(port)
    else:
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group(backend=backend, world_size=world, rank=rank)
    return rank, world
