"""Drop-in for u2pl/utils/dist_helper.py:13-46: one process per GPU, NCCL over NVLink."""
import os
import subprocess

import torch
import torch.distributed as dist


def setup_distributed(backend="nccl", port=None):
    num_gpus = torch.cuda.device_count()
    if "SLURM_JOB_ID" in os.environ:
        rank = int(os.environ["SLURM_PROCID"])
        world_size = int(os.environ["SLURM_NTASKS"])
        addr = subprocess.getoutput(f"scontrol show hostname {os.environ['SLURM_NODELIST']} | head -n1")
        if port is not None:
            os.environ["MASTER_PORT"] = str(port)
        os.environ.setdefault("MASTER_PORT", "10685")
        os.environ.setdefault("MASTER_ADDR", addr)
        os.environ["WORLD_SIZE"] = str(world_size)
        os.environ["LOCAL_RANK"] = str(rank % num_gpus)
        os.environ["RANK"] = str(rank)
    else:
        rank = int(os.environ["RANK"])
        world_size = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank % num_gpus)
    dist.init_process_group(backend=backend, world_size=world_size, rank=rank)
    return rank, world_size
