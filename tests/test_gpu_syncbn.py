"""2 GPUs (NCCL): (1) the peer-memory all-reduce (csrc/peer_reduce.cu) against a rank-ordered sum of the all-gathered
inputs -- bitwise, on both ranks, over hundreds of back-to-back calls (slot reuse); (2) the SyncBN-aware fused
BatchNorm (fused.bn_act on nn.SyncBatchNorm modules, statistics exchanged through that kernel) against torch's own
nn.SyncBatchNorm (reference base.py:6-8) forward, backward and running statistics.

Runs whenever >= 2 GPUs are visible:  gpurun --gpus 2 -- 'python -m pytest tests/test_gpu_syncbn.py -x -q'"""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]


def _init(rank, world, port_no):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    return dist


def _run_allreduce(rank, world, port_no, ret):
    dist = _init(rank, world, port_no)
    from u2pl_b200 import peer
    ok, used_kernel = True, False
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    for it in range(300):                                          # back-to-back: both parities, slot reuse
        n = [2 * 64, 2 * 256, 2 * 2048, 7, 4096][it % 5]
        x = torch.randn(n, device="cuda", generator=g)
        gathered = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(gathered, x)
        want = torch.zeros_like(x)
        for r in range(world):                                     # the kernel's summation order
            want += gathered[r]
        got = peer.allreduce_small_(x.clone())
        used_kernel = used_kernel or peer._REDUCER is not None
        ok = ok and bool(torch.equal(got, want))
    ret[rank] = (ok, used_kernel)
    dist.destroy_process_group()


def test_peer_allreduce_is_rank_ordered_sum():
    import torch.multiprocessing as mp
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_run_allreduce, args=(2, 34100 + os.getpid() % 1000, ret), nprocs=2, join=True)
        assert ret[0] == (True, True) and ret[1] == (True, True), dict(ret)


def _run_syncbn(rank, world, port_no, ret):
    dist = _init(rank, world, port_no)
    import torch.nn as nn
    from u2pl_b200 import fused
    torch.manual_seed(7)                                           # same parameters on both ranks
    C = 256
    ref = nn.SyncBatchNorm(C).cuda()
    mine = nn.SyncBatchNorm(C).cuda()
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5)
        ref.bias.normal_()
        mine.load_state_dict(ref.state_dict())
    g = torch.Generator(device="cuda").manual_seed(50 + rank)       # different data per rank
    out = {}
    for step in range(2):
        x = (torch.randn(3, C, 17, 19, device="cuda", generator=g) * (1 + rank) + 0.3 * rank).bfloat16()
        x = x.contiguous(memory_format=torch.channels_last)
        go = torch.randn(3, C, 17, 19, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
        xa = x.clone().requires_grad_(True)
        xb = x.float().requires_grad_(True)                        # torch's SyncBatchNorm in fp32 on the same values
        ya = fused.bn_act(xa, mine, relu=True)
        yb = torch.relu(ref(xb))
        ya.backward(go)
        yb.backward(go.float())
        out[f"y{step}"] = float((ya.float() - yb).abs().max() / yb.abs().max())
        out[f"dx{step}"] = float((xa.grad.float() - xb.grad).abs().max() / xb.grad.abs().max())
        out[f"dw{step}"] = float((mine.weight.grad - ref.weight.grad).abs().max() / ref.weight.grad.abs().max())
        out[f"db{step}"] = float((mine.bias.grad - ref.bias.grad).abs().max() / ref.bias.grad.abs().max())
    out["mean"] = float((mine.running_mean - ref.running_mean).abs().max())
    out["var"] = float((mine.running_var - ref.running_var).abs().max() / ref.running_var.abs().max())
    out["nbt"] = int(mine.num_batches_tracked) == int(ref.num_batches_tracked)
    ret[rank] = out
    dist.destroy_process_group()


def test_fused_syncbn_matches_torch_syncbatchnorm():
    import torch.multiprocessing as mp
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_run_syncbn, args=(2, 34300 + os.getpid() % 1000, ret), nprocs=2, join=True)
        for rank in range(2):
            o = ret[rank]
            for k in ("y0", "y1", "dx0", "dx1"):
                assert o[k] <= 1.5e-2, (rank, k, o[k])             # bf16 outputs vs fp32 reference
            for k in ("dw0", "dw1", "db0", "db1"):
                assert o[k] <= 2e-3, (rank, k, o[k])
            assert o["mean"] <= 1e-5 and o["var"] <= 1e-5 and o["nbt"], (rank, o)
