"""2 GPUs (NCCL): the class-sharded, peer-mapped memory bank (bank.ShardedBank + u2pl_infonce_forward_sharded) must
give bit-identical losses, gradients and bank contents to the replicated device bank (bank.DeviceBank), which the
single-GPU tests pin to the oracle.  Each rank feeds its own slice of the golden fixture's steps, so keys from both
ranks interleave in rank order exactly as in utils.py:21-38.

Runs whenever >= 2 GPUs are visible:  gpurun --gpus 2 -- 'python -m pytest tests/test_gpu_sharded_bank.py -x -q'"""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "contra_c21.npz")


def _run(rank, world, port_no, sharded, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    os.environ["U2PL_BANK_SHARDED"] = "1" if sharded else "0"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from u2pl_b200 import bank as bank_mod, contra
    g = np.load(GOLDEN)
    cfg = {k: v for k, v in zip(g["cfg_keys"].tolist(), g["cfg_vals"].tolist())}
    for k in ("low_rank", "high_rank", "num_negatives", "num_queries"):
        cfg[k] = int(cfg[k])
    C, D = g["s0_label_l"].shape[1], g["s0_rep"].shape[1]
    qsize = g["queue_size"].tolist()
    memobank = [[torch.zeros(0, D)] for _ in range(C)]
    ptrs = [torch.zeros(1, dtype=torch.long) for _ in range(C)]
    steps = int(g["steps"])
    out = {"loss": [], "grad": [], "keys": []}
    for it in range(steps):
        s = (it + rank) % steps                                   # ranks see different batches
        dev = lambda a, dt=None: (torch.from_numpy(np.ascontiguousarray(a)).to(dt) if dt else torch.from_numpy(np.ascontiguousarray(a))).cuda()  # noqa: E731
        rep = dev(g[f"s{s}_rep"]).requires_grad_(True)
        args = [dev(g[f"s{s}_label_l"].astype(np.int64)), dev(g[f"s{s}_label_u"].astype(np.int64)), dev(g[f"s{s}_prob_l"]),
                dev(g[f"s{s}_prob_u"]), dev(g[f"s{s}_low_mask"].astype(np.float32)), dev(g[f"s{s}_high_mask"].astype(np.float32))]
        torch.manual_seed(500 + 10 * it + rank)
        new_keys, loss = contra.compute_contra_memobank_loss(rep, *args, cfg, memobank, ptrs, qsize, dev(g[f"s{s}_rep_teacher"]))
        loss.backward()
        out["loss"].append(float(loss))
        out["grad"].append(rep.grad.cpu().numpy().copy() if rep.grad is not None else None)
        out["keys"].append(list(new_keys))
    bank = contra.bank_for(memobank, qsize, D, torch.device("cuda", rank))
    assert isinstance(bank, bank_mod.ShardedBank) == bool(sharded)
    dist.barrier()
    out["bank"] = [bank.materialize(c).cpu().numpy() for c in range(C)]
    dist.barrier()
    ret[(rank, sharded)] = out
    if sharded:
        bank.close()
    contra.forget_banks()
    dist.destroy_process_group()


def test_sharded_bank_equals_replicated_bank():
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        for i, sharded in enumerate((False, True)):
            mp.spawn(_run, args=(world, 33000 + 7 * i + os.getpid() % 1000, sharded, ret), nprocs=world, join=True)
        for rank in range(world):
            a, b = ret[(rank, False)], ret[(rank, True)]
            assert a["keys"] == b["keys"]
            assert a["loss"] == b["loss"], (a["loss"], b["loss"])                      # same rows, same kernel maths
            assert any(x > 0 for x in a["loss"])
            for ga, gb in zip(a["grad"], b["grad"]):
                # the backward accumulates duplicate anchors (sampling with replacement, loss_helper.py:179-181) with
                # floating-point atomics: the sum is the same, its rounding order is not reproducible run to run
                assert (ga is None) == (gb is None)
                if ga is not None:
                    assert np.array_equal(ga != 0, gb != 0) and np.allclose(ga, gb, rtol=1e-5, atol=1e-9)
            for ca, cb in zip(a["bank"], b["bank"]):
                assert np.array_equal(ca, cb)
        for c in range(len(ret[(0, True)]["bank"])):                                   # both ranks see the same bank
            assert np.array_equal(ret[(0, True)]["bank"][c], ret[(1, True)]["bank"][c])
