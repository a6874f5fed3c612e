"""CPU, world_size 2 over gloo: the multi-rank bank protocol of u2pl_b200/contra.py -- all_gather of the
per-class key counts, all_gather of the padded packed keys, rank-ordered ring append -- must leave every
rank with the same bank, equal to the reference's rank-order concatenation (utils.py:16-38).  The GPU
kernels are replaced here by numpy copies; the host logic (bank.plan_append, offsets, padding) is the
code under test."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import port
from u2pl_b200.bank import Ring, physical_rows, plan_append


def _worker(rank, world, port_no, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C, D, caps = 4, 3, [7, 5, 6, 9]
    rings, base = [], 0
    for cap in caps:
        rings.append(Ring(cap=cap, row_base=base))
        base += cap
    store = np.zeros((base, D), np.float32)
    o_bank = [[np.zeros((0, D), np.float32)] for _ in range(C)]
    o_ptr = [[0] for _ in range(C)]
    for step in range(6):
        rng = np.random.default_rng(1000 * step)                  # same stream on every rank
        counts_all = rng.integers(0, 6, (world, C))
        keys_all = [[rng.standard_normal((counts_all[r, c], D)).astype(np.float32) for c in range(C)] for r in range(world)]
        mine = torch.from_numpy(counts_all[rank].astype(np.int32))
        gathered_counts = [torch.zeros(C, dtype=torch.int32) for _ in range(world)]
        dist.all_gather(gathered_counts, mine)                    # contra.py: one collective for all classes
        neg = torch.stack(gathered_counts).numpy().astype(np.int64)
        assert np.array_equal(neg, counts_all)
        class_base = np.concatenate([np.zeros((world, 1), np.int64), np.cumsum(neg, axis=1)[:, :-1]], axis=1)
        kmax = int(neg.sum(1).max())
        packed = np.zeros((max(kmax, 1), D), np.float32)
        for c in range(C):
            packed[class_base[rank, c]: class_base[rank, c] + neg[rank, c]] = keys_all[rank][c]
        chunks = [torch.zeros(max(kmax, 1), D) for _ in range(world)]
        dist.all_gather(chunks, torch.from_numpy(packed))
        gathered = torch.cat(chunks).numpy()
        for c in range(C):
            descs, k = plan_append(rings[c], neg[:, c], [r * max(kmax, 1) + class_base[r, c] for r in range(world)])
            for (src, rb, first, cap, cnt) in descs:
                for i in range(cnt):
                    store[rb + (first + i) % cap] = gathered[src + i]
            port.dequeue_and_enqueue([keys_all[r][c] for r in range(world)], o_bank[c], o_ptr[c], caps[c])
            got = store[physical_rows(rings[c], np.arange(rings[c].length))]
            assert np.array_equal(got, o_bank[c][0]) and rings[c].ptr == o_ptr[c][0]
    digest = torch.tensor([float(np.abs(store).sum())])
    both = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(both, digest)
    assert both[0].item() == both[1].item()                       # replicated banks are identical across ranks
    ret[rank] = True
    dist.destroy_process_group()


def test_two_rank_bank_protocol():
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, 29000 + os.getpid() % 2000, ret), nprocs=world, join=True)
        assert ret.get(0) and ret.get(1)


# ------------------------------------------------------------------ class-sharded bank (ShardedBank's host logic)
def _sharded_worker(rank, world, port_no, ret):
    """Same protocol, but every rank stores only the classes it owns (owner = c % world).  Ring state is advanced for
    ALL classes on every rank, rows are written by the owner only; a reader addresses `shard[owner][row_base + ...]`.
    The peer-memory read of the real system (CUDA IPC) is emulated by all-gathering the shards at the end."""
    from u2pl_b200.bank import shard_layout
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C, D, caps = 5, 3, [7, 5, 6, 9, 4]
    owner, bases, shard_rows = shard_layout(caps, world)
    assert owner == [c % world for c in range(C)] and shard_rows == [7 + 6 + 4, 5 + 9]
    rings = [Ring(cap=cap, row_base=b) for cap, b in zip(caps, bases)]
    mine = np.zeros((max(shard_rows), D), np.float32)             # padded to a common size for the final all_gather
    o_bank = [[np.zeros((0, D), np.float32)] for _ in range(C)]
    o_ptr = [[0] for _ in range(C)]
    for step in range(6):
        rng = np.random.default_rng(77 * step)
        counts_all = rng.integers(0, 6, (world, C))
        keys_all = [[rng.standard_normal((counts_all[r, c], D)).astype(np.float32) for c in range(C)] for r in range(world)]
        class_base = np.concatenate([np.zeros((world, 1), np.int64), np.cumsum(counts_all, axis=1)[:, :-1]], axis=1)
        kmax = max(int(counts_all.sum(1).max()), 1)
        packed = np.zeros((kmax, D), np.float32)
        for c in range(C):
            packed[class_base[rank, c]: class_base[rank, c] + counts_all[rank, c]] = keys_all[rank][c]
        chunks = [torch.zeros(kmax, D) for _ in range(world)]
        dist.all_gather(chunks, torch.from_numpy(packed))
        gathered = torch.cat(chunks).numpy()
        for c in range(C):
            descs, _ = plan_append(rings[c], counts_all[:, c], [r * kmax + class_base[r, c] for r in range(world)])
            if owner[c] == rank:
                for (src, rb, first, cap, cnt) in descs:
                    for i in range(cnt):
                        mine[rb + (first + i) % cap] = gathered[src + i]
            port.dequeue_and_enqueue([keys_all[r][c] for r in range(world)], o_bank[c], o_ptr[c], caps[c])
    shards = [torch.zeros(max(shard_rows), D) for _ in range(world)]
    dist.all_gather(shards, torch.from_numpy(mine))               # stands in for the peer mapping
    for c in range(C):
        got = shards[owner[c]].numpy()[physical_rows(rings[c], np.arange(rings[c].length))]
        assert np.array_equal(got, o_bank[c][0]) and rings[c].ptr == o_ptr[c][0], c
    ret[rank] = True
    dist.destroy_process_group()


def test_two_rank_sharded_bank_protocol():
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_sharded_worker, args=(world, 31000 + os.getpid() % 2000, ret), nprocs=world, join=True)
        assert ret.get(0) and ret.get(1)


# ------------------------------------------------------------------ contra.py, world 2, replicated vs class-sharded bank
def _contra_worker(rank, world, port_no, sharded, ret):
    """compute_contra_memobank_loss on two gloo ranks over the emulated C ABI (tests/emulated_abi.py; CUDA IPC is stood in
    for by POSIX shared memory): the CPU analogue of tests/test_gpu_sharded_bank.py."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emulated_abi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    os.environ["U2PL_BANK_SHARDED"] = "1" if sharded else "0"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emulated_abi.install(emulated_abi.DirectPatcher())
    from u2pl_b200 import bank as bank_mod, contra
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "contra_c21.npz"))
    cfg = {k: v for k, v in zip(g["cfg_keys"].tolist(), g["cfg_vals"].tolist())}
    for k in ("low_rank", "high_rank", "num_negatives", "num_queries"):
        cfg[k] = int(cfg[k])
    C, D = g["s0_label_l"].shape[1], g["s0_rep"].shape[1]
    qsize = g["queue_size"].tolist()
    memobank = [[torch.zeros(0, D)] for _ in range(C)]
    ptrs = [torch.zeros(1, dtype=torch.long) for _ in range(C)]
    steps = int(g["steps"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))                                 # noqa: E731
    out = {"loss": [], "grad": [], "keys": []}
    for it in range(steps):
        s = (it + rank) % steps
        rep = t(g[f"s{s}_rep"]).requires_grad_(True)
        args = [t(g[f"s{s}_label_l"].astype(np.int64)), t(g[f"s{s}_label_u"].astype(np.int64)), t(g[f"s{s}_prob_l"]), t(g[f"s{s}_prob_u"]),
                t(g[f"s{s}_low_mask"].astype(np.float32)), t(g[f"s{s}_high_mask"].astype(np.float32))]
        torch.manual_seed(500 + 10 * it + rank)
        new_keys, loss = contra.compute_contra_memobank_loss(rep, *args, cfg, memobank, ptrs, qsize, t(g[f"s{s}_rep_teacher"]))
        loss.backward()
        out["loss"].append(float(loss.detach()))
        out["grad"].append(rep.grad.numpy().copy() if rep.grad is not None else None)
        out["keys"].append(list(new_keys))
    bank = contra.bank_for(memobank, qsize, D, torch.device("cpu"))
    assert isinstance(bank, bank_mod.ShardedBank) == bool(sharded)
    dist.barrier()
    out["bank"] = [bank.materialize(c).numpy().copy() for c in range(C)]
    out["owned"] = [bank.owns(c) for c in range(C)] if sharded else None
    dist.barrier()
    ret[(rank, sharded)] = out
    if sharded:
        bank.close()
    dist.destroy_process_group()


def test_two_rank_contra_sharded_equals_replicated():
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        for i, sharded in enumerate((False, True)):
            mp.spawn(_contra_worker, args=(world, 35000 + 11 * i + os.getpid() % 1500, sharded, ret), nprocs=world, join=True)
        for rank in range(world):
            a, b = ret[(rank, False)], ret[(rank, True)]
            assert a["keys"] == b["keys"] and any(x > 0 for x in a["loss"])
            assert a["loss"] == b["loss"], (a["loss"], b["loss"])                           # same rows, same maths
            for ga, gb in zip(a["grad"], b["grad"]):
                assert (ga is None) == (gb is None) and (ga is None or np.array_equal(ga, gb))
            for ca, cb in zip(a["bank"], b["bank"]):
                assert np.array_equal(ca, cb)
        assert ret[(0, True)]["owned"] == [c % 2 == 0 for c in range(len(ret[(0, True)]["owned"]))]
        for c in range(len(ret[(0, True)]["bank"])):
            assert np.array_equal(ret[(0, True)]["bank"][c], ret[(1, True)]["bank"][c])     # every rank sees the same bank
