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
