"""CPU: the FIFO bookkeeping of u2pl_b200/bank.py against the oracle's dequeue_and_enqueue
(utils.py:28-47 restated), on a numpy-simulated ring, incl. multi-rank segments and overflow."""
import numpy as np

from oracle import port
from u2pl_b200.bank import Ring, physical_rows, plan_append


def _simulate(cap, steps, world, rng):
    D = 3
    ring = Ring(cap=cap, row_base=5)
    store = np.full((5 + cap, D), np.nan, np.float32)
    q, p = [np.zeros((0, D), np.float32)], [0]
    nxt = 0
    for _ in range(steps):
        counts = [int(rng.integers(0, cap + 3)) if rng.random() < 0.8 else 0 for _ in range(world)]
        segs = []
        for k in counts:
            segs.append((np.arange(nxt, nxt + k, dtype=np.float32)[:, None] + np.zeros((1, D), np.float32)))
            nxt += k
        gathered = np.concatenate(segs + [np.zeros((0, D), np.float32)], axis=0)
        firsts = np.concatenate([[0], np.cumsum(counts)[:-1]])
        descs, k = plan_append(ring, counts, firsts)
        for (src, base, first, c, cnt) in descs:
            for i in range(cnt):
                store[base + (first + i) % c] = gathered[src + i]
        n = port.dequeue_and_enqueue(segs, q, p, cap)
        assert n == k
        assert ring.length == q[0].shape[0] and ring.ptr == p[0]
        got = store[physical_rows(ring, np.arange(ring.length))]
        assert np.array_equal(got, q[0])


def test_ring_matches_reference_fifo():
    rng = np.random.default_rng(0)
    for cap in (1, 2, 5, 17):
        for world in (1, 2, 4):
            _simulate(cap, 40, world, rng)


def test_exact_fill_sets_ptr_to_capacity():
    ring = Ring(cap=4, row_base=0)
    plan_append(ring, [3], [0])
    assert (ring.length, ring.ptr) == (3, 3)
    plan_append(ring, [1], [0])
    assert (ring.length, ring.ptr) == (4, 4)          # `>= queue_size` branch (utils.py:39-41)
    plan_append(ring, [0], [0])
    assert (ring.length, ring.ptr) == (4, 4)
