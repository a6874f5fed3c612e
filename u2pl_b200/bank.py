"""Device-resident class-wise memory bank (reference: train_semi.py:161-169 builds
`memobank` / `queue_ptrlis` / `queue_size` as CPU lists; utils.py:28-47 appends to them).

The reference keeps every bank as a CPU tensor, re-allocates it on every append and copies the
whole bank host->device once per class per step (loss_helper.py:192, up to 665 MB/step).  Here
all C banks live in ONE device tensor [sum(queue_size), D]; each class is a ring buffer
(start, length) over its own row range, so `keep the newest queue_size rows` is pointer
arithmetic and an append only writes the new rows.

This file is host bookkeeping only (pure Python/numpy, unit-tested on CPU against the oracle's
dequeue_and_enqueue); the row copies are u2pl_bank_append (csrc/contra.cu).
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class Ring:
    cap: int            # queue_size[c]
    row_base: int       # first row of this class inside the shared bank tensor
    start: int = 0      # ring position of logical row 0
    length: int = 0     # rows held (== memobank[c][0].shape[0] in the reference)
    ptr: int = 0        # queue_ptrlis[c][0] in the reference (utils.py:36-45)


def plan_append(ring, seg_counts, seg_src_first):
    """FIFO arithmetic of utils.py:38-43 for one class.

    seg_counts[r]    : keys contributed by rank r (rank order = reference's concat order, :21-32)
    seg_src_first[r] : row of rank r's first key for this class inside the gathered key buffer
    Returns (descriptors, total_new) and updates `ring`.  A descriptor is
    (src_first_row, dst_row_base, dst_first_pos, capacity, count): source row i goes to
    bank[dst_row_base + (dst_first_pos + i) % capacity]."""
    k = int(sum(seg_counts))
    total = ring.length + k
    if total < ring.cap:                               # :42-43
        drop_old, skip, new_len = 0, 0, total
        ring.ptr = (ring.ptr + k) % ring.cap
    else:                                              # :39-41  queue[0][-queue_size:]
        drop_total = total - ring.cap
        drop_old = min(drop_total, ring.length)
        skip = drop_total - drop_old                   # oldest NEW rows that never make it in
        new_len = ring.cap
        ring.ptr = ring.cap
    new_start = (ring.start + drop_old) % ring.cap
    kept_old = ring.length - drop_old
    descs = []
    s = 0
    for k_r, src in zip(seg_counts, seg_src_first):
        k_r = int(k_r)
        lo = max(s, skip)
        if lo < s + k_r:
            descs.append((int(src) + (lo - s), ring.row_base, (new_start + kept_old + (lo - skip)) % ring.cap,
                          ring.cap, s + k_r - lo))
        s += k_r
    ring.start, ring.length = new_start, new_len
    return descs, k


def physical_rows(ring, logical_idx):
    """Bank-tensor rows of logical indices (what `negative_feat[high_entropy_idx]` addresses, :197)."""
    idx = np.asarray(logical_idx, dtype=np.int64)
    return (ring.row_base + (ring.start + idx) % ring.cap).astype(np.int32)


def shard_layout(queue_size, world):
    """Class-sharded placement: class c lives on rank c % world.  Returns (owner[c], row_base[c] inside
    the owner's shard, rows per shard)."""
    owner, row_base, rows = [], [], [0] * world
    for c, cap in enumerate(queue_size):
        r = c % world
        owner.append(r)
        row_base.append(rows[r])
        rows[r] += int(cap)
    return owner, row_base, rows


class DeviceBank:
    """All class banks in one device tensor + per-class Ring state."""

    def __init__(self, queue_size, dim, device):
        import torch
        self.dim = int(dim)
        self.rings = []
        base = 0
        for cap in queue_size:
            self.rings.append(Ring(cap=int(cap), row_base=base))
            base += int(cap)
        self.rows = torch.zeros((base, self.dim), dtype=torch.float32, device=device)

    def length(self, c):
        return self.rings[c].length

    def materialize(self, c):
        """Logical-order copy of bank c (what memobank[c][0] holds in the reference)."""
        import torch
        r = self.rings[c]
        idx = (r.start + torch.arange(r.length, device=self.rows.device)) % r.cap + r.row_base
        return self.rows[idx]

    def load(self, c, rows):
        """Adopt pre-existing rows (e.g. a bank the caller filled on the CPU)."""
        r = self.rings[c]
        n = min(int(rows.shape[0]), r.cap)
        self.rows[r.row_base:r.row_base + n] = rows[-n:].to(self.rows.device, self.rows.dtype)
        r.start, r.length = 0, n
        r.ptr = r.cap if rows.shape[0] >= r.cap else n % r.cap


class _DevicePtrView:
    """Zero-copy torch view of raw device memory (own cudaMalloc shard or a peer's IPC mapping)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def _tensor_from_ptr(ptr, shape, device=None):
    """Zero-copy fp32 tensor over raw device memory.  With `device` (the pointer's own device) nothing is copied; without it
    torch places the tensor on whatever device owns the pointer (a mapped peer shard lives on the peer's device)."""
    import torch
    view = _DevicePtrView(ptr, shape)
    return torch.as_tensor(view, device=device) if device is not None else torch.as_tensor(view)


class ShardedBank:
    """Banks sharded BY CLASS over the GPUs of one box: rank c % world holds class c's ring in a cudaMalloc'd shard
    that every other rank maps through CUDA IPC.  Ring bookkeeping (host integers) is replicated for all classes --
    it is a pure function of the all-gathered key counts -- so every rank can draw `torch.randint(len(bank))` and
    address the rows exactly as the reference does (loss_helper.py:192-197) without asking the owner anything; the
    InfoNCE kernel then reads the sampled rows directly from the owner's memory over NVLink
    (u2pl_infonce_forward_sharded).  Appends touch only the owner's shard (1/world of the replicated traffic).

    Ordering between ranks (all on each rank's compute stream):
      append(step n) -> tiny all_reduce -> peers' loss kernels read          (read-after-write)
      peers' loss kernels(step n) -> counts all_gather(step n+1) -> append   (write-after-read; contra._prepare)
    """

    def __init__(self, queue_size, dim, device, rank, world, group=None):
        import ctypes

        import torch
        import torch.distributed as dist

        from . import _lib
        lib = _lib.load()
        self.dim, self.rank, self.world, self.group = int(dim), int(rank), int(world), group
        self.owner, bases, shard_rows = shard_layout(queue_size, world)
        self.rings = [Ring(cap=int(cap), row_base=b) for cap, b in zip(queue_size, bases)]
        self.shard_rows = shard_rows
        nbytes = max(shard_rows[rank], 1) * self.dim * 4
        ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        _lib.check(lib.u2pl_shard_alloc(nbytes, ctypes.byref(ptr), handle), "u2pl_shard_alloc")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.base = []
        for r in range(world):
            if r == rank:
                self.base.append(int(ptr.value))
                continue
            peer = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
            _lib.check(lib.u2pl_shard_open(buf, ctypes.byref(peer)), "u2pl_shard_open")
            self.base.append(int(peer.value))
        self.device = device
        # the local shard as a tensor (what u2pl_bank_append writes; same device, so as_tensor does not copy)
        self.rows = _tensor_from_ptr(self.base[rank], (max(shard_rows[rank], 1), self.dim), device)
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)

    def owns(self, c):
        return self.owner[c] == self.rank

    def length(self, c):
        return self.rings[c].length

    def fence(self):
        """Orders this rank's appends before any peer's subsequent reads (see class docstring)."""
        import torch.distributed as dist
        dist.all_reduce(self._flag, group=self.group)

    def class_base(self, c):
        return self.base[self.owner[c]]

    def materialize(self, c):
        """Logical-order copy of bank c on this rank's device (debug / tests; a peer's shard is copied over NVLink)."""
        import torch
        r, o = self.rings[c], self.owner[c]
        if o == self.rank:
            shard = self.rows
        else:       # torch places a tensor built from a mapped peer pointer on the peer's device: copy the ring's rows over
            peer = _tensor_from_ptr(self.base[o], (max(self.shard_rows[o], 1), self.dim))
            shard = torch.empty((r.cap, self.dim), dtype=torch.float32, device=self.device)
            shard.copy_(peer[r.row_base:r.row_base + r.cap])
            idx = (r.start + torch.arange(r.length, device=self.device)) % r.cap
            return shard[idx]
        idx = (r.start + torch.arange(r.length, device=self.device)) % r.cap + r.row_base
        return shard[idx]

    def load(self, c, rows):
        """Adopt pre-existing rows: every rank updates the ring state, only the owner stores the rows."""
        r = self.rings[c]
        n = min(int(rows.shape[0]), r.cap)
        if self.owns(c):
            self.rows[r.row_base:r.row_base + n] = rows[-n:].to(self.rows.device, self.rows.dtype)
        r.start, r.length = 0, n
        r.ptr = r.cap if rows.shape[0] >= r.cap else n % r.cap

    def close(self):
        from . import _lib
        lib = _lib.load()
        for r, b in enumerate(self.base):
            lib.u2pl_shard_close(b, int(r == self.rank))
        self.base = []
