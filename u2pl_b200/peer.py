"""Small-vector all-reduce over peer-mapped memory (csrc/peer_reduce.cu) for the SyncBatchNorm statistics exchange
(reference base.py:6-8).  One `PeerReducer` per process: an exchange region on this GPU (cudaMalloc'd, exported through
CUDA IPC) plus the mapped regions of every other rank of the box.  `allreduce_(t)` sums a contiguous fp32 tensor of at
most 4096 elements in place across ranks with ONE single-CTA kernel; larger tensors, or boxes where peer mapping is not
possible, use torch.distributed."""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib
from .ops import _p, _stream

_REDUCER = None
_DISABLED = False


class PeerReducer:
    def __init__(self, device, group=None):
        lib = _lib.load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device, self.group = device, group
        if self.world > 8:
            raise RuntimeError("peer all-reduce is built for one NVSwitch box (<= 8 ranks)")
        nbytes = int(lib.u2pl_peer_region_bytes())
        ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        _lib.check(lib.u2pl_shard_alloc(nbytes, ctypes.byref(ptr), handle), "u2pl_shard_alloc")      # zero-filled
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        bases = []
        for r in range(self.world):
            if r == self.rank:
                bases.append(int(ptr.value))
                continue
            peer = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
            _lib.check(lib.u2pl_shard_open(buf, ctypes.byref(peer)), "u2pl_shard_open")
            bases.append(int(peer.value))
        self.bases = (ctypes.c_void_p * self.world)(*bases)
        self.max_floats = int(lib.u2pl_peer_max_floats())
        self.seq = 0
        dist.barrier(group=group)                            # every region is mapped before anybody pushes

    def allreduce_(self, t):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() <= self.max_floats
        self.seq += 1
        rc = _lib.load().u2pl_peer_allreduce_f32(_p(t), t.numel(), self.bases, self.rank, self.world, self.seq, _stream())
        _lib.check(rc, "u2pl_peer_allreduce_f32")
        return t


def allreduce_small_(t, group=None):
    """Sum `t` across ranks in place: peer-memory kernel when possible (U2PL_PEER_SYNCBN != 0, default group, fp32,
    <= 4096 elements), else torch.distributed."""
    global _REDUCER, _DISABLED
    if (not _DISABLED and group is None and os.environ.get("U2PL_PEER_SYNCBN", "1") == "1" and t.is_cuda
            and t.dtype == torch.float32 and t.is_contiguous() and dist.get_backend() == "nccl"):
        if _REDUCER is None:
            try:
                _REDUCER = PeerReducer(t.device)
            except Exception as ex:                          # e.g. ranks on different nodes: keep NCCL, say so once
                _DISABLED = True
                if dist.get_rank() == 0:
                    print(f"[u2pl_b200] peer all-reduce unavailable ({ex!r}); SyncBN statistics go through NCCL")
        if _REDUCER is not None and t.numel() <= _REDUCER.max_floats:
            return _REDUCER.allreduce_(t)
    dist.all_reduce(t, group=group)
    return t
