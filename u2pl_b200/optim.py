"""Fused parameter update of one training step (csrc/sgd_ema.cu): torch.optim.SGD (momentum, weight decay, per-group LR as
written by the scheduler of lr_helper.py:78-113) + the EMA of the teacher parameters (train_semi.py:531-548) in ONE kernel
over all ~360 parameter tensors.  State stays where torch keeps it -- `optimizer.state[p]["momentum_buffer"]`,
`param_group["lr"]` -- so `optimizer.state_dict()` / checkpoints and a later plain `optimizer.step()` remain valid."""
import struct

import numpy as np
import torch

from . import _lib
from .ops import _p, _stream


def _dense(t):
    """Non-overlapping and dense: numel elements laid out without holes in some dimension order."""
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)) or t.numel() <= 1


class FusedSGDEMA:
    def __init__(self, optimizer, student_params=None, teacher_params=None):
        assert isinstance(optimizer, torch.optim.SGD), "fused update is built for torch.optim.SGD (lr_helper.get_optimizer 'SGD')"
        for g in optimizer.param_groups:
            assert not g.get("nesterov", False) and g.get("dampening", 0) == 0 and not g.get("maximize", False)
            assert g.get("momentum", 0) > 0, "momentum == 0 has no buffer; use optimizer.step()"
        self.opt = optimizer
        self.pairs = None
        if student_params is not None:
            self.pairs = {id(s): t for s, t in zip(student_params, teacher_params)}
        lib = _lib.load()
        self.rec = int(lib.u2pl_sgd_tensor_bytes())
        self.chunk = int(lib.u2pl_sgd_chunk_elems())
        assert self.rec == 56
        self._slots = [dict(pin=None, dev=None, event=None), dict(pin=None, dev=None, event=None)]
        self._turn = 0

    @torch.no_grad()
    def step(self, ema_decay=None):
        """optimizer.step() (+ EMA of the paired teacher parameters with `ema_decay` when given)."""
        lib = _lib.load()
        recs, chunks, touched = [], [], []
        momentum = None
        dev = None
        for group in self.opt.param_groups:
            momentum = group["momentum"] if momentum is None else momentum
            assert group["momentum"] == momentum, "one momentum value per optimizer"
            for p in group["params"]:
                if p.grad is None:
                    continue
                # The update is element-wise, so any dense layout works as long as parameter, gradient, momentum buffer and
                # teacher parameter share it (the network's conv weights are channels-last: dense, not "contiguous").
                assert p.dtype == torch.float32 and _dense(p), "fused SGD needs dense fp32 parameters"
                g = p.grad
                if g.dtype != torch.float32 or g.stride() != p.stride():
                    g = torch.empty_like(p).copy_(g)              # empty_like keeps p's (dense) strides
                dev = p.device
                state = self.opt.state[p]
                m = state.get("momentum_buffer")
                first = m is None
                if first or m.stride() != p.stride():
                    m = torch.empty_like(p) if first else torch.empty_like(p).copy_(m)
                    state["momentum_buffer"] = m
                t = self.pairs.get(id(p)) if (self.pairs is not None and ema_decay is not None) else None
                if t is not None:
                    assert t.dtype == torch.float32 and t.data_ptr() != p.data_ptr()
                    if t.stride() != p.stride():                  # re-lay the teacher tensor once, in place of the old storage
                        t.data = torch.empty_like(p).copy_(t.data)
                n = p.numel()
                touched.append(p)
                if t is not None:
                    touched.append(t)
                idx = len(recs)
                recs.append(struct.pack("<QQQQqffii", p.data_ptr(), g.data_ptr(), m.data_ptr(), t.data_ptr() if t is not None else 0,
                                        n, float(group["lr"]), float(group["weight_decay"]), int(first), 0))
                chunks.extend((idx, c) for c in range((n + self.chunk - 1) // self.chunk))
        if not recs:
            return
        table = np.frombuffer(b"".join(recs), dtype=np.uint8)
        ch = np.asarray(chunks, dtype=np.uint32).reshape(-1)
        nbytes = table.size + ch.size * 4
        # two pinned staging buffers used alternately: the async H2D copy of step k may still be queued when the host
        # builds the tables of step k+1 (the host runs ahead of the GPU), so a buffer is rewritten only after ITS copy ran
        slot = self._slots[self._turn]
        self._turn ^= 1
        if slot["pin"] is None or slot["pin"].numel() < nbytes:
            slot["pin"] = torch.empty(nbytes * 2, dtype=torch.uint8)
            if dev.type == "cuda":                            # (CPU tensors only ever reach here under the tests' emulated ABI)
                slot["pin"] = slot["pin"].pin_memory()
            slot["dev"] = torch.empty(nbytes * 2, dtype=torch.uint8, device=dev)
            slot["event"] = None
        if slot["event"] is not None:
            slot["event"].synchronize()
        pin, dbuf = slot["pin"], slot["dev"]
        pin[:table.size].copy_(torch.from_numpy(table.copy()))
        pin[table.size:nbytes].copy_(torch.from_numpy(ch.view(np.uint8).copy()))
        dbuf[:nbytes].copy_(pin[:nbytes], non_blocking=True)
        if dev.type == "cuda":
            slot["event"] = torch.cuda.Event()
            slot["event"].record()
        rc = lib.u2pl_sgd_ema_step(_p(dbuf), _p(dbuf[table.size:]), len(chunks), float(momentum),
                                   float(ema_decay if ema_decay is not None else 0.0),
                                   float(1 - ema_decay) if ema_decay is not None else 1.0, int(ema_decay is not None), _stream())
        _lib.check(rc, "u2pl_sgd_ema_step")
        # the kernel wrote through raw pointers: move the version counters so that caches keyed on them (ops.kernel_weight)
        # and autograd's saved-tensor checks see the update
        torch.autograd.graph.increment_version(touched)
