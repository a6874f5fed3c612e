"""Fused parameter update of one training step (csrc/sgd_ema.cu): torch.optim.SGD (momentum, weight decay, per-group LR as
written by the scheduler of lr_helper.py:78-113) + the EMA of the teacher parameters (train_semi.py:531-548) in ONE kernel
over all ~360 parameter tensors.  State stays where torch keeps it -- `optimizer.state[p]["momentum_buffer"]`,
`param_group["lr"]` -- so `optimizer.state_dict()` / checkpoints and a later plain `optimizer.step()` remain valid."""
import struct

import numpy as np
import torch

from . import _lib
from .ops import _p, _stream


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
        self._pin = None
        self._dev = None

    @torch.no_grad()
    def step(self, ema_decay=None):
        """optimizer.step() (+ EMA of the paired teacher parameters with `ema_decay` when given)."""
        lib = _lib.load()
        recs, chunks = [], []
        momentum = None
        dev = None
        for group in self.opt.param_groups:
            momentum = group["momentum"] if momentum is None else momentum
            assert group["momentum"] == momentum, "one momentum value per optimizer"
            for p in group["params"]:
                if p.grad is None:
                    continue
                assert p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous() and p.grad.dtype == torch.float32
                dev = p.device
                state = self.opt.state[p]
                first = "momentum_buffer" not in state or state["momentum_buffer"] is None
                if first:
                    state["momentum_buffer"] = torch.empty_like(p)
                m = state["momentum_buffer"]
                t = self.pairs.get(id(p)) if (self.pairs is not None and ema_decay is not None) else None
                if t is not None:
                    assert t.is_contiguous() and t.dtype == torch.float32 and t.data_ptr() != p.data_ptr()
                n = p.numel()
                idx = len(recs)
                recs.append(struct.pack("<QQQQqffii", p.data_ptr(), p.grad.data_ptr(), m.data_ptr(), t.data_ptr() if t is not None else 0,
                                        n, float(group["lr"]), float(group["weight_decay"]), int(first), 0))
                chunks.extend((idx, c) for c in range((n + self.chunk - 1) // self.chunk))
        if not recs:
            return
        table = np.frombuffer(b"".join(recs), dtype=np.uint8)
        ch = np.asarray(chunks, dtype=np.uint32).reshape(-1)
        nbytes = table.size + ch.size * 4
        if self._pin is None or self._pin.numel() < nbytes:
            self._pin = torch.empty(nbytes * 2, dtype=torch.uint8)
            if dev.type == "cuda":                            # (CPU tensors only ever reach here under the tests' emulated ABI)
                self._pin = self._pin.pin_memory()
            self._dev = torch.empty(nbytes * 2, dtype=torch.uint8, device=dev)
        self._pin[:table.size].copy_(torch.from_numpy(table.copy()))
        self._pin[table.size:nbytes].copy_(torch.from_numpy(ch.view(np.uint8).copy()))
        self._dev[:nbytes].copy_(self._pin[:nbytes], non_blocking=True)
        rc = lib.u2pl_sgd_ema_step(_p(self._dev), _p(self._dev[table.size:]), len(chunks), float(momentum),
                                   float(ema_decay if ema_decay is not None else 0.0), int(ema_decay is not None), _stream())
        _lib.check(rc, "u2pl_sgd_ema_step")
