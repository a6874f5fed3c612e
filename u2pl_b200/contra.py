"""compute_contra_memobank_loss on the GPU (reference: u2pl/utils/loss_helper.py:51-235,
u2pl/utils/utils.py:16-47).

Control flow on the host, data on the device:

  device  onehot->bits, classify+scan, prototypes                      (u2pl_contra_* kernels)
  host    ONE device->host read of the 3*C member counts (the reference syncs >= 2*C times and
          pickles every class through all_gather_object); torch.randint draws on the CPU default
          generator in exactly the reference's order (:179-181,194-196) so the sampled index
          sets -- hence the loss -- are identical; FIFO bookkeeping of the banks (bank.py)
  device  pack keys -> (all_gather across ranks) -> ring append, InfoNCE forward (+ gradient
          rows), scatter-add backward

World size > 1: one all_gather of the [3*C] counts and one all_gather of the padded packed keys
replace the reference's C x (barrier + all_gather_object); rank order is preserved so every
rank's bank stays identical to the reference's concatenation order (utils.py:21-38).
"""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .bank import DeviceBank, ShardedBank, physical_rows, plan_append
from .ops import _need_cuda, _p, _stream

_BANKS = []          # [(memobank list object, DeviceBank)]; identity-keyed like the driver's own lists


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def bank_for(memobank, queue_size, dim, device):
    """Device bank behind the driver-owned `memobank` list (train_semi.py:161-169)."""
    for obj, bank in _BANKS:
        if obj is memobank:
            return bank
    rank, world = _world()
    if world > 1 and os.environ.get("U2PL_BANK_SHARDED", "0") == "1":   # opt-in until validated on a multi-GPU box
        bank = ShardedBank(queue_size, dim, device, rank, world)
    else:
        bank = DeviceBank(queue_size, dim, device)
    for c, slot in enumerate(memobank):
        if slot[0].shape[0] > 0:                       # caller pre-filled the CPU bank: adopt it
            bank.load(c, slot[0])
    _BANKS.append((memobank, bank))
    return bank


def forget_banks():
    del _BANKS[:]


def _strides(t):
    """(sn, sd, sp) element strides of a [N, D, h, w] feature tensor, or None if pixels are not
    uniformly strided."""
    N, D, h, w = t.shape
    sn, sd, sh, sw = t.stride()
    if sh != w * sw:
        return None
    return sn, sd, sw


def _feat(t):
    t = t if t.dtype == torch.float32 else t.float()
    if _strides(t) is None:
        t = t.contiguous()
    return t


_PIN = {}


def _to_device_i32(name, arr, device):
    """numpy int32 array -> device tensor through a reused pinned staging buffer."""
    a = np.ascontiguousarray(arr, dtype=np.int32).ravel()
    buf = _PIN.get(name)
    if buf is None or buf.numel() < a.size:
        buf = torch.empty(max(a.size, 1024), dtype=torch.int32).pin_memory()
        _PIN[name] = buf
    buf[:a.size].copy_(torch.from_numpy(a))
    return buf[:a.size].to(device, non_blocking=True)


class _InfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rep, plan):
        lib = _lib.load()
        sn, sd, sp = _strides(rep)
        N2, D, h, w = rep.shape
        nact, nq, nneg = plan["nact"], plan["nq"], plan["nneg"]
        dev = rep.device
        loss_q = torch.empty(nact * nq, dtype=torch.float32, device=dev)
        grad_rows = torch.empty((nact * nq, D), dtype=torch.float32, device=dev)
        anchor_pix = torch.empty(nact * nq, dtype=torch.int32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        sharded = "class_bank" in plan
        fn = lib.u2pl_infonce_forward_sharded if sharded else lib.u2pl_infonce_forward
        rc = fn(_p(rep), sn, sd, sp, N2 * h * w, D, h * w,
                                      _p(plan["an_bits"]), _p(plan["blockoff_an"]),
                                      _p(plan["act_class"]), _p(plan["a_ord"]), _p(plan["neg_rows"]),
                                      _p(plan["proto"]), _p(plan["class_bank"] if sharded else plan["bank_rows"]),
                                      nact, nq, nneg, float(plan["temp"]), int(plan["valid_seg"]),
                                      _p(loss_q), _p(grad_rows), _p(anchor_pix), _p(loss), _stream())
        _lib.check(rc, "u2pl_infonce_forward")
        ctx.save_for_backward(grad_rows, anchor_pix)
        ctx.meta = (tuple(rep.shape), rep.stride(), (sn, sd, sp))
        plan["loss_q"] = loss_q
        plan["anchor_pix"] = anchor_pix
        return loss

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        grad_rows, anchor_pix = ctx.saved_tensors
        shape, stride, (sn, sd, sp) = ctx.meta
        grad = torch.empty_strided(shape, stride, dtype=torch.float32, device=grad_rows.device).zero_()
        up = gout.to(torch.float32).reshape(1).contiguous()
        rc = lib.u2pl_infonce_backward(_p(grad_rows), _p(anchor_pix), grad_rows.shape[0], shape[1],
                                       shape[2] * shape[3], sn, sd, sp, _p(up), _p(grad), _stream())
        _lib.check(rc, "u2pl_infonce_backward")
        return grad, None


@torch.no_grad()
def _prepare(rep, label_bits, prob_l, prob_u, low_mask, high_mask, cfg, memobank, queue_prtlis, queue_size,
             rep_teacher):
    """Everything of loss_helper.py:80-222 that carries no gradient."""
    lib = _lib.load()
    dev = rep.device
    N2, D, h, w = rep.shape
    hw = h * w
    Bl, C = prob_l.shape[0], prob_l.shape[1]
    Bu = prob_u.shape[0]
    P = N2 * hw
    assert Bl + Bu == N2
    nq, nneg = int(cfg["num_queries"]), int(cfg["num_negatives"])
    nb = int(lib.u2pl_contra_num_blocks(P))
    bits3 = torch.empty((3, P), dtype=torch.int32, device=dev)
    blockcnt = torch.empty((3, C, nb), dtype=torch.int32, device=dev)
    blockoff = torch.empty((3, C, nb), dtype=torch.int32, device=dev)
    totals = torch.empty((3, C), dtype=torch.int32, device=dev)
    prob_l = prob_l.contiguous().float()
    prob_u = prob_u.contiguous().float()
    low_mask = low_mask.contiguous().float()
    high_mask = high_mask.contiguous().float()
    rc = lib.u2pl_contra_classify(_p(label_bits), _p(prob_l), _p(prob_u), _p(low_mask), _p(high_mask),
                                  Bl, Bu, C, hw, float(cfg["current_class_threshold"]),
                                  float(cfg["current_class_negative_threshold"]),
                                  int(cfg["low_rank"]), int(cfg["high_rank"]),
                                  _p(bits3), _p(blockcnt), _p(blockoff), _p(totals), _stream())
    _lib.check(rc, "u2pl_contra_classify")
    # prototypes (:119-123) -- device only, overlaps with the count read-back below
    tn, td, tp = _strides(rep_teacher)
    parts = int(lib.u2pl_contra_proto_parts())
    partial = torch.empty((parts, C, D), dtype=torch.float32, device=dev)
    proto = torch.empty((C, D), dtype=torch.float32, device=dev)
    rc = lib.u2pl_contra_proto(_p(rep_teacher), tn, td, tp, P, C, D, hw, _p(bits3[0]), _p(totals[0]),
                                _p(partial), _p(proto), _stream())
    _lib.check(rc, "u2pl_contra_proto")

    rank, world = _world()
    if world > 1:                                       # one collective instead of C barriers + object gathers
        all_tot = torch.empty((world, 3, C), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_tot.view(-1), totals.view(-1))   # flat: the same call is valid on NCCL and gloo
    else:
        all_tot = totals[None]
    tot = all_tot.cpu().numpy().astype(np.int64)        # the step's single device->host sync on this path
    mine = tot[rank]
    neg_counts = tot[:, 2, :]                           # [world, C]
    class_base = np.concatenate([np.zeros((world, 1), np.int64), np.cumsum(neg_counts, axis=1)[:, :-1]], axis=1)
    k_rank = neg_counts.sum(axis=1)                     # keys per rank
    kmax = int(k_rank.max())

    bank = bank_for(memobank, queue_size, D, dev)
    new_keys = [int(neg_counts[:, c].sum()) for c in range(C)]          # :143-150 return values
    if kmax > 0:
        packed = torch.empty((kmax, D), dtype=torch.float32, device=dev)
        if k_rank[rank] > 0:
            cb = _to_device_i32("class_base", class_base[rank], dev)
            rc = lib.u2pl_contra_pack_keys(_p(rep_teacher), tn, td, tp, P, C, D, hw, _p(bits3[2]),
                                           _p(blockoff[2]), _p(cb), _p(packed), _stream())
            _lib.check(rc, "u2pl_contra_pack_keys")
        if world > 1:
            gathered = torch.empty((world * kmax, D), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(gathered, packed)
        else:
            gathered = packed
        descs = []
        sharded = isinstance(bank, ShardedBank)
        for c in range(C):                              # ring state advances for every class on every rank ...
            d, _ = plan_append(bank.rings[c], neg_counts[:, c], [r * kmax + class_base[r, c] for r in range(world)])
            if not sharded or bank.owns(c):             # ... the rows are written by the owner only
                descs.extend(d)
        if descs:
            dd = _to_device_i32("append_desc", np.asarray(descs, dtype=np.int64), dev)
            rc = lib.u2pl_bank_append(_p(gathered), _p(bank.rows), D, _p(dd), len(descs),
                                      int(max(x[4] for x in descs)), _stream())
            _lib.check(rc, "u2pl_bank_append")
    if isinstance(bank, ShardedBank):
        bank.fence()                                    # peers may read my shard only after this step's append
    for c in range(C):                                  # keep the driver-visible state coherent
        queue_prtlis[c][0] = bank.rings[c].ptr          # utils.py:45
        n = bank.rings[c].length
        if memobank[c][0].shape[0] != n:
            memobank[c][0] = torch.zeros(1, D).expand(n, D)   # shape proxy; the rows live in bank.rows

    valid_classes = [c for c in range(C) if mine[0, c] > 0]             # :152-154
    valid_seg = len(valid_classes)
    plan = dict(new_keys=new_keys, valid_classes=valid_classes, valid_seg=valid_seg, nact=0, nq=nq, nneg=nneg,
                temp=cfg["temperature"], proto=proto, bank_rows=bank.rows, an_bits=bits3[1],
                blockoff_an=blockoff[1], totals=mine, bits3=bits3, blockoff=blockoff)
    if valid_seg <= 1:                                                  # :156-162
        return plan
    act, a_ord, n_rows, sampled = [], [], [], []
    for j in range(valid_seg):                                          # :173
        n_anchor = int(mine[1, j])                  # Q1: list position j, not valid_classes[j]
        ring = bank.rings[valid_classes[j]]
        if not (n_anchor > 0 and ring.length > 0):                      # :174-188
            continue
        a_idx = torch.randint(n_anchor, size=(nq,))                     # :179-181  CPU default generator
        n_idx = torch.randint(ring.length, size=(nq * nneg,))           # :194-196
        act.append(j)
        a_ord.append(a_idx.numpy())
        n_rows.append(physical_rows(ring, n_idx.numpy()))
        sampled.append(dict(j=j, cls=valid_classes[j], a_idx=a_idx.numpy(), n_idx=n_idx.numpy()))
    plan["sampled"] = sampled
    plan["nact"] = len(act)
    if act:
        plan["act_class"] = _to_device_i32("act_class", np.asarray(act), dev)
        plan["a_ord"] = _to_device_i32("a_ord", np.stack(a_ord), dev)
        plan["neg_rows"] = _to_device_i32("neg_rows", np.stack(n_rows), dev)
        if isinstance(bank, ShardedBank):               # base of the shard holding each active class (maybe a peer's)
            ptrs = np.asarray([bank.class_base(valid_classes[j]) for j in act], dtype=np.int64)
            plan["class_bank"] = torch.from_numpy(ptrs).to(dev)
    return plan


def label_bits_from_onehot(label_l, label_u):
    """[Bl,C,h,w] + [Bu,C,h,w] int64 (multi-)hot -> uint32 class bitmask per pixel."""
    lib = _lib.load()
    Bl, C, h, w = label_l.shape
    Bu = label_u.shape[0]
    bits = torch.empty((Bl + Bu) * h * w, dtype=torch.int32, device=label_l.device)
    for lab, off, B in ((label_l, 0, Bl), (label_u, Bl * h * w, Bu)):
        lab = lab.contiguous()
        if lab.dtype != torch.int64:
            lab = lab.long()
        rc = lib.u2pl_onehot_to_bits(_p(lab), B, C, h * w, _p(bits[off:]), _stream())
        _lib.check(rc, "u2pl_onehot_to_bits")
    return bits


def compute_contra_memobank_loss(rep, label_l, label_u, prob_l, prob_u, low_mask, high_mask, cfg,
                                 memobank, queue_prtlis, queue_size, rep_teacher,
                                 momentum_prototype=None, i_iter=0, label_bits=None, return_plan=False):
    """Same positional signature and return values as loss_helper.py:51-66,232-235.
    `label_bits` (optional) lets the fused step pass class bitmasks directly instead of one-hots."""
    if momentum_prototype is not None:
        raise NotImplementedError("anchor_ema / momentum_prototype (loss_helper.py:209-218) is not built yet; "
                                  "no shipped config sets trainer.contrastive.anchor_ema")
    _need_cuda(rep, prob_l, prob_u, low_mask, high_mask, rep_teacher)
    rep_f = _feat(rep)
    rep_teacher = _feat(rep_teacher.detach())
    if label_bits is None:
        label_bits = label_bits_from_onehot(label_l, label_u)
    plan = _prepare(rep_f.detach(), label_bits, prob_l.detach(), prob_u.detach(), low_mask, high_mask, cfg,
                    memobank, queue_prtlis, queue_size, rep_teacher)
    if plan["valid_seg"] <= 1 or plan["nact"] == 0:
        loss = rep[0, 0, 0, 0] * 0.0                      # :160,187  zero with a graph through rep (the reference's
                                                          # `0 * rep.sum()` without the 545 MB reduction)
    else:
        loss = _InfoNCE.apply(rep_f, plan)
    if return_plan:
        return plan["new_keys"], loss, plan
    return plan["new_keys"], loss
