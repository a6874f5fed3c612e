"""CPU: u2pl_b200/contra.py END TO END (host control flow, ring-bank bookkeeping, sampling order, plan construction,
autograd wrapper) with the contrastive C-ABI entry points EMULATED inside this test from their header contracts, against
the oracle and the reference's own fixtures (tests/golden/contra_*).  The real kernels are compared with the same oracle
on the GPU (tests/test_gpu_contra.py); this file keeps the layer above them under test where there is no GPU.
Test-only emulation: the product has no CPU path."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import port
from u2pl_b200 import _lib, contra, ops

BLK = 256


def _addr(p):
    return p.value if isinstance(p, ctypes.c_void_p) else (int(p) if p is not None else None)


def _arr(p, n, ctype):
    return np.ctypeslib.as_array((ctype * int(n)).from_address(_addr(p)))


def _feat(p, sn, sd, sp, N, D, hw):
    """[N, hw, D] strided view of a feature tensor addressed as n*sn + d*sd + pixel*sp (elements)."""
    size = (N - 1) * sn + (D - 1) * sd + (hw - 1) * sp + 1
    flat = _arr(p, size, ctypes.c_float)
    return np.lib.stride_tricks.as_strided(flat, (N, hw, D), (4 * sn, 4 * sp, 4 * sd))


class FakeContraLib:
    def __init__(self):
        self.cfg = None                                  # set by the test: classify needs the config-independent args only

    def u2pl_last_error(self):
        return b""

    def u2pl_onehot_to_bits(self, onehot, B, C, hw, bits, stream):
        oh = _arr(onehot, B * C * hw, ctypes.c_int64).reshape(B, C, hw)
        out = _arr(bits, B * hw, ctypes.c_uint32).reshape(B, hw)
        out[:] = 0
        for c in range(C):
            out |= ((oh[:, c] != 0).astype(np.uint32) << np.uint32(c))
        return 0

    def u2pl_contra_num_blocks(self, P):
        return (P + BLK - 1) // BLK

    def u2pl_contra_classify(self, label_bits, prob_l, prob_u, low_mask, high_mask, Bl, Bu, C, hw, thr, nthr, low_rank, high_rank,
                             bits3, blockcnt, blockoff, totals, stream):
        P = (Bl + Bu) * hw
        lb = _arr(label_bits, P, ctypes.c_uint32).reshape(Bl + Bu, hw)
        onehot = np.stack([((lb >> np.uint32(c)) & 1).astype(np.int64) for c in range(C)], axis=1)      # [N, C, hw]
        pl = _arr(prob_l, Bl * C * hw, ctypes.c_float).reshape(Bl, C, hw)
        pu = _arr(prob_u, Bu * C * hw, ctypes.c_float).reshape(Bu, C, hw)
        lm = _arr(low_mask, P, ctypes.c_float).reshape(Bl + Bu, 1, hw)
        hm = _arr(high_mask, P, ctypes.c_float).reshape(Bl + Bu, 1, hw)
        cfg = dict(current_class_threshold=thr, current_class_negative_threshold=nthr, low_rank=low_rank, high_rank=high_rank)
        sel = port.contra_select(onehot[:Bl], onehot[Bl:], pl, pu, lm, hm, cfg)
        nb = (P + BLK - 1) // BLK
        b3 = _arr(bits3, 3 * P, ctypes.c_uint32).reshape(3, P)
        cnt = _arr(blockcnt, 3 * C * nb, ctypes.c_uint32).reshape(3, C, nb)
        off = _arr(blockoff, 3 * C * nb, ctypes.c_uint32).reshape(3, C, nb)
        tot = _arr(totals, 3 * C, ctypes.c_uint32).reshape(3, C)
        b3[:] = 0
        for k, name in enumerate(("lowvalid", "anchors", "negs")):
            for c in range(C):
                idx = sel[name][c]
                b3[k, idx] |= np.uint32(1 << c)
                per_block = np.bincount(idx // BLK, minlength=nb).astype(np.uint32)
                cnt[k, c] = per_block
                off[k, c] = np.concatenate([[0], np.cumsum(per_block)[:-1]]).astype(np.uint32)
                tot[k, c] = idx.size
        return 0

    def u2pl_contra_proto_parts(self):
        return 2

    def u2pl_contra_proto(self, rep_t, sn, sd, sp, P, C, D, hw, lv_bits, lv_totals, partial, proto, stream):
        rows = _feat(rep_t, sn, sd, sp, P // hw, D, hw).reshape(P, D)
        lv = _arr(lv_bits, P, ctypes.c_uint32)
        out = _arr(proto, C * D, ctypes.c_float).reshape(C, D)
        for c in range(C):
            idx = np.flatnonzero((lv >> np.uint32(c)) & 1)
            out[c] = rows[idx].mean(axis=0, dtype=np.float32) if idx.size else np.nan
        return 0

    def u2pl_contra_pack_keys(self, rep_t, sn, sd, sp, P, C, D, hw, ng_bits, blockoff_ng, class_base, packed, stream):
        rows = _feat(rep_t, sn, sd, sp, P // hw, D, hw).reshape(P, D)
        ng = _arr(ng_bits, P, ctypes.c_uint32)
        base = _arr(class_base, C, ctypes.c_uint32)
        total = int(base[-1]) + int(np.count_nonzero((ng >> np.uint32(C - 1)) & 1))
        out = _arr(packed, max(total, 1) * D, ctypes.c_float).reshape(-1, D)
        for c in range(C):
            idx = np.flatnonzero((ng >> np.uint32(c)) & 1)
            out[base[c]:base[c] + idx.size] = rows[idx]
        return 0

    def u2pl_bank_append(self, src_rows, bank, D, desc, ndesc, max_count, stream):
        d = _arr(desc, ndesc * 5, ctypes.c_uint32).reshape(ndesc, 5).astype(np.int64)
        for src, base, first, cap, count in d:
            s = _arr(src_rows, (src + count) * D, ctypes.c_float).reshape(-1, D)
            b = _arr(bank, (base + cap) * D, ctypes.c_float).reshape(-1, D)
            for r in range(count):
                b[base + (first + r) % cap] = s[src + r]
        return 0

    def u2pl_infonce_forward(self, rep, sn, sd, sp, P, D, hw, an_bits, blockoff_an, act_class, a_ord, neg_rows, proto, bank,
                             nact, nq, nneg, temp, valid_seg, loss_q, grad_rows, anchor_pix, loss, stream):
        rows = torch.from_numpy(np.ascontiguousarray(_feat(rep, sn, sd, sp, P // hw, D, hw).reshape(P, D)))
        an = _arr(an_bits, P, ctypes.c_uint32)
        act = _arr(act_class, nact, ctypes.c_int32)
        ao = _arr(a_ord, nact * nq, ctypes.c_int32).reshape(nact, nq)
        nr = _arr(neg_rows, nact * nq * nneg, ctypes.c_int32).reshape(nact, nq, nneg).astype(np.int64)
        C = int(act.max()) + 1
        pr = torch.from_numpy(_arr(proto, C * D, ctypes.c_float).reshape(C, D).copy())
        bk = torch.from_numpy(_arr(bank, (int(nr.max()) + 1) * D, ctypes.c_float).reshape(-1, D).copy())
        lq = _arr(loss_q, nact * nq, ctypes.c_float)
        gr = _arr(grad_rows, nact * nq * D, ctypes.c_float).reshape(nact * nq, D)
        ap = _arr(anchor_pix, nact * nq, ctypes.c_int32)
        total = torch.zeros(())
        scale = 1.0 / (nq * valid_seg)
        with torch.enable_grad():
            for a in range(nact):
                cls = int(act[a])
                members = np.flatnonzero((an >> np.uint32(cls)) & 1)
                pix = members[ao[a]]
                ap[a * nq:(a + 1) * nq] = pix
                anchor = rows[torch.from_numpy(pix)].clone().requires_grad_(True)
                keys = torch.cat((pr[cls].reshape(1, 1, D).repeat(nq, 1, 1), bk[torch.from_numpy(nr[a])]), dim=1)
                logits = torch.cosine_similarity(anchor.unsqueeze(1), keys, dim=2)
                ce = torch.nn.functional.cross_entropy(logits / temp, torch.zeros(nq, dtype=torch.long), reduction="none")
                (ce.sum() * scale).backward()
                lq[a * nq:(a + 1) * nq] = ce.detach().numpy()
                gr[a * nq:(a + 1) * nq] = anchor.grad.numpy()
                total = total + ce.detach().sum() * scale
        _arr(loss, 1, ctypes.c_float)[0] = float(total)
        return 0

    def u2pl_infonce_backward(self, grad_rows, anchor_pix, nrows, D, hw, sn, sd, sp, upstream, grad_rep, stream):
        gr = _arr(grad_rows, nrows * D, ctypes.c_float).reshape(nrows, D)
        ap = _arr(anchor_pix, nrows, ctypes.c_int32)
        up = float(_arr(upstream, 1, ctypes.c_float)[0])
        n_img = int(ap.max()) // hw + 1
        out = _feat(grad_rep, sn, sd, sp, n_img, D, hw)
        for r in range(nrows):
            out[ap[r] // hw, ap[r] % hw] += up * gr[r]
        return 0


@pytest.fixture
def emulated(monkeypatch):
    fake = FakeContraLib()
    monkeypatch.setattr(_lib, "load", lambda *a, **k: fake)
    monkeypatch.setattr(ops, "_need_cuda", lambda *ts: None)
    monkeypatch.setattr(contra, "_need_cuda", lambda *ts: None)
    monkeypatch.setattr(contra, "_stream", lambda: None)
    monkeypatch.setattr(contra, "_to_device_i32",
                        lambda name, arr, device: torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int32).ravel().copy()))
    yield fake
    contra.forget_banks()


def _cfg(g):
    cfg = {k: v for k, v in zip(g["cfg_keys"].tolist(), g["cfg_vals"].tolist())}
    for k in ("low_rank", "high_rank", "num_negatives", "num_queries"):
        cfg[k] = int(cfg[k])
    return cfg


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("name", ["contra_c21", "contra_c19_missing", "contra_c21_driver_onehot"])
def test_contra_host_logic_matches_oracle(golden, emulated, name, channels_last):
    g = golden(name)
    cfg = _cfg(g)
    C, D = g["s0_label_l"].shape[1], g["s0_rep"].shape[1]
    qsize = g["queue_size"].tolist()
    memobank = [[torch.zeros(0, D)] for _ in range(C)]
    ptrs = [torch.zeros(1, dtype=torch.long) for _ in range(C)]
    o_bank = [[np.zeros((0, D), np.float32)] for _ in range(C)]
    o_ptr = [[0] for _ in range(C)]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))                                 # noqa: E731
    for s in range(int(g["steps"])):
        rep, rep_t = t(g[f"s{s}_rep"]), t(g[f"s{s}_rep_teacher"])
        if channels_last:
            rep, rep_t = rep.contiguous(memory_format=torch.channels_last), rep_t.contiguous(memory_format=torch.channels_last)
        rep.requires_grad_(True)
        args = (g[f"s{s}_label_l"].astype(np.int64), g[f"s{s}_label_u"].astype(np.int64), g[f"s{s}_prob_l"],
                g[f"s{s}_prob_u"], g[f"s{s}_low_mask"].astype(np.float32), g[f"s{s}_high_mask"].astype(np.float32))
        torch.manual_seed(1000 + s)
        new_keys, loss, plan = contra.compute_contra_memobank_loss(rep, *[t(a) for a in args], cfg, memobank, ptrs, qsize, rep_t,
                                                                   return_plan=True)
        loss.backward()
        torch.manual_seed(1000 + s)
        out = port.compute_contra_memobank_loss(g[f"s{s}_rep"], *args, cfg, o_bank, o_ptr, qsize, g[f"s{s}_rep_teacher"], want_grad=True)
        assert new_keys == out["new_keys"] == g[f"s{s}_new_keys"].tolist()
        assert plan["valid_classes"] == out["valid_classes"]
        assert [int(p[0]) for p in ptrs] == [int(p[0]) for p in o_ptr] == g[f"s{s}_ptr"].tolist()
        assert [m[0].shape[0] for m in memobank] == g[f"s{s}_bank_len"].tolist()
        bank = contra.bank_for(memobank, qsize, D, rep.device)
        for c in range(C):
            assert np.array_equal(bank.materialize(c).numpy(), o_bank[c][0])                 # FIFO content
        if plan["nact"]:
            pix = plan["anchor_pix"].numpy().reshape(plan["nact"], -1)
            for a, smp in enumerate(out["sampled"]):
                assert np.array_equal(pix[a], out["sel"]["anchors"][smp["j"]][smp["a_idx"]])  # same anchors, same order of draws
        assert abs(float(loss.detach()) - float(out["loss"])) <= 1e-5
        grad = rep.grad.numpy() if rep.grad is not None else np.zeros_like(g[f"s{s}_rep"])
        assert np.abs(grad - out["rep_grad"]).max() <= 1e-6
