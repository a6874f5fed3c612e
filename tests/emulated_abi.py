"""tests/emulated_abi.py -- TEST INFRASTRUCTURE: a Python emulation of libu2pl_b200.so's C ABI for CPU tests.

Every method receives exactly what the ctypes call would pass (raw pointers as c_void_p / ints, scalars), reinterprets
HOST memory the way include/u2pl_b200.h documents the argument, and computes the documented result with torch / numpy
/ the oracle.  It lets the Python layers above the kernels (ops.py, fused.py, contra.py, step.py) run on the CPU in
tests; the kernels themselves are compared with the oracle on the GPU (tests/test_gpu_*.py, tools/cu/tc_selftest.cu).
Nothing outside tests/ may import this: the product has no CPU path and fails loudly without the CUDA library."""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from oracle import port

def _addr(p):
    return p.value if isinstance(p, ctypes.c_void_p) else (int(p) if p is not None else None)


def _view(p, shape, dtype):
    """Host memory at pointer p as a torch tensor (bf16 through its uint16 bit pattern)."""
    a = _addr(p)
    if a is None:
        return None
    n = int(np.prod(shape))
    if dtype is torch.bfloat16:
        arr = np.ctypeslib.as_array((ctypes.c_uint16 * n).from_address(a))
        return torch.from_numpy(arr).view(torch.bfloat16).view(*shape)
    arr = np.ctypeslib.as_array((ctypes.c_float * n).from_address(a))
    return torch.from_numpy(arr).view(*shape)


class FakeTC:
    """Semantics of include/u2pl_b200.h for the entry points the glue under test calls."""

    def u2pl_conv_bf16_nhwc(self, x, w, out, n, h, wd, cin, cout, k, d, scale, shift, res, relu, stream):
        xt = _view(x, (n, h, wd, cin), torch.bfloat16).float().permute(0, 3, 1, 2)
        wt = _view(w, (cout, k, k, cin), torch.bfloat16).float().permute(0, 3, 1, 2)
        y = F.conv2d(xt, wt, None, 1, d * (k // 2), d)
        if _addr(scale) is not None:
            y = y * _view(scale, (cout,), torch.float32)[None, :, None, None]
        if _addr(shift) is not None:
            y = y + _view(shift, (cout,), torch.float32)[None, :, None, None]
        if _addr(res) is not None:
            y = y + _view(res, (n, h, wd, cout), torch.bfloat16).float().permute(0, 3, 1, 2)
        if relu:
            y = F.relu(y)
        _view(out, (n, h, wd, cout), torch.bfloat16).copy_(y.permute(0, 2, 3, 1).bfloat16())
        return 0

    def u2pl_conv_bf16_nhwc_ex(self, x, w, out, n, h, wd, cin, cout, k, d, in_scale, in_shift, in_relu, scale, shift, res, relu,
                               part, sums, stream):
        xt = _view(x, (n, h, wd, cin), torch.bfloat16).float()
        if _addr(in_scale) is not None:
            xt = xt * _view(in_scale, (cin,), torch.float32)
        if _addr(in_shift) is not None:
            xt = xt + _view(in_shift, (cin,), torch.float32)
        z = (F.relu(xt) if in_relu else xt).bfloat16().contiguous()      # the kernel rewrites the tile in bf16
        rc = self.u2pl_conv_bf16_nhwc(ctypes.c_void_p(z.data_ptr()), w, out, n, h, wd, cin, cout, k, d, scale, shift, res, relu, stream)
        if _addr(sums) is not None:
            y = _view(out, (n * h * wd, cout), torch.bfloat16).float()
            _view(sums, (2, cout), torch.float32).copy_(torch.stack([y.sum(0), (y * y).sum(0)]))
        return rc

    def u2pl_conv_stat_parts(self, n, h, w, k):
        return (n * h * w + 127) // 128 if k == 1 else n * ((h + 7) // 8) * ((w + 15) // 16)

    def u2pl_conv_bf16_nhwc_stats(self, x, w, out, n, h, wd, cin, cout, k, d, part, sums, stream):
        self.u2pl_conv_bf16_nhwc(x, w, out, n, h, wd, cin, cout, k, d, None, None, None, 0, stream)
        y = _view(out, (n * h * wd, cout), torch.bfloat16).float()
        _view(sums, (2, cout), torch.float32).copy_(torch.stack([y.sum(0), (y * y).sum(0)]))
        return 0

    def u2pl_conv_wgrad_splits(self, n, h, w, cin, cout):
        return 3

    def u2pl_conv_wgrad_bf16_nhwc(self, x, g, part, n, h, wd, cin, cout, d, stream):
        xt = _view(x, (n, h, wd, cin), torch.bfloat16).float().permute(0, 3, 1, 2)
        gt = _view(g, (n, h, wd, cout), torch.bfloat16).float().permute(0, 3, 1, 2)
        with torch.enable_grad():                                          # (this fake may run inside an autograd backward)
            wz = torch.zeros(cout, cin, 3, 3, requires_grad=True)
            F.conv2d(xt, wz, None, 1, d, d).backward(gt)
        dw = wz.grad.permute(2, 3, 0, 1).reshape(9, cout, cin)           # [tap][co][ci]
        p = _view(part, (3, 9, cout, cin), torch.float32)
        p[0].copy_(dw * 0.5); p[1].copy_(dw * 0.25); p[2].copy_(dw * 0.25)   # the caller must sum the splits
        return 0

    def u2pl_bn_fold(self, C, gamma, beta, mean, var, eps, scale, shift, stream):
        g, b = _view(gamma, (C,), torch.float32), _view(beta, (C,), torch.float32)
        m, v = _view(mean, (C,), torch.float32), _view(var, (C,), torch.float32)
        s = g / torch.sqrt(v + eps)
        _view(scale, (C,), torch.float32).copy_(s)
        _view(shift, (C,), torch.float32).copy_(b - m * s)
        return 0

    # ---- csrc/bn.cu entry points (x / y / dy / residual: [M, C] bf16 rows = channels-last pixels)
    def u2pl_bn_parts(self):
        return 4

    def u2pl_bn_stats(self, x, M, C, partial, sums, stream):
        xv = _view(x, (M, C), torch.bfloat16).float()
        _view(sums, (2, C), torch.float32).copy_(torch.stack([xv.sum(0), (xv * xv).sum(0)]))
        return 0

    def u2pl_bn_finalize(self, sums, C, count, gamma, beta, rmean, rvar, momentum, eps, mean, invstd, scale, shift, stream):
        n = count.value if hasattr(count, "value") else float(count)
        sm = _view(sums, (2, C), torch.float32).double()
        mu = sm[0] / n
        var = (sm[1] / n - mu * mu).clamp_min(0)
        inv = 1.0 / torch.sqrt(var + eps)
        g, b = _view(gamma, (C,), torch.float32).double(), _view(beta, (C,), torch.float32).double()
        _view(mean, (C,), torch.float32).copy_(mu.float())
        _view(invstd, (C,), torch.float32).copy_(inv.float())
        _view(scale, (C,), torch.float32).copy_((g * inv).float())
        _view(shift, (C,), torch.float32).copy_((b - mu * g * inv).float())
        rm, rv = _view(rmean, (C,), torch.float32), _view(rvar, (C,), torch.float32)
        rm.mul_(1 - momentum).add_(momentum * mu.float())
        rv.mul_(1 - momentum).add_(momentum * (var * n / (n - 1)).float())
        return 0

    def u2pl_bn_apply(self, x, res, scale, shift, M, C, relu, y, stream):
        v = _view(x, (M, C), torch.bfloat16).float() * _view(scale, (C,), torch.float32) + _view(shift, (C,), torch.float32)
        if _addr(res) is not None:
            v = v + _view(res, (M, C), torch.bfloat16).float()
        _view(y, (M, C), torch.bfloat16).copy_((F.relu(v) if relu else v).bfloat16())
        return 0

    @staticmethod
    def _masked(dy, y, M, C):
        g = _view(dy, (M, C), torch.bfloat16).float()
        return g * (_view(y, (M, C), torch.bfloat16).float() > 0) if _addr(y) is not None else g

    def u2pl_bn_backward_reduce(self, dy, x, y, mean, invstd, M, C, partial, sums, stream):
        g = self._masked(dy, y, M, C)
        xh = (_view(x, (M, C), torch.bfloat16).float() - _view(mean, (C,), torch.float32)) * _view(invstd, (C,), torch.float32)
        _view(sums, (2, C), torch.float32).copy_(torch.stack([g.sum(0), (g * xh).sum(0)]))
        return 0

    def u2pl_bn_backward_elemt(self, dy, x, y, mean, invstd, gamma, sums, count, M, C, coef, dx, dres, stream):
        n = count.value if hasattr(count, "value") else float(count)
        g = self._masked(dy, y, M, C)
        sm = _view(sums, (2, C), torch.float32)
        inv, mu, ga = _view(invstd, (C,), torch.float32), _view(mean, (C,), torch.float32), _view(gamma, (C,), torch.float32)
        A = ga * inv
        B = -ga * inv * inv * sm[1] / n
        D = -A * sm[0] / n - B * mu
        _view(dx, (M, C), torch.bfloat16).copy_((A * g + B * _view(x, (M, C), torch.bfloat16).float() + D).bfloat16())
        if _addr(dres) is not None:
            _view(dres, (M, C), torch.bfloat16).copy_(g.bfloat16())
        return 0

    def u2pl_last_error(self):
        return b""



BLK = 256


def _arr(p, n, ctype):
    return np.ctypeslib.as_array((ctype * int(n)).from_address(_addr(p)))


def _feat(p, sn, sd, sp, N, D, hw):
    """[N, hw, D] strided view of a feature tensor addressed as n*sn + d*sd + pixel*sp (elements)."""
    size = (N - 1) * sn + (D - 1) * sd + (hw - 1) * sp + 1
    flat = _arr(p, size, ctypes.c_float)
    return np.lib.stride_tricks.as_strided(flat, (N, hw, D), (4 * sn, 4 * sp, 4 * sd))


class FakeContra:
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




class FakeLoss:
    """entropy / percentile / partition / cross-entropy entry points (A6-A8, A12 CE) through the oracle."""

    def u2pl_entropy_ws_bytes(self, B, HW):
        return 64

    u2pl_entropy_fast_ws_bytes = u2pl_ce_ws_bytes = u2pl_entropy_ws_bytes

    def u2pl_entropy_thresholds(self, logits, target, B, C, HW, ignore, hq, nq, ent, thresh, n_valid, ws, ws_bytes, stream):
        x = _arr(logits, B * C * HW, ctypes.c_float).reshape(B, C, HW)
        t = _arr(target, B * HW, ctypes.c_int64).reshape(B, HW)
        e = port.entropy(x)
        _arr(ent, B * HW, ctypes.c_float).reshape(B, HW)[:] = e
        valid = t != ignore
        th = _arr(thresh, nq, ctypes.c_float)
        for j in range(nq):
            th[j] = port.percentile(e[valid], float(hq[j]))
        _arr(n_valid, 1, ctypes.c_int64)[0] = int(valid.sum())
        return 0

    u2pl_entropy_thresholds_fast = u2pl_entropy_thresholds

    def u2pl_entropy_partition_fused(self, logits, target_in, B, C, HW, ignore, hq, nq, part_idx, ent, thresh, n_valid,
                                     target_out, mask, n_kept, ws, ws_bytes, stream):
        self.u2pl_entropy_thresholds(logits, target_in, B, C, HW, ignore, hq, nq, ent, thresh, n_valid, ws, ws_bytes, stream)
        _arr(target_out, B * HW, ctypes.c_int64)[:] = _arr(target_in, B * HW, ctypes.c_int64)
        return self.u2pl_partition_target(ent, target_out, B * HW, ignore, thresh, part_idx, mask, n_kept, stream)

    def u2pl_partition_target(self, entropy, target, n, ignore, thresh, idx, mask, n_kept, stream):
        e = _arr(entropy, n, ctypes.c_float)
        t = _arr(target, n, ctypes.c_int64)
        th = np.float32(_arr(thresh, idx + 1, ctypes.c_float)[idx])
        drop = (e >= th) & (t != ignore)
        t[drop] = ignore
        if _addr(mask) is not None:
            _arr(mask, n, ctypes.c_uint8)[:] = drop
        _arr(n_kept, 1, ctypes.c_int64)[0] = int((t != ignore).sum())
        return 0

    def u2pl_upsample_fused_supported(self, C):
        return 1 if C in (19, 21) else 0

    def u2pl_upce_ws_bytes(self):
        return 148 * 8 * 8

    @staticmethod
    def _up(low, B, C, h, w, H, W):
        x = torch.from_numpy(_arr(low, B * C * h * w, ctypes.c_float).reshape(B, C, h, w).copy())
        return F.interpolate(x, (H, W), mode="bilinear", align_corners=True)

    def u2pl_up_softmax_max(self, low, B, C, h, w, H, W, out_prob, out_label, stream):
        prob, lab = torch.max(F.softmax(self._up(low, B, C, h, w, H, W), dim=1), dim=1)
        _arr(out_prob, B * H * W, ctypes.c_float).reshape(B, H, W)[:] = prob.numpy()
        _arr(out_label, B * H * W, ctypes.c_int64).reshape(B, H, W)[:] = lab.numpy()
        return 0

    def u2pl_upce_forward(self, low, target, B, C, h, w, H, W, ignore, nll, n_used, ws, ws_bytes, stream):
        x = self._up(low, B, C, h, w, H, W).numpy().reshape(B, C, H * W)
        t = _arr(target, B * H * W, ctypes.c_int64).reshape(B, H * W)
        s, n = port.cross_entropy_sum(x, t, ignore)
        _arr(nll, 1, ctypes.c_float)[0] = s
        _arr(n_used, 1, ctypes.c_int64)[0] = n
        return 0

    def u2pl_upce_backward(self, low, target, B, C, h, w, H, W, ignore, scale, grad_low, stream):
        x = torch.from_numpy(_arr(low, B * C * h * w, ctypes.c_float).reshape(B, C, h, w).copy()).double()
        x.requires_grad_(True)
        t = torch.from_numpy(_arr(target, B * H * W, ctypes.c_int64).reshape(B, H, W).copy())
        with torch.enable_grad():                              # called from inside an autograd backward
            up = F.interpolate(x, (H, W), mode="bilinear", align_corners=True)
            loss = F.cross_entropy(up, t, ignore_index=ignore, reduction="sum") * float(_arr(scale, 1, ctypes.c_float)[0])
            loss.backward()
        _arr(grad_low, B * C * h * w, ctypes.c_float).reshape(B, C, h, w)[:] = x.grad.float().numpy()
        return 0

    def u2pl_sgd_tensor_bytes(self):
        return 56

    def u2pl_sgd_chunk_elems(self):
        return 8192

    def u2pl_sgd_ema_step(self, tensor_table, chunk_table, n_chunks, momentum, decay, one_minus, do_ema, stream):
        import struct
        ch = _arr(chunk_table, 2 * n_chunks, ctypes.c_uint32).reshape(-1, 2)
        n_tensors = int(ch[:, 0].max()) + 1
        raw = ctypes.string_at(_addr(tensor_table), 56 * n_tensors)
        for k in range(n_tensors):
            pp, gp, mp, tp, n, lr, wd, first, _ = struct.unpack_from("<QQQQqffii", raw, 56 * k)
            p = _arr(ctypes.c_void_p(pp), n, ctypes.c_float)
            g = _arr(ctypes.c_void_p(gp), n, ctypes.c_float)
            m = _arr(ctypes.c_void_p(mp), n, ctypes.c_float)
            d = g + np.float32(wd) * p
            m[:] = d if first else np.float32(momentum) * m + d
            p[:] = p - np.float32(lr) * m
            if do_ema and tp:
                t = _arr(ctypes.c_void_p(tp), n, ctypes.c_float)
                t[:] = np.float32(decay) * t + np.float32(one_minus) * p
        return 0

    def u2pl_ce_forward(self, logits, target, B, C, HW, ignore, nll, n_used, ws, ws_bytes, stream):
        x = _arr(logits, B * C * HW, ctypes.c_float).reshape(B, C, HW)
        t = _arr(target, B * HW, ctypes.c_int64).reshape(B, HW)
        s, n = port.cross_entropy_sum(x, t, ignore)
        _arr(nll, 1, ctypes.c_float)[0] = s
        _arr(n_used, 1, ctypes.c_int64)[0] = n
        return 0

    def u2pl_ce_backward(self, logits, target, B, C, HW, ignore, scale, grad, stream):
        x = torch.from_numpy(_arr(logits, B * C * HW, ctypes.c_float).reshape(B, C, HW).copy()).double()
        t = _arr(target, B * HW, ctypes.c_int64).reshape(B, HW)
        valid = torch.from_numpy(t != ignore)
        oh = F.one_hot(torch.from_numpy(np.where(t == ignore, 0, t)), C).permute(0, 2, 1).double()
        g = (torch.softmax(x, 1) - oh) * valid[:, None] * float(_arr(scale, 1, ctypes.c_float)[0])
        _arr(grad, B * C * HW, ctypes.c_float).reshape(B, C, HW)[:] = g.float().numpy()
        return 0

    def u2pl_unsup_finalize(self, nll, n_kept, total, upstream, loss, bwd_scale, stream):
        s, n = float(_arr(nll, 1, ctypes.c_float)[0]), int(_arr(n_kept, 1, ctypes.c_int64)[0])
        up = float(_arr(upstream, 1, ctypes.c_float)[0]) if _addr(upstream) is not None else 1.0
        with np.errstate(all="ignore"):
            if _addr(loss) is not None:
                _arr(loss, 1, ctypes.c_float)[0] = np.float32(total) / np.float32(n) * (np.float32(s) / np.float32(n)) if n else np.nan
            if _addr(bwd_scale) is not None:
                _arr(bwd_scale, 1, ctypes.c_float)[0] = up * total / n / n if n else np.nan
        return 0

    def u2pl_ohem_select(self, logits, target, B, C, HW, ignore, thresh, min_kept, new_target, kth_value, n_valid, ws, ws_bytes, stream):
        x = _arr(logits, B * C * HW, ctypes.c_float).reshape(B, C, HW)
        t = _arr(target, B * HW, ctypes.c_int64).reshape(-1)
        valid = t != ignore
        nv = int(valid.sum())
        keep = valid.copy()
        kth = np.float32(np.nan)
        if nv > 0 and min_kept <= nv:                                       # loss_helper.py:512-526
            prob = port.softmax(x).transpose(1, 0, 2).reshape(C, -1)
            mp = np.where(valid, prob[np.where(valid, t, 0), np.arange(t.size)], np.float32(1))
            th = np.float32(thresh)
            if min_kept > 0:
                kth = np.sort(mp, kind="stable")[min(mp.size, min_kept) - 1]
                th = max(th, kth)
                keep = valid & (mp <= th)
        _arr(new_target, B * HW, ctypes.c_int64)[:] = np.where(keep, t, ignore)
        _arr(kth_value, 1, ctypes.c_float)[0] = kth
        _arr(n_valid, 1, ctypes.c_int64)[0] = nv
        return 0

    def u2pl_contra_prep_lowres(self, label_l, label_u, entropy, thresh, lo_idx, hi_idx, Bl, Bu, H, W, h, w, C, ignore, neg_high,
                                bits, low, high, stream):
        ll = _arr(label_l, Bl * H * W, ctypes.c_int64).reshape(Bl, H, W)
        lu = _arr(label_u, Bu * H * W, ctypes.c_int64).reshape(Bu, H, W)
        e = _arr(entropy, Bu * H * W, ctypes.c_float).reshape(Bu, H, W)
        th = _arr(thresh, max(lo_idx, hi_idx) + 1, ctypes.c_float)
        valid_u = lu != ignore
        lo = (e <= np.float32(th[lo_idx])).astype(np.float32) * valid_u                     # train_semi.py:408-410
        hi = (e >= np.float32(th[hi_idx])).astype(np.float32) * valid_u                     # :416-418
        lab = (ll != ignore).astype(np.float32)
        sy, sx = port.nearest_src_index(h, H), port.nearest_src_index(w, W)
        down = lambda a: a[..., sy[:, None], sx[None, :]]                                   # noqa: E731
        _arr(low, (Bl + Bu) * h * w, ctypes.c_float).reshape(Bl + Bu, h, w)[:] = down(np.concatenate([lab, lo]))
        second = hi if neg_high else np.ones_like(hi)
        _arr(high, (Bl + Bu) * h * w, ctypes.c_float).reshape(Bl + Bu, h, w)[:] = down(np.concatenate([lab, second]))
        oh = np.concatenate([down(port.label_onehot(ll, C, ignore)), down(port.label_onehot(lu, C, ignore))])   # [Bl+Bu, C, h, w]
        out = _arr(bits, (Bl + Bu) * h * w, ctypes.c_uint32).reshape(Bl + Bu, h, w)
        out[:] = 0
        for c in range(C):
            out |= ((oh[:, c] != 0).astype(np.uint32) << np.uint32(c))
        return 0


class FakeShards:
    """u2pl_shard_alloc / open / close with POSIX shared memory standing in for CUDA IPC: the 64-byte handle carries the
    segment name, a peer process maps the same bytes.  Plus the sharded InfoNCE entry point, which reads each active
    class's negatives through its own base pointer."""

    def __init__(self):
        self._segments = {}

    def u2pl_shard_alloc(self, nbytes, dptr, handle64):
        from multiprocessing import shared_memory
        seg = shared_memory.SharedMemory(create=True, size=int(nbytes))
        np.frombuffer(seg.buf, dtype=np.uint8)[:] = 0
        addr = ctypes.addressof(ctypes.c_char.from_buffer(seg.buf))
        self._segments[addr] = seg
        dptr._obj.value = addr                                     # ctypes.byref(c_void_p)
        name = seg.name.encode()
        ctypes.memmove(handle64, name + b"\0" * (64 - len(name)), 64)
        return 0

    def u2pl_shard_open(self, handle64, dptr):
        from multiprocessing import shared_memory
        name = bytes(handle64).split(b"\0")[0].decode()
        seg = shared_memory.SharedMemory(name=name)
        addr = ctypes.addressof(ctypes.c_char.from_buffer(seg.buf))
        self._segments[addr] = seg
        dptr._obj.value = addr
        return 0

    def u2pl_shard_close(self, ptr, owned):
        seg = self._segments.pop(_addr(ptr) if not isinstance(ptr, int) else ptr, None)
        if seg is not None:
            try:
                seg.close()
                if owned:
                    seg.unlink()
            except (BufferError, FileNotFoundError):
                pass
        return 0

    def u2pl_infonce_forward_sharded(self, rep, sn, sd, sp, P, D, hw, an_bits, blockoff_an, act_class, a_ord, neg_rows, proto,
                                     class_bank, nact, nq, nneg, temp, valid_seg, loss_q, grad_rows, anchor_pix, loss, stream):
        bases = _arr(class_bank, nact, ctypes.c_int64)
        nr = _arr(neg_rows, nact * nq * nneg, ctypes.c_int32).reshape(nact, nq, nneg).astype(np.int64)
        # gather every active class's rows from its own shard into one temporary bank, then reuse the replicated-bank maths
        tmp = np.zeros((nact * nq * nneg, D), np.float32)
        remap = np.arange(nact * nq * nneg, dtype=np.int32).reshape(nact, nq, nneg)
        for a in range(nact):
            shard = _arr(int(bases[a]), (int(nr[a].max()) + 1) * D, ctypes.c_float).reshape(-1, D)
            tmp[remap[a].ravel()] = shard[nr[a].ravel()]
        return self.u2pl_infonce_forward(rep, sn, sd, sp, P, D, hw, an_bits, blockoff_an, act_class, a_ord,
                                         ctypes.c_void_p(remap.ctypes.data), proto, ctypes.c_void_p(tmp.ctypes.data), nact, nq, nneg,
                                         temp, valid_seg, loss_q, grad_rows, anchor_pix, loss, stream)


class FakePool:
    """csrc/pool.cu: 3x3 / stride 2 / pad 1 / ceil_mode max-pooling of [n,h,w,c] bf16 with a uint8 tap map."""

    def u2pl_maxpool3s2_out(self, n):
        return F.max_pool2d(torch.zeros(1, 1, n, n), 3, 2, 1, ceil_mode=True).shape[-1]

    def u2pl_maxpool3s2_forward(self, x, y, tap, n, h, w, c, stream):
        xt = _view(x, (n, h, w, c), torch.bfloat16).float().permute(0, 3, 1, 2)
        out, idx = F.max_pool2d(xt, 3, 2, 1, ceil_mode=True, return_indices=True)          # idx: flat h*W + w in the input plane
        ho, wo = out.shape[2:]
        _view(y, (n, ho, wo, c), torch.bfloat16).copy_(out.permute(0, 2, 3, 1).bfloat16())
        if _addr(tap) is not None:
            ih, iw = idx // w, idx % w
            kh = ih - (2 * torch.arange(ho).view(1, 1, ho, 1) - 1)
            kw = iw - (2 * torch.arange(wo).view(1, 1, 1, wo) - 1)
            t = (kh * 3 + kw).permute(0, 2, 3, 1).to(torch.uint8).contiguous()
            _arr(tap, n * ho * wo * c, ctypes.c_uint8)[:] = t.numpy().ravel()
        return 0

    def u2pl_maxpool3s2_backward(self, dy, tap, dx, n, h, w, c, stream):
        ho, wo = self.u2pl_maxpool3s2_out(h), self.u2pl_maxpool3s2_out(w)
        g = _view(dy, (n, ho, wo, c), torch.bfloat16).float()
        t = torch.from_numpy(_arr(tap, n * ho * wo * c, ctypes.c_uint8).reshape(n, ho, wo, c).astype(np.int64))
        out = torch.zeros(n, h, w, c)
        hh = (2 * torch.arange(ho).view(1, ho, 1, 1) - 1) + t // 3
        ww = (2 * torch.arange(wo).view(1, 1, wo, 1) - 1) + t % 3
        nn_ = torch.arange(n).view(n, 1, 1, 1).expand_as(t)
        cc = torch.arange(c).view(1, 1, 1, c).expand_as(t)
        out.index_put_((nn_, hh, ww, cc), g, accumulate=True)
        _view(dx, (n, h, w, c), torch.bfloat16).copy_(out.bfloat16())
        return 0


class FakeAll(FakeTC, FakeContra, FakeLoss, FakeShards, FakePool):
    def __init__(self):
        FakeShards.__init__(self)

    def u2pl_last_error(self):
        return b""

    def u2pl_launch_count(self):
        return 0


class DirectPatcher:
    """monkeypatch stand-in for spawned worker processes (nothing to undo there)."""

    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)


def _host_tensor_from_ptr(ptr, shape, device=None):
    n = int(np.prod(shape))
    return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_float * n).from_address(int(ptr)))).view(*shape)


def install(monkeypatch, fake=None):
    """Route u2pl_b200's ctypes layer to the emulation and lift its CUDA-only guards (monkeypatch scope)."""
    from u2pl_b200 import _lib, bank, contra, fused, ops
    monkeypatch.setattr(bank, "_tensor_from_ptr", _host_tensor_from_ptr)
    fake = fake or FakeAll()
    monkeypatch.setattr(_lib, "load", lambda *a, **k: fake)
    for mod in (ops, contra):
        monkeypatch.setattr(mod, "_need_cuda", lambda *ts: None)
    for mod in (ops, contra, fused):
        monkeypatch.setattr(mod, "_stream", lambda: None)
    monkeypatch.setattr(fused, "_is_cl_bf16", lambda x: x.dtype == torch.bfloat16 and x.dim() == 4
                        and x.is_contiguous(memory_format=torch.channels_last))
    monkeypatch.setattr(contra, "_to_device_i32",
                        lambda name, arr, device: torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int32).ravel().copy()))
    return fake
