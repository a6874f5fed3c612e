"""Torch-facing wrappers over the C ABI (include/u2pl_b200.h).

PyTorch is plumbing here: it owns device memory, the current stream and autograd
bookkeeping; every computation below is a kernel of libu2pl_b200.so.  All functions
require CUDA tensors and raise otherwise -- there is no CPU path.
"""
import ctypes

import torch

from . import _lib

_WS = {}


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.U2PLNativeError("u2pl_b200 kernels need CUDA tensors (no CPU fallback exists)")


def _workspace(name, nbytes, device):
    key = (name, device)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# --------------------------------------------------------------------------- entropy / percentiles
def entropy_thresholds(logits, target, percents, ignore=255, exact_map=False):
    """Softmax entropy of `logits` [B,C,H,W] + np.percentile-compatible thresholds.

    Mirrors loss_helper.py:35-40 / train_semi.py:402-415.  Returns (entropy [B,H,W] fp32,
    thresh [len(percents)] fp32 device tensor, n_valid int64 device scalar).
    exact_map=False (default): two-level path -- thresholds and every comparison against them are exact,
    the entropy map itself is exact only near the thresholds (elsewhere within 1e-4 of the contract).
    exact_map=True: contract arithmetic for every pixel."""
    _need_cuda(logits, target)
    lib = _lib.load()
    logits = _f32c(logits)
    target = target.contiguous()
    assert target.dtype == torch.int64
    B, C, H, W = logits.shape
    HW = H * W
    nq = len(percents)
    ent = torch.empty((B, H, W), dtype=torch.float32, device=logits.device)
    thresh = torch.empty(nq, dtype=torch.float32, device=logits.device)
    n_valid = torch.empty((), dtype=torch.int64, device=logits.device)
    hq = (ctypes.c_float * nq)(*[float(q) for q in percents])
    if exact_map:
        ws = _workspace("entropy", lib.u2pl_entropy_ws_bytes(B, HW), logits.device)
        fn, name = lib.u2pl_entropy_thresholds, "u2pl_entropy_thresholds"
    else:
        ws = _workspace("entropy_fast", lib.u2pl_entropy_fast_ws_bytes(B, HW), logits.device)
        fn, name = lib.u2pl_entropy_thresholds_fast, "u2pl_entropy_thresholds_fast"
    rc = fn(_p(logits), _p(target), B, C, HW, int(ignore), hq, nq,
            _p(ent), _p(thresh), _p(n_valid), _p(ws), ws.numel(), _stream())
    _lib.check(rc, name)
    return ent, thresh, n_valid


def entropy_partition(logits, target, percents, part_idx=0, ignore=255, want_mask=False):
    """loss_helper.py:35-44 (+ the extra percentiles of train_semi.py:402-415) in ONE kernel launch
    (u2pl_entropy_partition_fused): returns (entropy [B,H,W], thresh [len(percents)], n_valid, new_target, n_kept, drop_mask).
    `target` is left untouched; `new_target` is the clone-with-unreliable-pixels-ignored the reference builds in place."""
    _need_cuda(logits, target)
    lib = _lib.load()
    logits = _f32c(logits)
    target = target.contiguous()
    assert target.dtype == torch.int64
    B, C, H, W = logits.shape
    HW = H * W
    nq = len(percents)
    dev = logits.device
    ent = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    thresh = torch.empty(nq, dtype=torch.float32, device=dev)
    n_valid = torch.empty((), dtype=torch.int64, device=dev)
    n_kept = torch.empty((), dtype=torch.int64, device=dev)
    new_target = torch.empty_like(target)
    mask = torch.empty(target.shape, dtype=torch.uint8, device=dev) if want_mask else None
    hq = (ctypes.c_float * nq)(*[float(q) for q in percents])
    ws = _workspace("entropy_fast", lib.u2pl_entropy_fast_ws_bytes(B, HW), dev)
    rc = lib.u2pl_entropy_partition_fused(_p(logits), _p(target), B, C, HW, int(ignore), hq, nq, int(part_idx),
                                          _p(ent), _p(thresh), _p(n_valid), _p(new_target), _p(mask), _p(n_kept),
                                          _p(ws), ws.numel(), _stream())
    _lib.check(rc, "u2pl_entropy_partition_fused")
    return ent, thresh, n_valid, new_target, n_kept, mask


def partition_target_(entropy, target, thresh, thresh_idx=0, ignore=255, want_mask=False):
    """In place: target[(entropy >= thresh[idx]) & (target != ignore)] = ignore  (loss_helper.py:41-43).
    Returns (n_kept int64 device scalar, drop mask uint8 or None)."""
    _need_cuda(entropy, target, thresh)
    lib = _lib.load()
    assert target.is_contiguous() and target.dtype == torch.int64 and entropy.is_contiguous()
    n = target.numel()
    mask = torch.empty(target.shape, dtype=torch.uint8, device=target.device) if want_mask else None
    n_kept = torch.empty((), dtype=torch.int64, device=target.device)
    rc = lib.u2pl_partition_target(_p(entropy), _p(target), n, int(ignore), _p(thresh), int(thresh_idx),
                                   _p(mask), _p(n_kept), _stream())
    _lib.check(rc, "u2pl_partition_target")
    return n_kept, mask


def entropy_masks(entropy, target, thresh, lo_idx, hi_idx, idx=None, ignore=255, out_shape=None):
    """low/high entropy masks of train_semi.py:408-418, optionally gathered at flat positions `idx`."""
    _need_cuda(entropy, target, thresh, idx)
    lib = _lib.load()
    n_out = idx.numel() if idx is not None else entropy.numel()
    shape = out_shape if out_shape is not None else (tuple(idx.shape) if idx is not None else tuple(entropy.shape))
    low = torch.empty(shape, dtype=torch.float32, device=entropy.device)
    high = torch.empty(shape, dtype=torch.float32, device=entropy.device)
    rc = lib.u2pl_entropy_masks(_p(entropy), _p(target), _p(idx), n_out, int(ignore), _p(thresh),
                                int(lo_idx), int(hi_idx), _p(low), _p(high), _stream())
    _lib.check(rc, "u2pl_entropy_masks")
    return low, high


# --------------------------------------------------------------------------- cross entropy
def _ce_forward(logits, target, ignore):
    lib = _lib.load()
    B, C, H, W = logits.shape
    nll = torch.empty((), dtype=torch.float32, device=logits.device)
    n_used = torch.empty((), dtype=torch.int64, device=logits.device)
    ws = _workspace("ce", lib.u2pl_ce_ws_bytes(B, H * W), logits.device)
    rc = lib.u2pl_ce_forward(_p(logits), _p(target), B, C, H * W, int(ignore), _p(nll), _p(n_used),
                             _p(ws), ws.numel(), _stream())
    _lib.check(rc, "u2pl_ce_forward")
    return nll, n_used


def _ce_backward(logits, target, ignore, scale):
    lib = _lib.load()
    B, C, H, W = logits.shape
    grad = torch.empty_like(logits)
    rc = lib.u2pl_ce_backward(_p(logits), _p(target), B, C, H * W, int(ignore), _p(scale), _p(grad), _stream())
    _lib.check(rc, "u2pl_ce_backward")
    return grad


class _CrossEntropyMean(torch.autograd.Function):
    """nn.CrossEntropyLoss(ignore_index) with mean reduction (loss_helper.py:265,319)."""

    @staticmethod
    def forward(ctx, logits, target, ignore):
        _need_cuda(logits, target)
        logits = _f32c(logits)
        target = target.contiguous()
        nll, n_used = _ce_forward(logits, target, ignore)
        ctx.save_for_backward(logits, target, n_used)
        ctx.ignore = ignore
        return nll / n_used.to(torch.float32)

    @staticmethod
    def backward(ctx, gout):
        logits, target, n_used = ctx.saved_tensors
        scale = (gout.to(torch.float32) / n_used.to(torch.float32)).reshape(1).contiguous()
        return _ce_backward(logits, target, ctx.ignore, scale), None, None


def cross_entropy_mean(logits, target, ignore=255):
    return _CrossEntropyMean.apply(logits, target, ignore)


class _UnsupCE(torch.autograd.Function):
    """Masked-CE half of compute_unsupervised_loss (loss_helper.py:44-46) on an already partitioned
    target: CE forward and the weight = B*H*W / #kept algebra, all without a host round trip."""

    @staticmethod
    def forward(ctx, predict, target, n_kept, ignore):
        _need_cuda(predict, target, n_kept)
        lib = _lib.load()
        predict = _f32c(predict)
        assert target.is_contiguous() and target.dtype == torch.int64
        nll, _ = _ce_forward(predict, target, ignore)
        loss = torch.empty((), dtype=torch.float32, device=predict.device)
        total = target.numel()
        rc = lib.u2pl_unsup_finalize(_p(nll), _p(n_kept), total, None, _p(loss), None, _stream())
        _lib.check(rc, "u2pl_unsup_finalize")
        ctx.save_for_backward(predict, target, n_kept, nll)
        ctx.ignore = ignore
        ctx.total = total
        return loss

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        predict, target, n_kept, nll = ctx.saved_tensors
        gout = gout.to(torch.float32).contiguous()
        scale = torch.empty(1, dtype=torch.float32, device=predict.device)
        rc = lib.u2pl_unsup_finalize(_p(nll), _p(n_kept), ctx.total, _p(gout), None, _p(scale), _stream())
        _lib.check(rc, "u2pl_unsup_finalize(bwd)")
        return _ce_backward(predict, target, ctx.ignore, scale), None, None, None


def unsup_ce(predict, target_partitioned, n_kept, ignore=255):
    return _UnsupCE.apply(predict, target_partitioned, n_kept, ignore)


def unsup_loss_from_entropy(predict, target, ent, thresh, thresh_idx=0, ignore=255):
    """Partition (rewrites `target` in place like loss_helper.py:43) + weighted masked CE."""
    n_kept, _ = partition_target_(ent, target, thresh, int(thresh_idx), ignore)
    return _UnsupCE.apply(predict, target, n_kept, ignore)


def unsup_loss(predict, target, percent, pred_teacher, ignore=255):
    """compute_unsupervised_loss(predict, target, percent, pred_teacher) -- loss_helper.py:30-48."""
    _, _, _, new_target, n_kept, _ = entropy_partition(pred_teacher.detach(), target, [float(percent)], 0, ignore)
    target.copy_(new_target)                                  # the reference mutates the caller's tensor (loss_helper.py:43)
    return _UnsupCE.apply(predict, target, n_kept, ignore)


# --------------------------------------------------------------------------- fused x4 bilinear up-sampling consumers
def upsample_fused_supported(num_classes):
    return bool(_lib.load().u2pl_upsample_fused_supported(int(num_classes)))


def _low(pred_low):
    """Low-resolution logits as the kernels read them: fp32 NCHW contiguous (22 MB for V16 -- the small tensor is
    converted, the 354 MB full-resolution one is never built)."""
    return pred_low.detach().float().contiguous()


def up_softmax_max(pred_low, size):
    """torch.max(F.softmax(F.interpolate(pred_low, size, mode="bilinear", align_corners=True), dim=1), dim=1)
    (train_semi.py:318-324) in one kernel: returns (max probability fp32 [B,H,W], arg-max int64 [B,H,W])."""
    _need_cuda(pred_low)
    lib = _lib.load()
    low = _low(pred_low)
    B, C, h, w = low.shape
    H, W = size
    prob = torch.empty((B, H, W), dtype=torch.float32, device=low.device)
    label = torch.empty((B, H, W), dtype=torch.int64, device=low.device)
    rc = lib.u2pl_up_softmax_max(_p(low), B, C, h, w, H, W, _p(prob), _p(label), _stream())
    _lib.check(rc, "u2pl_up_softmax_max")
    return prob, label


class _UpCE(torch.autograd.Function):
    """F.cross_entropy(F.interpolate(pred_low, (H,W), bilinear, align_corners=True), target, ignore_index) with the
    gradient delivered directly at low resolution (csrc/upsample_ce.cu).  mode "mean": nll / #valid
    (loss_helper.py:313-319).  mode "unsup": (B*H*W / n_kept) * nll / n_kept on an already partitioned target
    (loss_helper.py:44-46)."""

    @staticmethod
    def forward(ctx, pred_low, target, n_kept, ignore):
        _need_cuda(pred_low, target, n_kept)
        lib = _lib.load()
        low = _low(pred_low)
        B, C, h, w = low.shape
        H, W = target.shape[1:]
        assert target.is_contiguous() and target.dtype == torch.int64 and target.shape[0] == B
        dev = low.device
        nll = torch.empty((), dtype=torch.float32, device=dev)
        n_used = torch.empty((), dtype=torch.int64, device=dev)
        ws = _workspace("upce", lib.u2pl_upce_ws_bytes(), dev)
        rc = lib.u2pl_upce_forward(_p(low), _p(target), B, C, h, w, H, W, int(ignore), _p(nll), _p(n_used),
                                   _p(ws), ws.numel(), _stream())
        _lib.check(rc, "u2pl_upce_forward")
        ctx.ignore, ctx.geom, ctx.in_dtype, ctx.unsup = ignore, (B, C, h, w, H, W), pred_low.dtype, n_kept is not None
        if n_kept is None:
            ctx.save_for_backward(low, target, n_used)
            return nll / n_used.to(torch.float32)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        rc = lib.u2pl_unsup_finalize(_p(nll), _p(n_kept), target.numel(), None, _p(loss), None, _stream())
        _lib.check(rc, "u2pl_unsup_finalize")
        ctx.save_for_backward(low, target, n_kept, nll)
        return loss

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        B, C, h, w, H, W = ctx.geom
        gout = gout.to(torch.float32).contiguous()
        if ctx.unsup:
            low, target, n_kept, nll = ctx.saved_tensors
            scale = torch.empty(1, dtype=torch.float32, device=low.device)
            rc = lib.u2pl_unsup_finalize(_p(nll), _p(n_kept), target.numel(), _p(gout), None, _p(scale), _stream())
            _lib.check(rc, "u2pl_unsup_finalize(bwd)")
        else:
            low, target, n_used = ctx.saved_tensors
            scale = (gout / n_used.to(torch.float32)).reshape(1).contiguous()
        grad = torch.empty_like(low)
        rc = lib.u2pl_upce_backward(_p(low), _p(target), B, C, h, w, H, W, int(ctx.ignore), _p(scale), _p(grad), _stream())
        _lib.check(rc, "u2pl_upce_backward")
        return grad.to(ctx.in_dtype), None, None, None


def upsampled_ce_mean(pred_low, target, ignore=255):
    return _UpCE.apply(pred_low, target.contiguous(), None, ignore)


def upsampled_unsup_ce(pred_low, target_partitioned, n_kept, ignore=255):
    return _UpCE.apply(pred_low, target_partitioned, n_kept, ignore)


# --------------------------------------------------------------------------- A8 helpers
def label_onehot(inputs, num_segments, ignore=255):
    """utils.py:50-59 with its scatter quirk (Q8): slot 0 = union of all images' classes (ignored
    pixels count as class 0), cleared where image 0 is ignored; slots b > 0 all-zero.  Only the
    unchanged reference driver calls this (train_semi.py:456-465); the fused step never builds the
    [B,C,H,W] tensor (contra_prep_lowres)."""
    _need_cuda(inputs)
    B, H, W = inputs.shape
    out = torch.zeros((B, num_segments, H, W), dtype=torch.float32, device=inputs.device)
    t = torch.where(inputs == ignore, torch.zeros_like(inputs), inputs)
    out[0].scatter_(0, t, 1.0)
    out[0] *= (inputs[0] != ignore)
    return out


def contra_prep_lowres(label_l, label_u_aug, entropy, thresh, lo_idx, hi_idx, out_hw, num_classes,
                       negative_high_entropy=True, ignore=255):
    """Fused train_semi.py:408-465 -> (label_bits [P] int32, low_mask [Bl+Bu,1,h,w], high_mask [Bl+Bu,1,h,w])."""
    _need_cuda(label_l, label_u_aug, entropy, thresh)
    lib = _lib.load()
    Bl, H, W = label_l.shape
    Bu = label_u_aug.shape[0]
    h, w = out_hw
    P = (Bl + Bu) * h * w
    dev = label_l.device
    bits = torch.empty(P, dtype=torch.int32, device=dev)
    low = torch.empty((Bl + Bu, 1, h, w), dtype=torch.float32, device=dev)
    high = torch.empty((Bl + Bu, 1, h, w), dtype=torch.float32, device=dev)
    label_l, label_u_aug, entropy = label_l.contiguous(), label_u_aug.contiguous(), entropy.contiguous()
    assert label_l.dtype == torch.int64 and label_u_aug.dtype == torch.int64
    rc = lib.u2pl_contra_prep_lowres(_p(label_l), _p(label_u_aug), _p(entropy), _p(thresh), int(lo_idx), int(hi_idx),
                                     Bl, Bu, H, W, h, w, int(num_classes), int(ignore), int(bool(negative_high_entropy)),
                                     _p(bits), _p(low), _p(high), _stream())
    _lib.check(rc, "u2pl_contra_prep_lowres")
    return bits, low, high


def ohem_select(pred, target, thresh, min_kept, ignore=255):
    """Kept-pixel target of OhemCrossEntropy2dTensor (loss_helper.py:502-529).  Returns
    (new_target int64, kth_value fp32 device scalar, n_valid int64 device scalar)."""
    _need_cuda(pred, target)
    lib = _lib.load()
    pred = _f32c(pred.detach())
    target = target.contiguous()
    B, C, H, W = pred.shape
    new_target = torch.empty_like(target)
    kth = torch.empty((), dtype=torch.float32, device=pred.device)
    n_valid = torch.empty((), dtype=torch.int64, device=pred.device)
    ws = _workspace("entropy", lib.u2pl_entropy_ws_bytes(B, H * W), pred.device)
    rc = lib.u2pl_ohem_select(_p(pred), _p(target), B, C, H * W, int(ignore), float(thresh), int(min_kept),
                              _p(new_target), _p(kth), _p(n_valid), _p(ws), ws.numel(), _stream())
    _lib.check(rc, "u2pl_ohem_select")
    return new_target, kth, n_valid


def ohem_cross_entropy(pred, target, thresh, min_kept, ignore=255):
    """OhemCrossEntropy2dTensor.forward (loss_helper.py:502-531): mean CE over the kept pixels."""
    new_target, _, _ = ohem_select(pred, target, thresh, min_kept, ignore)
    return cross_entropy_mean(pred, new_target, ignore)


# --------------------------------------------------------------------------- tcgen05 GEMM (1x1 conv)
def gemm_bf16_tn(a, b, scale=None, shift=None, relu=False):
    """D = act((a @ b.T) * scale + shift): a [M,K], b [N,K] bf16 row-major -> D [M,N] bf16 (csrc/gemm_tc.cu)."""
    _need_cuda(a, b, scale, shift)
    lib = _lib.load()
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_contiguous() and b.is_contiguous()
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K
    d = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    rc = lib.u2pl_gemm_bf16_tn(_p(a), _p(b), _p(d), M, N, K, _p(scale), _p(shift), int(bool(relu)), _stream())
    _lib.check(rc, "u2pl_gemm_bf16_tn")
    return d


def conv1x1_bn_relu_eval(x, weight, scale, shift, relu=True):
    """conv1x1 + folded eval-mode BatchNorm + ReLU on a channels-last bf16 activation [N,Cin,H,W]."""
    N, Cin, H, W = x.shape
    assert x.is_contiguous(memory_format=torch.channels_last) and x.dtype == torch.bfloat16
    a = x.permute(0, 2, 3, 1).reshape(N * H * W, Cin)            # view: NHWC is the physical layout
    w = weight.reshape(weight.shape[0], Cin).to(torch.bfloat16).contiguous()
    d = gemm_bf16_tn(a, w, scale, shift, relu)
    return d.view(N, H, W, -1).permute(0, 3, 1, 2)               # channels-last view of [N,Cout,H,W]


_W_CACHE = {}          # id(parameter) -> (weakref, version, [Cout,k,k,Cin] bf16 tensor)


def kernel_weight(weight):
    """The [Cout, k, k, Cin] bf16 matrix the convolution kernels read.  For an nn.Parameter the converted copy is cached
    until the parameter's version counter moves (optimizer / EMA updates bump it; u2pl_b200.optim bumps it explicitly
    after the fused kernel wrote through raw pointers), so the teacher's two forwards of a step share one conversion."""
    import weakref
    if not isinstance(weight, torch.nn.Parameter):
        return weight.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous()     # no copy when already channels-last bf16
    key = id(weight)
    hit = _W_CACHE.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2].device == weight.device:
        return hit[2]
    wk = weight.detach().to(torch.bfloat16).permute(0, 2, 3, 1).contiguous()
    _W_CACHE[key] = (weakref.ref(weight), weight._version, wk)
    return wk


def conv_bf16_nhwc(x, weight, dilation=1, scale=None, shift=None, residual=None, relu=False):
    """Stride-1 "same" convolution (1x1 or 3x3, padding = dilation * (k // 2)) of a channels-last bf16 activation
    [N,Cin,H,W] as an implicit GEMM on the tensor cores, epilogue act(conv * scale + shift + residual)
    (csrc/conv_tc.cu).  `weight` [Cout,Cin,k,k] in any layout/dtype (a channels-last bf16 weight is used in place)."""
    _need_cuda(x, weight, scale, shift, residual)
    lib = _lib.load()
    N, Cin, H, W = x.shape
    Cout, k = weight.shape[0], weight.shape[2]
    assert x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last)
    assert weight.shape[1] == Cin and weight.shape[3] == k and k in (1, 3)
    wk = kernel_weight(weight)                                                # [Cout,k,k,Cin] bf16; cached per version
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.shape == (N, Cout, H, W) \
            and residual.is_contiguous(memory_format=torch.channels_last)
    out = torch.empty((N, Cout, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    rc = lib.u2pl_conv_bf16_nhwc(_p(x), _p(wk), _p(out), N, H, W, Cin, Cout, k, int(dilation), _p(scale), _p(shift),
                                 _p(residual), int(bool(relu)), _stream())
    _lib.check(rc, "u2pl_conv_bf16_nhwc")
    return out


def conv_bf16_nhwc_stats(x, weight, dilation=1):
    """Raw convolution output (bf16, channels-last) and its per-channel [sum | sum of squares] (fp32 [2, Cout]) from the
    same kernel pass -- what a train-mode BatchNorm needs before it can normalise (csrc/conv_tc.cu, kStats epilogue)."""
    _need_cuda(x, weight)
    lib = _lib.load()
    N, Cin, H, W = x.shape
    Cout, k = weight.shape[0], weight.shape[2]
    assert x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last)
    assert weight.shape[1] == Cin and weight.shape[3] == k and k in (1, 3)
    wk = kernel_weight(weight)
    out = torch.empty((N, Cout, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    parts = int(lib.u2pl_conv_stat_parts(N, H, W, k))
    part = torch.empty((parts, 2, Cout), dtype=torch.float32, device=x.device)
    sums = torch.empty((2, Cout), dtype=torch.float32, device=x.device)
    rc = lib.u2pl_conv_bf16_nhwc_stats(_p(x), _p(wk), _p(out), N, H, W, Cin, Cout, k, int(dilation), _p(part), _p(sums),
                                       _stream())
    _lib.check(rc, "u2pl_conv_bf16_nhwc_stats")
    return out, sums


def conv_wgrad_bf16_nhwc(x, gout, dilation=1):
    """Weight gradient [Cout, Cin, 3, 3] (fp32) of a stride-1 "same" 3x3 convolution from the channels-last bf16 input
    and output gradient, on the tensor cores with MN-major operands read in place (csrc/wgrad_tc.cu)."""
    _need_cuda(x, gout)
    lib = _lib.load()
    N, Cin, H, W = x.shape
    Cout = gout.shape[1]
    assert gout.shape == (N, Cout, H, W)
    for t in (x, gout):
        assert t.dtype == torch.bfloat16 and t.is_contiguous(memory_format=torch.channels_last)
    splits = int(lib.u2pl_conv_wgrad_splits(N, H, W, Cin, Cout))
    part = torch.empty((splits, 9, Cout, Cin), dtype=torch.float32, device=x.device)
    rc = lib.u2pl_conv_wgrad_bf16_nhwc(_p(x), _p(gout), _p(part), N, H, W, Cin, Cout, int(dilation), _stream())
    _lib.check(rc, "u2pl_conv_wgrad_bf16_nhwc")
    dw = part.sum(0) if splits > 1 else part[0]
    return dw.view(3, 3, Cout, Cin).permute(2, 3, 0, 1)


def conv_bf16_nhwc_ex(x, weight, dilation=1, in_scale=None, in_shift=None, in_relu=False, scale=None, shift=None,
                      residual=None, relu=False, want_stats=False):
    """General form of the implicit-GEMM convolution (u2pl_conv_bf16_nhwc_ex): optional act(x * in_scale + in_shift)
    applied to the input inside shared memory (the previous layer's BatchNorm + ReLU), optional output epilogue, optional
    per-channel [sum | sum of squares] of the stored output.  Returns out, or (out, sums) with want_stats."""
    _need_cuda(x, weight, in_scale, in_shift, scale, shift, residual)
    lib = _lib.load()
    N, Cin, H, W = x.shape
    Cout, k = weight.shape[0], weight.shape[2]
    assert x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last)
    assert weight.shape[1] == Cin and weight.shape[3] == k and k in (1, 3)
    wk = kernel_weight(weight)
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.shape == (N, Cout, H, W) \
            and residual.is_contiguous(memory_format=torch.channels_last)
    out = torch.empty((N, Cout, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    part = sums = None
    if want_stats:
        part = torch.empty((int(lib.u2pl_conv_stat_parts(N, H, W, k)), 2, Cout), dtype=torch.float32, device=x.device)
        sums = torch.empty((2, Cout), dtype=torch.float32, device=x.device)
    rc = lib.u2pl_conv_bf16_nhwc_ex(_p(x), _p(wk), _p(out), N, H, W, Cin, Cout, k, int(dilation), _p(in_scale), _p(in_shift),
                                    int(bool(in_relu)), _p(scale), _p(shift), _p(residual), int(bool(relu)), _p(part), _p(sums),
                                    _stream())
    _lib.check(rc, "u2pl_conv_bf16_nhwc_ex")
    return (out, sums) if want_stats else out
