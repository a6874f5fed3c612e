"""Fused network blocks used by the model mirror on the GPU.

* `bn_act(x, bn, relu, residual)`: BatchNorm (+ReLU, +residual add) for channels-last bf16
  activations through the kernels of csrc/bn.cu (train and eval mode, SyncBN aware).  It reads the
  parameters / running statistics of the nn.BatchNorm2d / nn.SyncBatchNorm module it is given, so
  the module tree, state_dict keys and parameter order stay exactly the reference's
  (resnet.py:93-140, base.py:11-100, decoder.py:45-142).
* `run_sequential(seq, x)`: walks an nn.Sequential and fuses every (norm, ReLU) pair it finds.
* `DilatedConv2d`: nn.Conv2d whose weight gradient for large dilations is computed as nine cropped
  GEMMs.  cuDNN 9 picks `wgrad_alg0_engine_NHWC` (~43 TFLOP/s, 30 ms per layer per step on B200)
  for the three ASPP convolutions (2048 -> 256, dilation 12/24/36); the GEMM form also skips the
  taps' zero-padding region, which is most of the window at dilation 24/36 on a 65x65 map.

Selection rule: the fused path is taken for CUDA bf16 channels-last tensors (what the step runs
under autocast); fp32 tensors (CPU model-parity tests, `--fp32` runs) go through the nn.Modules.
"""
import ctypes

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .ops import _p, _stream

import os

def _env(name, default):
    return os.environ.get(name, default) == "1"


# Routing switches (env var = override for A/B runs; defaults are what B200 measurements picked, profiles/r02_*):
# "tc_conv"    eval-mode, no-grad conv + BN (+residual, +ReLU) as ONE tcgen05 implicit-GEMM kernel (csrc/conv_tc.cu).
#              "auto" (default) = per layer, where the fused kernel beats cuDNN conv + a separate bn_apply pass on B200
#              (every 1x1 convolution and the dilation >= 18 3x3s, `_tc_conv_wins`); "1" = every eligible layer; "0" = off.
# "tc_dilated" forward of 3x3 stride-1 convolutions with dilation >= 18 (ASPP d=24/36) through the tcgen05 kernel in
#              every mode: cuDNN 9 picks an sm80 kernel there (1.28 ms vs 0.52 ms per 16 images at 2048 -> 256).
# "wgrad_stack" stride-1 dilated weight gradient as ONE GEMM against nine shifted copies of the (small) output gradient
#              instead of nine GEMMs over cropped copies of the (large) input: -13 ms per V16 step.
# "tc_train"   train-mode forward AND data gradient of every eligible stride-1 convolution through conv_tc (statistics in
#              the epilogue); opt-in: the 1-CTA kernel runs at 0.65 of cuDNN's 2-CTA kernels, the step gets slower.
# "tc_t2"      train-mode forward WITHOUT autograd (the teacher's second forward, train_semi.py:362-364) of every eligible
#              stride-1 convolution through the flat tcgen05 kernel with the BatchNorm statistics taken in its epilogue
#              (no bn_stats pass over the output).
# "bias_fold"  convolution bias in front of a BatchNorm: folded into the fused epilogue (eval) / dropped from the train-mode
#              forward with the running mean corrected (`_conv_bias_bn_train`); no separate bias pass either way.
# "tc_wgrad"   3x3 stride-1 weight gradients of the DilatedConv2d layers via csrc/wgrad_tc.cu (MN-major operands read in
#              place, 0.74-1.16 PFLOP/s): -4.8 ms of backward per V16 step against the stacked-GEMM form.
# "tc_chain"   no-grad TRAIN-mode chains (teacher's second forward, train_semi.py:362-364): inner BatchNorm + ReLU applied
#              in the next convolution's operand load (conv_tc kXform), statistics from the producing conv's epilogue.
# "pool"       stem max-pooling through csrc/pool.cu instead of ATen's channels-last kernels (~0.6 TB/s).
ENABLED = {"bn": True, "wgrad": True,
           "tc_conv": os.environ.get("U2PL_TC_CONV", "auto"),
           "tc_dilated": _env("U2PL_TC_DILATED", "1"),
           "wgrad_stack": _env("U2PL_WGRAD_STACK", "1"),
           "tc_train": _env("U2PL_TC_TRAIN", "0"),
           "tc_t2": _env("U2PL_TC_T2", "0"),
           "bias_fold": _env("U2PL_BIAS_FOLD", "1"),
           "tc_wgrad": _env("U2PL_TC_WGRAD", "1"),
           "tc_chain": _env("U2PL_TC_CHAIN", "0"),
           "pool": _env("U2PL_POOL", "1")}
if ENABLED["tc_conv"] not in ("auto", "1", "1x1"):
    ENABLED["tc_conv"] = False
TC_DILATION_MIN = 18            # 3x3 convolutions at or above this dilation always run on conv_tc (see "tc_dilated")
FALLBACKS = {"bn_module": 0}    # calls that left the fused kernels for an nn.Module (reported by bench.py)


def _tc_conv_wins(conv, residual):
    """B200 measurements (profiles/r02_conv_bench.jsonl, 16 x 65x65 / 129x129): fused conv+BN(+res)+ReLU vs cuDNN conv +
    bn_apply.  1x1: the epilogue fusion always wins (44 vs 35+30 us at 1024->256; 88 vs 44+150 us at 256->1024 with
    the residual).  3x3: the 1-CTA kernel reaches 1.0-1.2 PF/s against cuDNN's 1.5-1.6 PF/s 2-CTA kernels, so only the
    large dilations (where cuDNN falls back to an sm80 kernel) win."""
    if ENABLED["tc_conv"] in ("1", True, "auto"):   # flat-tile kernel (conv_tc3.cu): at cuDNN's rate on every layer shape, so the
        return True                                 # fusion of BN (+bias, +residual, +ReLU) always wins; "0" / "1x1" for A/B runs
    if ENABLED["tc_conv"] == "1x1":                            # A/B switch: every 1x1 as well
        return conv.kernel_size == (1, 1) or conv.dilation[0] >= TC_DILATION_MIN
    # measured in the step (gpurun_out/r2b_*): routing every 1x1 of T1 through the 1-CTA kernel made T1 slower (22.1 vs
    # 18.2 ms) -- its per-thread 16-byte epilogue stores lose on the memory-bound 1x1 layers; only the large dilations win
    return conv.kernel_size == (3, 3) and conv.dilation[0] >= TC_DILATION_MIN


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _sync_sum_(sums, group=None):
    """Cross-rank sum of a [2,C] statistics tensor: one single-CTA kernel over peer-mapped memory on an NVSwitch box
    (csrc/peer_reduce.cu), torch.distributed otherwise."""
    from .peer import allreduce_small_
    return allreduce_small_(sums, group)                 # (a SyncBatchNorm built on a sub-group exchanges inside it, over NCCL)


def _is_cl_bf16(x):
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4
            and x.is_contiguous(memory_format=torch.channels_last))


def _bn_channels_ok(C, bn):
    return (ENABLED["bn"] and isinstance(bn, (nn.BatchNorm2d, nn.SyncBatchNorm))
            and C % 8 == 0 and C <= 2048 and 256 % (C // 8) == 0 and bn.momentum is not None
            and bn.track_running_stats)


def _bn_ok(x, bn):
    return _is_cl_bf16(x) and _bn_channels_ok(x.shape[1], bn)


def _scratch(dev, C):
    lib = _lib.load()
    parts = int(lib.u2pl_bn_parts())
    return torch.empty((parts, 2, C), dtype=torch.float32, device=dev), torch.empty((2, C), dtype=torch.float32, device=dev)


class _BNAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, bn, relu, sync, sums_in=None):
        lib = _lib.load()
        N, C, H, W = x.shape
        M = N * H * W
        dev = x.device
        scale = torch.empty(C, dtype=torch.float32, device=dev)
        shift = torch.empty(C, dtype=torch.float32, device=dev)
        training = bn.training
        count = float(M)
        if training:
            if sums_in is None:
                partial, sums = _scratch(dev, C)
                _lib.check(lib.u2pl_bn_stats(_p(x), M, C, _p(partial), _p(sums), _stream()), "u2pl_bn_stats")
            else:                                      # statistics already accumulated by the producing conv's epilogue
                sums = sums_in
            if sync:                                   # SyncBN: global sums; every rank holds the same number of
                group = getattr(bn, "process_group", None)
                _sync_sum_(sums, group)                # pixels (same per-GPU batch and crop, as in the reference configs)
                count = float(M) * _world(group)
            mean = torch.empty(C, dtype=torch.float32, device=dev)
            invstd = torch.empty(C, dtype=torch.float32, device=dev)
            _lib.check(lib.u2pl_bn_finalize(_p(sums), C, ctypes.c_double(count), _p(weight), _p(bias),
                                            _p(bn.running_mean), _p(bn.running_var), float(bn.momentum), float(bn.eps),
                                            _p(mean), _p(invstd), _p(scale), _p(shift), _stream()), "u2pl_bn_finalize")
            bn.num_batches_tracked.add_(1)
        else:
            mean = invstd = None
            _lib.check(lib.u2pl_bn_fold(C, _p(weight), _p(bias), _p(bn.running_mean), _p(bn.running_var), float(bn.eps),
                                        _p(scale), _p(shift), _stream()), "u2pl_bn_fold")
        y = torch.empty_like(x)                        # preserves channels-last
        _lib.check(lib.u2pl_bn_apply(_p(x), _p(residual), _p(scale), _p(shift), M, C, int(relu), _p(y), _stream()),
                   "u2pl_bn_apply")
        if training and any(ctx.needs_input_grad):
            ctx.save_for_backward(x, y if relu else None, weight, mean, invstd)
            ctx.meta = (M, C, count, sync, residual is not None, relu)
            ctx.group = getattr(bn, "process_group", None) if sync else None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, y, weight, mean, invstd = ctx.saved_tensors
        M, C, count, sync, has_res, relu = ctx.meta
        dy = dy.contiguous(memory_format=torch.channels_last)
        dev = x.device
        partial, sums = _scratch(dev, C)
        _lib.check(lib.u2pl_bn_backward_reduce(_p(dy), _p(x), _p(y), _p(mean), _p(invstd), M, C, _p(partial), _p(sums),
                                               _stream()), "u2pl_bn_backward_reduce")
        dweight, dbias = sums[1].clone(), sums[0].clone()            # local sums = this rank's parameter gradients
        if sync:
            _sync_sum_(sums, ctx.group)
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if has_res else None
        coef = torch.empty((3, C), dtype=torch.float32, device=dev)
        _lib.check(lib.u2pl_bn_backward_elemt(_p(dy), _p(x), _p(y), _p(mean), _p(invstd), _p(weight), _p(sums),
                                              ctypes.c_double(count), M, C, _p(coef), _p(dx), _p(dres), _stream()),
                   "u2pl_bn_backward_elemt")
        return dx, dweight, dbias, dres, None, None, None, None


def bn_act(x, bn, relu=None, residual=None, sums=None):
    """relu: an nn.ReLU module (or True) to fuse, or None.  residual: tensor added before the ReLU.
    sums: optional [2,C] batch sums of x from the producing convolution's epilogue (train mode)."""
    # (eval mode with autograd on -- frozen-BN fine-tuning, input gradients -- goes through the module: the fused eval
    # path keeps nothing for a backward pass)
    needs_eval_grad = (not bn.training) and torch.is_grad_enabled() and (x.requires_grad or bn.weight.requires_grad
                                                                         or (residual is not None and residual.requires_grad))
    if _bn_ok(x, bn) and (residual is None or _is_cl_bf16(residual)) and not needs_eval_grad:
        sync = isinstance(bn, nn.SyncBatchNorm) and bn.training and _world(getattr(bn, "process_group", None)) > 1
        return _BNAct.apply(x, bn.weight, bn.bias, residual, bn, relu is not None and relu is not False, sync,
                            sums if bn.training else None)
    if x.is_cuda:
        FALLBACKS["bn_module"] += 1                 # fp32 runs, Cout not a multiple of 8 (classifier), 1x1 pooled maps
    y = bn(x)
    if residual is not None:
        y = y + residual
    if relu is None or relu is False:
        return y
    return relu(y) if isinstance(relu, nn.Module) else F.relu(y)


def dgrad_weight(w):
    """Weight of the convolution that maps the output gradient of a stride-1 "same" convolution to its input gradient:
    w'[ci, co, r, s] = w[co, ci, k-1-r, k-1-s] (same dilation and padding)."""
    return w.transpose(0, 1).flip(2, 3)


class _ConvTCFn(torch.autograd.Function):
    """Stride-1 "same" convolution of channels-last bf16 tensors on the tcgen05 implicit-GEMM kernel, forward and data
    gradient; the weight gradient through ATen (cuDNN)."""

    @staticmethod
    def forward(ctx, x, w, dilation, want_stats=False):
        from .ops import conv_bf16_nhwc, conv_bf16_nhwc_stats
        ctx.save_for_backward(x, w)
        ctx.dilation = dilation
        if not want_stats:
            return conv_bf16_nhwc(x, w, dilation)
        y, sums = conv_bf16_nhwc_stats(x, w, dilation)
        ctx.mark_non_differentiable(sums)
        return y, sums

    @staticmethod
    def backward(ctx, gout, *unused):
        from .ops import conv_bf16_nhwc
        x, w = ctx.saved_tensors
        d, k = ctx.dilation, w.shape[2]
        gout = gout.contiguous(memory_format=torch.channels_last)
        dx = conv_bf16_nhwc(gout, dgrad_weight(w), d) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1] and ENABLED["tc_wgrad"] and k == 3:
            from .ops import conv_wgrad_bf16_nhwc
            dw = conv_wgrad_bf16_nhwc(x, gout, d).to(w.dtype)
        elif ctx.needs_input_grad[1]:
            pad = d * (k // 2)
            dw = torch.ops.aten.convolution_backward(gout, x, w, None, [1, 1], [pad, pad], [d, d], False, [0, 0], 1,
                                                     [False, True, False])[1]
        return dx, dw, None, None


class _ZeroBiasGrad(torch.autograd.Function):
    """Identity on y that hands the (unused) bias a zero gradient -- see _conv_bias_bn_train."""

    @staticmethod
    def forward(ctx, y, bias):
        ctx.shape, ctx.dev, ctx.dtype = bias.shape, bias.device, bias.dtype
        return y.view_as(y)

    @staticmethod
    def backward(ctx, g):
        return g, torch.zeros(ctx.shape, device=ctx.dev, dtype=ctx.dtype)


def _bias_fold_ok(x, conv, bn):
    return (ENABLED["bias_fold"] and type(conv) is nn.Conv2d and conv.bias is not None and bn.training and _is_cl_bf16(x)
            and isinstance(bn, (nn.BatchNorm2d, nn.SyncBatchNorm)) and _bn_channels_ok(conv.out_channels, bn))


def _conv_bias_bn_train(x, conv, bn, relu, residual):
    """Train-mode bn(conv(x) + b) without the bias pass (decoder.py:60-113: the heads' 3x3 convolutions and the low-level
    1x1 projection keep nn.Conv2d's default bias in front of a BatchNorm).  Batch normalisation removes the per-channel
    mean, so the normalised output does not depend on b, the gradient of b is identically zero (the reference computes
    rounding noise there; its weight decay still acts, so a zero gradient is returned rather than None), and only the
    running mean sees b: it is corrected by momentum * b after the statistics kernel updated it from the bias-free sums.
    Saves one read+write of the 129x129 activation per biased layer and direction (~0.27 ms each at 32 crops)."""
    w = conv.weight if torch.is_autocast_enabled() else conv.weight.to(x.dtype)      # (autocast casts once per context)
    y = F.conv2d(x, w, None, conv.stride, conv.padding, conv.dilation, conv.groups)
    if torch.is_grad_enabled() and conv.bias.requires_grad:
        y = _ZeroBiasGrad.apply(y, conv.bias)
    out = bn_act(y, bn, relu, residual)
    # (through .data: the module path of bn_act -- CPU / fp32 -- hands running_mean to native_batch_norm, whose train-mode
    # backward never reads it but whose saved-tensor version check would trip over a tracked in-place update)
    bn.running_mean.data.add_(conv.bias.detach().to(bn.running_mean.dtype), alpha=float(bn.momentum))
    return out


def _tc_geometry_ok(x, conv):
    k, d = conv.kernel_size[0], conv.dilation[0]
    if type(x).__name__ == "_ShapeProxy":             # output of an eligible stride-1 conv: channels-last bf16 by construction
        return _tc_geometry_ok(x.like, x.conv) and _geometry_only(conv)
    return (_is_cl_bf16(x) and isinstance(conv, nn.Conv2d) and conv.kernel_size in ((1, 1), (3, 3))
            and conv.stride == (1, 1) and conv.dilation == (d, d) and conv.padding == (d * (k // 2), d * (k // 2))
            and conv.groups == 1 and conv.bias is None and conv.padding_mode == "zeros"
            and conv.in_channels % 8 == 0 and conv.out_channels % 8 == 0)


def _geometry_only(conv):
    k, d = conv.kernel_size[0], conv.dilation[0]
    return (isinstance(conv, nn.Conv2d) and conv.kernel_size in ((1, 1), (3, 3)) and conv.stride == (1, 1)
            and conv.dilation == (d, d) and conv.padding == (d * (k // 2), d * (k // 2)) and conv.groups == 1
            and conv.bias is None and conv.padding_mode == "zeros" and conv.in_channels % 8 == 0 and conv.out_channels % 8 == 0)


def _tc_conv_ok(x, conv, bn, residual):
    k, d = conv.kernel_size[0], conv.dilation[0]
    return (ENABLED["tc_conv"] and not bn.training and not torch.is_grad_enabled() and _is_cl_bf16(x)
            and isinstance(conv, nn.Conv2d) and _tc_conv_wins(conv, residual) and isinstance(bn, (nn.BatchNorm2d, nn.SyncBatchNorm))
            and conv.kernel_size in ((1, 1), (3, 3)) and conv.stride == (1, 1) and conv.dilation == (d, d)
            and conv.padding == (d * (k // 2), d * (k // 2)) and conv.groups == 1
            and conv.padding_mode == "zeros" and conv.in_channels % 8 == 0 and conv.out_channels % 8 == 0
            and bn.track_running_stats and (residual is None or _is_cl_bf16(residual)))


def conv_bn_act(x, conv, bn, relu=None, residual=None):
    """relu(bn(conv(x)) + residual).  Eval mode without autograd (the teacher's pseudo-label forward) on channels-last
    bf16 activations: ONE implicit-GEMM kernel with the folded BatchNorm, the residual and the ReLU in its epilogue;
    otherwise the convolution module followed by `bn_act`."""
    if not _tc_conv_ok(x, conv, bn, residual):
        if _bias_fold_ok(x, conv, bn):
            return _conv_bias_bn_train(x, conv, bn, relu, residual)
        if (ENABLED["tc_t2"] and not torch.is_grad_enabled() and bn.training and isinstance(conv, nn.Conv2d)
                and _tc_geometry_ok(x, conv) and _bn_channels_ok(conv.out_channels, bn)):
            from .ops import conv_bf16_nhwc_stats
            y, sums = conv_bf16_nhwc_stats(x, conv.weight, conv.dilation[0])
            return bn_act(y, bn, relu, residual, sums=sums)
        if ENABLED["tc_train"] and type(conv) is nn.Conv2d and _tc_geometry_ok(x, conv):   # train mode / autograd on
            if bn.training and _bn_channels_ok(conv.out_channels, bn):
                y, sums = _ConvTCFn.apply(x, conv.weight.to(torch.bfloat16), conv.dilation[0], True)
                return bn_act(y, bn, relu, residual, sums=sums)
            y = _ConvTCFn.apply(x, conv.weight.to(torch.bfloat16), conv.dilation[0])
            return bn_act(y, bn, relu, residual)
        return bn_act(conv(x), bn, relu, residual)
    from .ops import conv_bf16_nhwc
    lib = _lib.load()
    C = conv.out_channels
    scale = torch.empty(C, dtype=torch.float32, device=x.device)
    shift = torch.empty(C, dtype=torch.float32, device=x.device)
    _lib.check(lib.u2pl_bn_fold(C, _p(bn.weight), _p(bn.bias), _p(bn.running_mean), _p(bn.running_var), float(bn.eps),
                                _p(scale), _p(shift), _stream()), "u2pl_bn_fold")
    if conv.bias is not None:                       # bn(conv + b) = conv * scale + (shift + scale * b)
        shift = torch.addcmul(shift, scale, conv.bias.detach().float())
    return conv_bf16_nhwc(x, conv.weight, conv.dilation[0], scale, shift, residual, relu is not None and relu is not False)


class _MaxPool3s2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        N, C, H, W = x.shape
        Ho, Wo = int(lib.u2pl_maxpool3s2_out(H)), int(lib.u2pl_maxpool3s2_out(W))
        y = torch.empty((N, C, Ho, Wo), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
        need_bwd = any(ctx.needs_input_grad)
        tap = torch.empty((N, Ho, Wo, C), dtype=torch.uint8, device=x.device) if need_bwd else None
        _lib.check(lib.u2pl_maxpool3s2_forward(_p(x), _p(y), _p(tap), N, H, W, C, _stream()), "u2pl_maxpool3s2_forward")
        if need_bwd:
            ctx.save_for_backward(tap)
            ctx.shape = (N, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        tap, = ctx.saved_tensors
        N, C, H, W = ctx.shape
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, C, H, W), dtype=torch.bfloat16, device=dy.device, memory_format=torch.channels_last)
        _lib.check(lib.u2pl_maxpool3s2_backward(_p(dy), _p(tap), _p(dx), N, H, W, C, _stream()), "u2pl_maxpool3s2_backward")
        return dx


def max_pool(x, pool):
    """The stem's nn.MaxPool2d(3, 2, 1, ceil_mode=True) on channels-last bf16 activations through csrc/pool.cu (opt-in)."""
    ok = (ENABLED["pool"] and _is_cl_bf16(x) and x.shape[1] % 8 == 0 and isinstance(pool, nn.MaxPool2d)
          and pool.kernel_size in (3, (3, 3)) and pool.stride in (2, (2, 2)) and pool.padding in (1, (1, 1))
          and pool.dilation in (1, (1, 1)) and pool.ceil_mode and not pool.return_indices)
    return _MaxPool3s2.apply(x) if ok else pool(x)


def _finalize_train_bn(bn, sums, count):
    """Batch statistics -> (scale, shift) of a train-mode BatchNorm, with the module's side effects (running statistics,
    num_batches_tracked, SyncBN all-reduce) -- the no-autograd half of _BNAct.forward."""
    lib = _lib.load()
    C = bn.num_features
    dev = sums.device
    group = getattr(bn, "process_group", None)
    if isinstance(bn, nn.SyncBatchNorm) and _world(group) > 1:
        _sync_sum_(sums, group)
        count = count * _world(group)
    mean, invstd = torch.empty(C, dtype=torch.float32, device=dev), torch.empty(C, dtype=torch.float32, device=dev)
    scale, shift = torch.empty(C, dtype=torch.float32, device=dev), torch.empty(C, dtype=torch.float32, device=dev)
    _lib.check(lib.u2pl_bn_finalize(_p(sums), C, ctypes.c_double(float(count)), _p(bn.weight), _p(bn.bias), _p(bn.running_mean),
                                    _p(bn.running_var), float(bn.momentum), float(bn.eps), _p(mean), _p(invstd), _p(scale),
                                    _p(shift), _stream()), "u2pl_bn_finalize")
    bn.num_batches_tracked.add_(1)
    return scale, shift


def chain_ok(x, convs, bns, residual=None):
    """All links of conv/BN pairs can run as one no-grad train-mode chain on the tensor-core kernel."""
    return (ENABLED["tc_chain"] and not torch.is_grad_enabled() and all(b.training for b in bns)
            and (residual is None or _is_cl_bf16(residual))
            and all(type(c) in (nn.Conv2d, DilatedConv2d) and _tc_geometry_ok(x if i == 0 else _ShapeProxy(x, convs[i - 1]), c)
                    for i, c in enumerate(convs))
            and all(_bn_channels_ok(c.out_channels, b) for c, b in zip(convs, bns)))


class _ShapeProxy:
    """Stands in for the (not yet computed) channels-last bf16 output of `conv` in eligibility checks."""

    def __init__(self, like, conv):
        self.like, self.conv = like, conv


def conv_bn_relu_chain(x, convs, bns, relu_last, residual=None):
    """relu?(bn_k(conv_k(... relu(bn_1(conv_1(x)))))) (+ residual before the last ReLU) in train mode without autograd.
    Link i's BatchNorm + ReLU is applied inside conv_{i+1}'s operand load; only the last BatchNorm is a separate pass
    (its output is the block's output).  Every convolution's epilogue delivers its own batch statistics."""
    from .ops import conv_bf16_nhwc_ex
    lib = _lib.load()
    in_scale = in_shift = None
    y = x
    for i, (conv, bn) in enumerate(zip(convs, bns)):
        y, sums = conv_bf16_nhwc_ex(y, conv.weight, conv.dilation[0], in_scale, in_shift, i > 0, want_stats=True)
        N, C, H, W = y.shape
        in_scale, in_shift = _finalize_train_bn(bn, sums, N * H * W)
    out = torch.empty_like(y)
    _lib.check(lib.u2pl_bn_apply(_p(y), _p(residual), _p(in_scale), _p(in_shift), N * H * W, C, int(bool(relu_last)), _p(out), _stream()),
               "u2pl_bn_apply")
    return out


def run_sequential(seq, x):
    """nn.Sequential.forward with (conv, norm[, ReLU]) and (norm, ReLU) groups fused."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Conv2d) and i + 1 < len(mods) and isinstance(mods[i + 1], (nn.BatchNorm2d, nn.SyncBatchNorm)) \
                and (_tc_conv_ok(x, m, mods[i + 1], None) or _bias_fold_ok(x, m, mods[i + 1])
                     or (ENABLED["tc_t2"] and not torch.is_grad_enabled() and mods[i + 1].training and _tc_geometry_ok(x, m))
                     or (ENABLED["tc_train"] and type(m) is nn.Conv2d and _tc_geometry_ok(x, m))):
            nxt = mods[i + 2] if i + 2 < len(mods) else None
            x = conv_bn_act(x, m, mods[i + 1], nxt if isinstance(nxt, nn.ReLU) else None)
            i += 3 if isinstance(nxt, nn.ReLU) else 2
            continue
        if isinstance(m, (nn.BatchNorm2d, nn.SyncBatchNorm)):
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if isinstance(nxt, nn.ReLU):
                x = bn_act(x, m, nxt)
                i += 2
                continue
            x = bn_act(x, m, None)
        else:
            x = m(x)
        i += 1
    return x


# ----------------------------------------------------------------------------- dilated conv
class _DilatedConvFn(torch.autograd.Function):
    """3x3 convolution, padding == dilation, stride s: cuDNN forward and data gradient, weight gradient as
    nine GEMMs over the (cropped, strided) input windows."""

    @staticmethod
    def forward(ctx, x, w, dilation, stride):
        ctx.save_for_backward(x, w)
        ctx.cfg = (dilation, stride)
        if stride == 1 and _is_cl_bf16(x) and (ENABLED["tc_train"] or (ENABLED["tc_dilated"] and dilation >= TC_DILATION_MIN)):
            from .ops import conv_bf16_nhwc
            return conv_bf16_nhwc(x, w, dilation)
        return F.conv2d(x, w, None, stride, dilation, dilation)

    @staticmethod
    def backward(ctx, gout):
        x, w = ctx.saved_tensors
        d, s = ctx.cfg
        gout = gout.contiguous(memory_format=torch.channels_last)
        dx = None
        if ctx.needs_input_grad[0] and ENABLED["tc_train"] and s == 1 and _is_cl_bf16(gout):
            from .ops import conv_bf16_nhwc
            dx = conv_bf16_nhwc(gout, dgrad_weight(w), d)
        elif ctx.needs_input_grad[0]:
            dx = torch.ops.aten.convolution_backward(gout, x, w, None, [s, s], [d, d], [d, d], False, [0, 0], 1,
                                                     [True, False, False])[0]
        dw = None
        if ctx.needs_input_grad[1]:
            N, Ci, H, W = x.shape
            Co, Ho, Wo = w.shape[0], gout.shape[2], gout.shape[3]
            xh = x.permute(0, 2, 3, 1)                   # NHWC views (no copies: both are channels-last)
            gh = gout.permute(0, 2, 3, 1)
            dw = torch.zeros((Co, 3, 3, Ci), dtype=torch.float32, device=x.device)

            def span(k, n_in, n_out):                    # output range whose input index s*o + k*d - d is in range
                off = k * d - d
                lo = max(0, -(off // s)) if off < 0 else 0
                hi = min(n_out - 1, (n_in - 1 - off) // s)
                return lo, hi, off

            if s == 1 and ENABLED["tc_wgrad"] and _is_cl_bf16(x) and _is_cl_bf16(gout):
                from .ops import conv_wgrad_bf16_nhwc
                return dx, conv_wgrad_bf16_nhwc(x, gout, d).to(w.dtype), None, None
            if s == 1 and ENABLED["wgrad_stack"]:
                # dw[co,ky,kx,ci] = sum_q g9[q, (ky,kx), co] * x[q, ci] with g9[q, tap] = g[q - offset(tap)] (zero where
                # that falls outside the map): x is read in place, only the Co-channel gradient is copied (9 shifted times)
                g9 = torch.zeros((N, H, W, 9, Co), dtype=gout.dtype, device=x.device)
                for ky in range(3):
                    y0, y1, oy = span(ky, H, Ho)
                    for kx in range(3):
                        x0, x1, ox = span(kx, W, Wo)
                        if y1 >= y0 and x1 >= x0:
                            g9[:, y0 + oy:y1 + oy + 1, x0 + ox:x1 + ox + 1, ky * 3 + kx, :] = gh[:, y0:y1 + 1, x0:x1 + 1, :]
                dw9 = torch.mm(g9.view(N * H * W, 9 * Co).t(), xh.reshape(N * H * W, Ci)).float()      # [9*Co, Ci]
                dw = dw9.view(3, 3, Co, Ci).permute(2, 3, 0, 1).to(w.dtype)
                return dx, dw, None, None
            for ky in range(3):
                y0, y1, oy = span(ky, H, Ho)
                if y1 < y0:
                    continue
                for kx in range(3):
                    x0, x1, ox = span(kx, W, Wo)
                    if x1 < x0:
                        continue
                    g = gh[:, y0:y1 + 1, x0:x1 + 1, :].reshape(-1, Co)              # zero padding contributes nothing:
                    a = xh[:, s * y0 + oy:s * y1 + oy + 1:s, s * x0 + ox:s * x1 + ox + 1:s, :].reshape(-1, Ci)   # crop to overlap
                    dw[:, ky, kx, :] = torch.mm(g.t(), a).float()
            dw = dw.permute(0, 3, 1, 2).to(w.dtype)
        return dx, dw, None, None


class DilatedConv2d(nn.Conv2d):
    """Same parameters / state_dict as nn.Conv2d; GEMM-based weight gradient for 3x3 convolutions with
    padding == dilation (any stride) on CUDA under bf16 autocast.  Used where cuDNN 9's heuristics fall back
    to wgrad_alg0_engine_NHWC on B200 (profiled: the three ASPP convs 2048->256 d=12/24/36, the 1280->256
    head conv, the stride-2 3x3 of layer2)."""

    def forward(self, x):
        d, s = self.dilation[0], self.stride[0]
        geom = (x.is_cuda and self.kernel_size == (3, 3) and self.stride == (s, s) and self.padding == (d, d)
                and self.dilation == (d, d) and self.groups == 1 and self.bias is None and torch.is_autocast_enabled())
        fast = ENABLED["wgrad"] and geom and torch.is_grad_enabled()
        if not fast:
            if (geom and ENABLED["tc_dilated"] and s == 1 and d >= TC_DILATION_MIN and self.in_channels % 8 == 0
                    and self.out_channels % 8 == 0):          # no-grad passes (teacher): same kernel, no autograd node
                from .ops import conv_bf16_nhwc
                xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
                return conv_bf16_nhwc(xb, self.weight, d)
            return super().forward(x)
        xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wb = self.weight.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", enabled=False):
            return _DilatedConvFn.apply(xb, wb, d, s)


class StemConv2d(nn.Conv2d):
    """The 3-channel input convolution (resnet.py:179).  With Cin = 3 cuDNN cannot use its 16-byte aligned
    bf16 kernels and falls back to wgrad_alg0_engine_NHWC (18 ms per step on B200 for 7 GFLOP).  Padding the
    input and the weight's Cin with zero channels up to 8 is mathematically the identity and takes the
    aligned paths; autograd slices the weight gradient back to 3 channels."""

    def forward(self, x):
        if not (x.is_cuda and self.in_channels < 8 and torch.is_autocast_enabled() and self.groups == 1):
            return super().forward(x)
        pad = 8 - self.in_channels
        x8 = F.pad(x, (0, 0, 0, 0, 0, pad)).contiguous(memory_format=torch.channels_last)
        w8 = F.pad(self.weight, (0, 0, 0, 0, 0, pad))
        return F.conv2d(x8, w8, self.bias, self.stride, self.padding, self.dilation, 1)
