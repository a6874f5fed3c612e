"""CPU: the pure-torch parts of u2pl_b200/fused.py that the GPU step relies on -- the nine-GEMM weight gradient of
`_DilatedConvFn` (cropped / strided windows, zero padding skipped), `StemConv2d`'s channel padding and the module
walk of `run_sequential` -- against plain autograd of nn.Conv2d / nn.Sequential in fp32."""
import pytest
import torch
import torch.nn as nn

from u2pl_b200 import fused


@pytest.mark.parametrize("d,stride,H,W", [(1, 1, 9, 11), (2, 1, 13, 13), (12, 1, 17, 15), (36, 1, 33, 33), (1, 2, 17, 19),
                                          (2, 2, 16, 13), (4, 1, 7, 5)])
@pytest.mark.parametrize("stack", [False, True])
def test_dilated_conv_fn_matches_autograd(d, stride, H, W, stack, monkeypatch):
    monkeypatch.setitem(fused.ENABLED, "wgrad_stack", stack)
    torch.manual_seed(d * 10 + stride)
    x = torch.randn(2, 6, H, W).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = torch.randn(5, 6, 3, 3).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = fused._DilatedConvFn.apply(x, w, d, stride)
    g = torch.randn_like(y)
    y.backward(g)
    xr, wr = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr, None, stride, d, d)
    yr.backward(g)
    assert torch.allclose(y, yr, atol=1e-5)
    assert torch.allclose(x.grad, xr.grad, atol=1e-4)
    assert torch.allclose(w.grad, wr.grad, atol=1e-4), (w.grad - wr.grad).abs().max()


def test_run_sequential_equals_sequential_on_cpu():
    torch.manual_seed(0)
    seq = nn.Sequential(nn.Conv2d(4, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8), nn.ReLU(inplace=True), nn.Dropout2d(0.0),
                        nn.Conv2d(8, 8, 1, bias=False), nn.BatchNorm2d(8), nn.Conv2d(8, 3, 1))
    x = torch.randn(2, 4, 9, 9)
    for mode in (True, False):
        seq.train(mode)
        a = fused.run_sequential(seq, x.clone())
        seq2 = __import__("copy").deepcopy(seq)
        assert torch.allclose(a, seq(x.clone()), atol=1e-6)
        del seq2


def test_conv_bn_act_falls_back_off_gpu():
    conv, bn, relu = nn.Conv2d(8, 8, 3, padding=2, dilation=2, bias=False), nn.BatchNorm2d(8).eval(), nn.ReLU()
    x, r = torch.randn(1, 8, 7, 7), torch.randn(1, 8, 7, 7)
    with torch.no_grad():
        got = fused.conv_bn_act(x, conv, bn, relu, r)
        assert torch.allclose(got, relu(bn(conv(x)) + r), atol=1e-6)
    assert fused.ENABLED["tc_conv"] in ("auto", "1", False)       # per-layer policy by default (U2PL_TC_CONV=auto)


@pytest.mark.parametrize("k,d", [(1, 1), (3, 1), (3, 2), (3, 5)])
def test_dgrad_weight_identity(k, d):
    """The data gradient of a stride-1 'same' convolution is the convolution of the output gradient with
    fused.dgrad_weight(w) -- what _ConvTCFn.backward feeds to the implicit-GEMM kernel."""
    import torch.nn.functional as F
    torch.manual_seed(k + d)
    x = torch.randn(2, 5, 9, 11, requires_grad=True)
    w = torch.randn(7, 5, k, k)
    pad = d * (k // 2)
    y = F.conv2d(x, w, None, 1, pad, d)
    g = torch.randn_like(y)
    y.backward(g)
    dx = F.conv2d(g, fused.dgrad_weight(w), None, 1, pad, d)
    assert torch.allclose(dx, x.grad, atol=1e-4)


def test_bias_free_train_conv_in_front_of_batchnorm_equals_modules():
    """fused._conv_bias_bn_train (decoder.py:60-113: nn.Conv2d's default bias in front of a BatchNorm): same normalised
    output, running statistics and weight gradient as the plain modules, a ZERO bias gradient (the reference's is rounding
    noise around zero), on the CPU in fp32 where bn_act runs the module itself."""
    import torch.nn as nn
    from u2pl_b200 import fused
    torch.manual_seed(3)
    conv, bn, relu = nn.Conv2d(8, 16, 3, 1, 1, bias=True), nn.BatchNorm2d(16), nn.ReLU()
    conv_r, bn_r = nn.Conv2d(8, 16, 3, 1, 1, bias=True), nn.BatchNorm2d(16)
    with torch.no_grad():
        conv.bias.normal_(0, 2.0)
    conv_r.load_state_dict(conv.state_dict())
    bn_r.load_state_dict(bn.state_dict())
    x = torch.randn(3, 8, 9, 11)
    y = fused._conv_bias_bn_train(x, conv, bn, relu, None)
    yr = torch.relu(bn_r(conv_r(x)))
    assert torch.allclose(y, yr, atol=1e-5)
    assert torch.allclose(bn.running_mean, bn_r.running_mean, atol=1e-6) and torch.allclose(bn.running_var, bn_r.running_var, atol=1e-6)
    g = torch.randn_like(yr)
    y.backward(g)
    yr.backward(g)
    assert torch.allclose(conv.weight.grad, conv_r.weight.grad, atol=1e-4)
    assert float(conv.bias.grad.abs().max()) == 0.0 and float(conv_r.bias.grad.abs().max()) <= 1e-4
    assert torch.allclose(bn.weight.grad, bn_r.weight.grad, atol=1e-4) and torch.allclose(bn.bias.grad, bn_r.bias.grad, atol=1e-4)


def test_kernel_weight_cache_follows_the_version_counter():
    """ops.kernel_weight caches the [Cout,k,k,Cin] bf16 copy of a Parameter until its version counter moves: in-place
    updates (optimizer / EMA) and the explicit bump u2pl_b200.optim applies after the fused kernel wrote through raw
    pointers both invalidate it; plain tensors are never cached."""
    from u2pl_b200 import ops
    w = torch.nn.Parameter(torch.randn(4, 8, 3, 3))
    a = ops.kernel_weight(w)
    assert a.shape == (4, 3, 3, 8) and a.dtype == torch.bfloat16 and ops.kernel_weight(w) is a
    with torch.no_grad():
        w.mul_(2.0)
    b = ops.kernel_weight(w)
    assert b is not a and torch.equal(b, (w.detach().to(torch.bfloat16)).permute(0, 2, 3, 1).contiguous())
    with torch.no_grad():
        w.data.view(-1)[0] = 7.0                                  # a write the counter does not see (what a raw-pointer kernel does)
    assert ops.kernel_weight(w) is b
    torch.autograd.graph.increment_version(w)
    c = ops.kernel_weight(w)
    assert c is not b and float(c.reshape(-1)[0]) == 7.0
    t = torch.randn(4, 8, 1, 1)
    assert ops.kernel_weight(t) is not ops.kernel_weight(t)
