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
