"""CPU: the PYTHON GLUE around the tensor-core entry points (u2pl_b200/ops.py conv_bf16_nhwc / conv_bf16_nhwc_stats /
conv_wgrad_bf16_nhwc, fused.conv_bn_act's eval path, fused._ConvTCFn forward / data gradient / weight gradient) with
the C ABI EMULATED in the test: a fake library object receives the same raw pointers and integer arguments the real
libu2pl_b200.so would, reinterprets the host memory exactly as the header documents it (x [n,h,w,cin], weight
[cout,k,k,cin], out [n,h,w,cout], partial [splits,9,cout,cin] ...) and computes with torch.  The kernels themselves are
checked against CPU loops on the GPU (tools/cu/tc_selftest.cu); what this pins is the layer in between -- argument
order, physical layouts, permutes, autograd plumbing -- which otherwise only a GPU run would exercise.
Nothing here is a product path: the product has no CPU fallback, the emulation exists only inside this test."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import emulated_abi
from u2pl_b200 import fused, ops


@pytest.fixture
def emulated(monkeypatch):
    return emulated_abi.install(monkeypatch)


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("k,d", [(1, 1), (3, 1), (3, 4)])
def test_ops_conv_layouts(emulated, k, d):
    torch.manual_seed(k + d)
    x = _cl(torch.randn(2, 16, 9, 7).bfloat16())
    w_cl = _cl((torch.randn(24, 16, k, k) / 4).bfloat16())
    w_plain = w_cl.contiguous()                                            # NCHW-contiguous weight must give the same result
    scale, shift = torch.rand(24) + 0.5, torch.randn(24)
    res = _cl(torch.randn(2, 24, 9, 7).bfloat16())
    ref = F.relu(F.conv2d(x.float(), w_cl.float(), None, 1, d * (k // 2), d) * scale[None, :, None, None]
                 + shift[None, :, None, None] + res.float())
    for w in (w_cl, w_plain, w_plain.float()):
        y = ops.conv_bf16_nhwc(x, w, d, scale, shift, res, True)
        assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
        assert (y.float() - ref).abs().max() <= 2e-2 * ref.abs().max()
    y2, sums = ops.conv_bf16_nhwc_stats(x, w_cl, d)
    raw = F.conv2d(x.float(), w_cl.float(), None, 1, d * (k // 2), d)
    assert (y2.float() - raw).abs().max() <= 2e-2 * raw.abs().max()
    assert torch.allclose(sums[0], y2.float().sum(dim=(0, 2, 3)), atol=1e-3) and sums.shape == (2, 24)


def test_ops_conv_ex_argument_order(emulated):
    torch.manual_seed(9)
    x = _cl(torch.randn(2, 16, 9, 7).bfloat16())
    w = _cl((torch.randn(24, 16, 3, 3) / 4).bfloat16())
    isc, ish = torch.rand(16) + 0.5, torch.randn(16)
    sc, sh = torch.rand(24) + 0.5, torch.randn(24)
    res = _cl(torch.randn(2, 24, 9, 7).bfloat16())
    z = F.relu(x.float() * isc[None, :, None, None] + ish[None, :, None, None]).bfloat16().float()
    ref = F.relu(F.conv2d(z, w.float(), None, 1, 2, 2) * sc[None, :, None, None] + sh[None, :, None, None] + res.float())
    y, sums = ops.conv_bf16_nhwc_ex(x, w, 2, isc, ish, True, sc, sh, res, True, want_stats=True)
    assert (y.float() - ref).abs().max() <= 2e-2 * ref.abs().max()
    assert torch.allclose(sums[1], (y.float() ** 2).sum(dim=(0, 2, 3)), rtol=1e-4)
    y2 = ops.conv_bf16_nhwc_ex(x, w, 2)                                     # everything optional off == plain convolution
    assert (y2.float() - F.conv2d(x.float(), w.float(), None, 1, 2, 2)).abs().max() <= 2e-2 * 8


def test_ops_wgrad_layout(emulated):
    torch.manual_seed(0)
    x = _cl(torch.randn(2, 16, 9, 7).bfloat16())
    g = _cl(torch.randn(2, 8, 9, 7).bfloat16())
    dw = ops.conv_wgrad_bf16_nhwc(x, g, 2)
    w = torch.zeros(8, 16, 3, 3, requires_grad=True)
    F.conv2d(x.float(), w, None, 1, 2, 2).backward(g.float())
    assert dw.shape == (8, 16, 3, 3) and torch.allclose(dw, w.grad, atol=1e-3)


@pytest.mark.parametrize("k,d,wgrad", [(3, 2, False), (3, 2, True), (1, 1, False)])
def test_conv_tc_fn_autograd_plumbing(emulated, monkeypatch, k, d, wgrad):
    monkeypatch.setitem(fused.ENABLED, "tc_wgrad", wgrad)
    torch.manual_seed(k)
    x = _cl(torch.randn(2, 16, 9, 7).bfloat16()).requires_grad_(True)
    wm = nn.Parameter(_cl(torch.randn(8, 16, k, k) / 4))                   # fp32 master weight, channels-last
    y = fused._ConvTCFn.apply(x, wm.to(torch.bfloat16), d)
    g = _cl(torch.randn_like(y))
    y.backward(g)
    xr = x.detach().float().requires_grad_(True)
    wr = wm.detach().bfloat16().float().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, 1, d * (k // 2), d)
    yr.backward(g.float())
    assert (y.float() - yr).abs().max() <= 2e-2 * yr.abs().max()
    assert (x.grad.float() - xr.grad).abs().max() <= 2e-2 * xr.grad.abs().max()
    assert wm.grad.dtype == torch.float32 and (wm.grad - wr.grad).abs().max() <= 2e-2 * wr.grad.abs().max()
    ys, sums = fused._ConvTCFn.apply(x, wm.to(torch.bfloat16), d, True)   # statistics variant: same output, extra sums
    assert torch.equal(ys, y) and sums.shape == (2, 8) and not sums.requires_grad


def test_conv_bn_act_eval_path(emulated, monkeypatch):
    monkeypatch.setitem(fused.ENABLED, "tc_conv", True)
    torch.manual_seed(3)
    conv = nn.Conv2d(16, 24, 3, padding=2, dilation=2, bias=False).to(memory_format=torch.channels_last)
    bn = nn.BatchNorm2d(24).eval()
    with torch.no_grad():
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0); bn.weight.uniform_(0.5, 1.5); bn.bias.normal_()
    x = _cl(torch.randn(2, 16, 9, 7).bfloat16())
    r = _cl(torch.randn(2, 24, 9, 7).bfloat16())
    with torch.no_grad():
        assert fused._tc_conv_ok(x, conv, bn, r)
        got = fused.conv_bn_act(x, conv, bn, nn.ReLU(), r)
        ref = F.relu(bn(F.conv2d(x.float(), conv.weight.bfloat16().float(), None, 1, 2, 2)) + r.float())
        seq = nn.Sequential(conv, bn, nn.ReLU(), nn.Conv2d(24, 24, 1, bias=False).to(memory_format=torch.channels_last),
                            nn.BatchNorm2d(24).eval())
        out = fused.run_sequential(seq, x)
        ref2 = seq[4](F.conv2d(F.relu(bn(F.conv2d(x.float(), conv.weight.bfloat16().float(), None, 1, 2, 2))).bfloat16().float(),
                               seq[3].weight.bfloat16().float()))
    assert got.dtype == torch.bfloat16 and (got.float() - ref).abs().max() <= 3e-2 * ref.abs().max()
    assert (out.float() - ref2).abs().max() <= 3e-2 * ref2.abs().max()
    with torch.enable_grad():                                              # autograd on -> not the fused eval path
        assert not fused._tc_conv_ok(x, conv, bn, None)


@pytest.mark.parametrize("train,relu,res", [(True, True, True), (True, False, False), (False, True, True), (True, True, False)])
def test_bn_act_function_matches_batchnorm(emulated, train, relu, res):
    """fused._BNAct (the DEFAULT path's Python) over the emulated csrc/bn.cu ABI against nn.BatchNorm2d autograd."""
    torch.manual_seed(int(train) * 4 + int(relu) * 2 + int(res))
    C = 16
    x32 = torch.randn(3, C, 7, 5).bfloat16().float()
    r32 = torch.randn(3, C, 7, 5).bfloat16().float() if res else None
    bn, ref = nn.BatchNorm2d(C), nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2)
    ref.load_state_dict(bn.state_dict())
    bn.train(train); ref.train(train)
    xa = _cl(x32.bfloat16()).requires_grad_(True)
    ra = _cl(r32.bfloat16()).requires_grad_(True) if res else None
    y = fused.bn_act(xa, bn, nn.ReLU() if relu else None, ra)
    xb = x32.clone().requires_grad_(True)
    rb = r32.clone().requires_grad_(True) if res else None
    yr = ref(xb)
    yr = yr + rb if res else yr
    yr = F.relu(yr) if relu else yr
    assert y.dtype == torch.bfloat16 and (y.float() - yr).abs().max() <= 3e-2 * max(1.0, yr.abs().max().item())
    assert torch.allclose(bn.running_mean, ref.running_mean, atol=1e-4) and torch.allclose(bn.running_var, ref.running_var, atol=1e-3)
    if train:
        g = torch.randn_like(yr).bfloat16().float()
        y.backward(_cl(g.bfloat16()))
        yr.backward(g)
        assert (xa.grad.float() - xb.grad).abs().max() <= 4e-2 * max(1.0, xb.grad.abs().max().item())
        assert torch.allclose(bn.weight.grad, ref.weight.grad, atol=3e-2 * ref.weight.grad.abs().max().item())
        assert torch.allclose(bn.bias.grad, ref.bias.grad, atol=3e-2 * ref.bias.grad.abs().max().item())
        if res:
            assert (ra.grad.float() - rb.grad).abs().max() <= 2e-2 * max(1.0, rb.grad.abs().max().item())
        assert int(bn.num_batches_tracked) == 1


def test_train_mode_conv_stats_bn_chain(emulated, monkeypatch):
    """U2PL_TC_TRAIN routing: conv (+statistics epilogue) -> _BNAct consuming those sums, against Conv2d + BatchNorm2d."""
    monkeypatch.setitem(fused.ENABLED, "tc_train", True)
    torch.manual_seed(5)
    conv = nn.Conv2d(16, 24, 3, padding=2, dilation=2, bias=False).to(memory_format=torch.channels_last)
    bn = nn.BatchNorm2d(24)
    conv_r, bn_r = nn.Conv2d(16, 24, 3, padding=2, dilation=2, bias=False), nn.BatchNorm2d(24)
    with torch.no_grad():
        conv.weight.copy_(conv.weight.bfloat16().float())
    conv_r.load_state_dict(conv.state_dict()); bn_r.load_state_dict(bn.state_dict())
    x32 = torch.randn(2, 16, 9, 7).bfloat16().float()
    xa = _cl(x32.bfloat16()).requires_grad_(True)
    calls = []
    real = emulated.u2pl_bn_stats
    monkeypatch.setattr(emulated, "u2pl_bn_stats", lambda *a: (calls.append(1), real(*a))[1])
    y = fused.conv_bn_act(xa, conv, bn, nn.ReLU())
    assert not calls                                                       # statistics came from the conv epilogue
    xb = x32.clone().requires_grad_(True)
    yr = F.relu(bn_r(conv_r(xb)))
    assert (y.float() - yr).abs().max() <= 4e-2 * max(1.0, yr.abs().max().item())
    g = torch.randn_like(yr).bfloat16().float()
    y.backward(_cl(g.bfloat16()))
    yr.backward(g)
    assert (xa.grad.float() - xb.grad).abs().max() <= 6e-2 * max(1.0, xb.grad.abs().max().item())
    assert (conv.weight.grad - conv_r.weight.grad).abs().max() <= 6e-2 * conv_r.weight.grad.abs().max().item()
    assert torch.allclose(bn.running_var, bn_r.running_var, atol=2e-3)


def _small_net():
    import copy
    import u2pl_b200
    u2pl_b200.install()
    from u2pl.models.model_helper import ModelBuilder
    net = {"num_classes": 21, "sync_bn": False, "ema_decay": 0.99,
           "encoder": {"type": "u2pl.models.resnet.resnet50",
                       "kwargs": {"multi_grid": True, "zero_init_residual": False, "fpn": True,
                                  "replace_stride_with_dilation": [False, True, True], "pretrained": False}},
           "decoder": {"type": "u2pl.models.decoder.dec_deeplabv3_plus", "kwargs": {"inner_planes": 256, "dilations": [12, 24, 36]}}}
    torch.manual_seed(0)
    m = ModelBuilder(copy.deepcopy(net)).to(memory_format=torch.channels_last)
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
    return m


def test_whole_network_eval_routing(emulated, monkeypatch):
    """Mirror network, eval / no_grad / bf16 autocast: every eligible conv+BN(+residual)(+ReLU) group through the fused
    tensor-core entry point (U2PL_TC_CONV) vs the default routing (conv module + BN kernel) -- same network output."""
    m = _small_net().eval()
    x = _cl(torch.randn(1, 3, 49, 49))
    outs, convs = {}, []
    real = emulated.u2pl_conv_bf16_nhwc
    monkeypatch.setattr(emulated, "u2pl_conv_bf16_nhwc", lambda *a: (convs.append(a[8:10]), real(*a))[1])
    for flag in (False, True):
        monkeypatch.setitem(fused.ENABLED, "tc_conv", flag)
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            outs[flag] = {k: v.float() for k, v in m(x).items()}
    assert len(convs) >= 50                                               # 16 bottlenecks x 3 + downsamples + ASPP + decoder
    assert {int(d[1]) for d in convs} >= {1, 2, 4, 8, 16, 12, 24, 36}      # every dilation of the network went through it
    for k in ("pred", "rep"):
        a, b = outs[False][k], outs[True][k]
        assert (a - b).norm() <= 0.05 * a.norm(), (k, float((a - b).norm() / a.norm()))


def test_whole_encoder_train_routing(emulated, monkeypatch):
    """Train mode with autograd: U2PL_TC_TRAIN (+ tensor-core weight gradients) routing vs the default routing."""
    import copy
    ma = _small_net().train()
    mb = copy.deepcopy(ma)
    x = _cl(torch.randn(2, 3, 33, 33))
    res = []
    for m, flag in ((ma, False), (mb, True)):
        monkeypatch.setitem(fused.ENABLED, "tc_train", flag)
        monkeypatch.setitem(fused.ENABLED, "tc_wgrad", flag)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            feats = m.encoder(x)
        sum(t.float().pow(2).mean() for t in feats).backward()
        res.append((feats, dict(m.named_parameters())))
    for a, b in zip(res[0][0], res[1][0]):
        assert (a.float() - b.float()).norm() <= 0.05 * a.float().norm()
    for name in ("encoder.layer4.2.conv2.weight", "encoder.layer3.1.conv1.weight", "encoder.layer1.0.bn1.weight", "encoder.conv1.3.weight"):
        ga, gb = res[0][1][name].grad.float(), res[1][1][name].grad.float()
        assert (ga - gb).norm() <= 0.15 * ga.norm(), (name, float((ga - gb).norm() / ga.norm()))


def test_nograd_train_chain_routing(emulated, monkeypatch):
    """U2PL_TC_CHAIN: the teacher's no-grad TRAIN-mode forward with the inner BatchNorm+ReLU of every bottleneck applied in
    the next convolution's operand load -- same features, same running statistics, a third of the BN-apply passes."""
    import copy
    ma = _small_net().train()
    mb = copy.deepcopy(ma)
    x = _cl(torch.randn(2, 3, 33, 33))
    applies, res = [], []
    real = emulated.u2pl_bn_apply
    monkeypatch.setattr(emulated, "u2pl_bn_apply", lambda *a: (applies.append(1), real(*a))[1])
    for m, flag in ((ma, False), (mb, True)):
        monkeypatch.setitem(fused.ENABLED, "tc_chain", flag)
        n0 = len(applies)
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            feats = m.encoder(x)
        res.append((feats, dict(m.named_buffers()), len(applies) - n0))
    for a, b in zip(res[0][0], res[1][0]):
        assert (a.float() - b.float()).norm() <= 0.05 * a.float().norm()
    for name in ("encoder.layer2.1.bn2.running_var", "encoder.layer4.0.bn1.running_mean", "encoder.layer3.2.bn3.running_var"):
        va, vb = res[0][1][name], res[1][1][name]
        assert (va - vb).abs().max() <= 0.03 * max(1e-3, va.abs().max().item()), name
    assert int(res[1][1]["encoder.layer3.2.bn2.num_batches_tracked"]) == 1
    assert res[1][2] <= res[0][2] - 2 * 14                                 # >= 14 of the 16 bottlenecks chained: 2 fewer passes each


@pytest.mark.parametrize("H,W", [(17, 19), (16, 12), (9, 9)])
def test_max_pool_function(emulated, monkeypatch, H, W):
    """fused.max_pool / _MaxPool3s2 (csrc/pool.cu routing) vs nn.MaxPool2d autograd, with ReLU-like inputs full of ties."""
    monkeypatch.setitem(fused.ENABLED, "pool", True)
    torch.manual_seed(H)
    pool = nn.MaxPool2d(3, 2, 1, ceil_mode=True)
    x32 = F.relu(torch.randn(2, 16, H, W)).bfloat16().float()                 # many exact zeros and bf16 ties
    xa = _cl(x32.bfloat16()).requires_grad_(True)
    y = fused.max_pool(xa, pool)
    xb = x32.clone().requires_grad_(True)
    yr = pool(xb)
    assert y.shape == yr.shape and torch.equal(y.float(), yr)
    g = torch.randn_like(yr).bfloat16().float()
    y.backward(_cl(g.bfloat16()))
    yr.backward(g)
    assert (xa.grad.float() - xb.grad).abs().max() <= 2e-2 * xb.grad.abs().max()  # bf16 rounding of sums of <= 4 terms
    with torch.no_grad():
        assert torch.equal(fused.max_pool(xa.detach(), pool).float(), yr)      # no tap map without autograd
    monkeypatch.setitem(fused.ENABLED, "pool", False)
    assert torch.equal(fused.max_pool(xa.detach(), pool).float(), yr)


def test_smoke_tensor_core_helper_runs(emulated, monkeypatch, capsys):
    """__graft_entry__._smoke_tensor_core: the informational tensor-core check smoke() appends on the GPU box, exercised
    here on the CPU device over the emulated ABI (so that a typo cannot turn it into a false alarm there)."""
    import __graft_entry__ as G
    from u2pl_b200 import _lib
    monkeypatch.setattr(_lib, "launch_count", lambda: 0)
    assert G._smoke_tensor_core(torch, ops, _lib, device="cpu") is True
    assert "smoke tensor-core ok" in capsys.readouterr().out
