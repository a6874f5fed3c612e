"""GPU: fused BatchNorm(+ReLU,+residual) kernels and the GEMM-based dilated-conv weight gradient against
plain PyTorch fp32 references of the same ops (inputs are bf16-representable; outputs are bf16, so the
tolerance is bf16 rounding: 1e-2 relative on activations / gradients, 1e-5 on fp32 statistics)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("C,shape,relu,res", [(64, (3, 17, 19), True, False), (256, (2, 9, 11), True, True),
                                              (2048, (2, 5, 7), False, False), (1024, (4, 13, 13), True, True),
                                              (128, (1, 33, 35), False, True)])
def test_bn_act_train_matches_torch(C, shape, relu, res):
    from u2pl_b200 import fused
    torch.manual_seed(C)
    N, H, W = shape
    x32 = (torch.randn(N, C, H, W, device="cuda") * 1.7 + 0.3).bfloat16().float()
    r32 = torch.randn(N, C, H, W, device="cuda").bfloat16().float() if res else None
    bn = nn.BatchNorm2d(C).cuda()
    ref = nn.BatchNorm2d(C).cuda()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
        ref.load_state_dict(bn.state_dict())
    xa = _cl(x32.bfloat16()).requires_grad_(True)
    ra = _cl(r32.bfloat16()).requires_grad_(True) if res else None
    y = fused.bn_act(xa, bn, nn.ReLU() if relu else None, ra)
    assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    xb = x32.clone().requires_grad_(True)
    rb = r32.clone().requires_grad_(True) if res else None
    yr = ref(xb)
    if res:
        yr = yr + rb
    if relu:
        yr = F.relu(yr)
    assert (y.float() - yr).abs().max() <= 2e-2 * max(1.0, yr.abs().max().item())
    assert torch.allclose(bn.running_mean, ref.running_mean, atol=1e-5, rtol=1e-5)
    assert torch.allclose(bn.running_var, ref.running_var, atol=1e-5, rtol=1e-4)
    assert int(bn.num_batches_tracked) == 1
    g = torch.randn_like(yr).bfloat16().float()
    y.backward(_cl(g.bfloat16()))
    yr.backward(g)
    scale = max(1.0, xb.grad.abs().max().item())
    assert (xa.grad.float() - xb.grad).abs().max() <= 3e-2 * scale
    if res:
        assert (ra.grad.float() - rb.grad).abs().max() <= 1e-2 * max(1.0, rb.grad.abs().max().item())
    assert torch.allclose(bn.weight.grad, ref.weight.grad, atol=2e-2 * ref.weight.grad.abs().max().item(), rtol=2e-2)
    assert torch.allclose(bn.bias.grad, ref.bias.grad, atol=2e-2 * ref.bias.grad.abs().max().item(), rtol=2e-2)


def test_bn_act_eval_and_fallbacks():
    from u2pl_b200 import fused
    torch.manual_seed(0)
    bn = nn.BatchNorm2d(64).cuda().eval()
    with torch.no_grad():
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2); bn.weight.uniform_(0.5, 1.5); bn.bias.normal_()
    x = torch.randn(2, 64, 7, 9, device="cuda").bfloat16()
    y = fused.bn_act(_cl(x), bn, nn.ReLU())
    ref = F.relu(bn(x.float()))
    assert (y.float() - ref).abs().max() <= 2e-2 * max(1.0, ref.abs().max().item())
    # fp32 / NCHW inputs keep using the nn.Module (exact)
    x32 = torch.randn(2, 64, 7, 9, device="cuda")
    assert torch.equal(fused.bn_act(x32, bn, nn.ReLU()), F.relu(bn(x32)))


@pytest.mark.parametrize("d,stride", [(12, 1), (24, 1), (36, 1), (1, 1), (1, 2), (2, 2)])
def test_dilated_conv_weight_grad(d, stride):
    from u2pl_b200 import fused
    torch.manual_seed(d)
    conv = fused.DilatedConv2d(64, 32, 3, stride=stride, padding=d, dilation=d, bias=False).cuda()
    ref = nn.Conv2d(64, 32, 3, stride=stride, padding=d, dilation=d, bias=False).cuda()
    ref.load_state_dict(conv.state_dict())
    x = torch.randn(2, 64, 33, 31, device="cuda").bfloat16().float()
    xa = _cl(x.clone()).requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ya = conv(xa)
    yb = ref(xb)
    assert (ya.float() - yb).abs().max() <= 2e-2 * yb.abs().max().item()
    g = torch.randn_like(yb).bfloat16().float()
    ya.backward(_cl(g.bfloat16()))
    yb.backward(g)
    assert (conv.weight.grad - ref.weight.grad).abs().max() <= 2e-2 * ref.weight.grad.abs().max().item()
    assert (xa.grad - xb.grad).abs().max() <= 2e-2 * xb.grad.abs().max().item()


def test_model_bf16_fused_close_to_fp32():
    """Whole mirror network: fused bf16 path vs its own fp32 module path (same weights).
    eval mode: full outputs.  train mode: encoder features and gradients (the ASPP image-pool branch
    batch-normalises N values per channel at 1x1 resolution -- with N=4 its output sign is decided by
    rounding noise, so full train-mode outputs are not comparable across precisions)."""
    import copy
    import u2pl_b200
    u2pl_b200.install()
    from u2pl.models.model_helper import ModelBuilder
    net = {"num_classes": 21, "sync_bn": False, "ema_decay": 0.99,
           "encoder": {"type": "u2pl.models.resnet.resnet50",
                       "kwargs": {"multi_grid": True, "zero_init_residual": False, "fpn": True,
                                  "replace_stride_with_dilation": [False, True, True], "pretrained": False}},
           "decoder": {"type": "u2pl.models.decoder.dec_deeplabv3_plus", "kwargs": {"inner_planes": 256, "dilations": [12, 24, 36]}}}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    m32 = ModelBuilder(copy.deepcopy(net)).cuda()
    mbf = copy.deepcopy(m32).to(memory_format=torch.channels_last)
    x = torch.randn(4, 3, 97, 97, device="cuda")
    # ---- train mode, encoder only.  Control = the same network under autocast with ATen's own BN/ReLU
    # (fused path switched off): the fused kernels must not be further from fp32 than that control.
    from u2pl_b200 import fused
    f32 = m32.encoder(x)
    mctl = copy.deepcopy(mbf)
    fused.ENABLED["bn"] = fused.ENABLED["wgrad"] = False
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            fctl = mctl.encoder(_cl(x))
        sum(t.float().pow(2).mean() for t in fctl).backward()
    finally:
        fused.ENABLED["bn"] = fused.ENABLED["wgrad"] = True
    with torch.autocast("cuda", dtype=torch.bfloat16):
        fbf = mbf.encoder(_cl(x))
    for a, c, b in zip(fbf, fctl, f32):
        err, ctl = (a.float() - b).norm() / b.norm(), (c.float() - b).norm() / b.norm()
        assert err <= max(1.5 * ctl, 0.02), (float(err), float(ctl))
    sum(t.float().pow(2).mean() for t in f32).backward()
    sum(t.float().pow(2).mean() for t in fbf).backward()
    p32, pbf, pctl = dict(m32.named_parameters()), dict(mbf.named_parameters()), dict(mctl.named_parameters())
    for name in ("encoder.layer4.2.conv2.weight", "encoder.layer3.2.conv2.weight", "encoder.layer1.0.bn1.weight",
                 "encoder.conv1.0.weight", "encoder.layer2.0.downsample.0.weight"):
        b = p32[name].grad
        err = (pbf[name].grad.float() - b).norm() / b.norm()
        ctl = (pctl[name].grad.float() - b).norm() / b.norm()
        assert err <= max(1.5 * ctl, 0.05), (name, float(err), float(ctl))
    rm32 = dict(m32.named_buffers())["encoder.layer2.1.bn2.running_var"]
    rmbf = dict(mbf.named_buffers())["encoder.layer2.1.bn2.running_var"]
    assert (rm32 - rmbf).abs().max() <= 0.05 * rm32.abs().max().item()
    # ---- eval mode, whole network (running statistics folded into scale/shift)
    m32.eval(); mbf.eval()
    with torch.no_grad():
        out32 = m32(x)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outbf = mbf(_cl(x))
    for k in ("pred", "rep"):
        assert (outbf[k].float() - out32[k]).norm() <= 0.1 * out32[k].norm(), k


@pytest.mark.parametrize("k", [1, 3])
def test_conv_bias_in_front_of_batchnorm(k):
    """decoder.py:60-113 keeps nn.Conv2d's default bias in front of a BatchNorm.  Train mode: the fused path drops the bias
    pass (the normalised output does not depend on it), corrects the running mean and returns a zero bias gradient;
    eval mode: the bias is folded into the fused conv epilogue.  Both against the plain modules in fp32."""
    from u2pl_b200 import fused
    torch.manual_seed(0)
    seq = nn.Sequential(nn.Conv2d(64, 128, k, 1, k // 2, bias=True), nn.BatchNorm2d(128), nn.ReLU(inplace=True)).cuda()
    ref = nn.Sequential(nn.Conv2d(64, 128, k, 1, k // 2, bias=True), nn.BatchNorm2d(128), nn.ReLU(inplace=True)).cuda()
    with torch.no_grad():
        seq[0].bias.normal_(0, 2.0)                                 # a bias large enough to matter if mishandled
    ref.load_state_dict(seq.state_dict())
    x = torch.randn(4, 64, 33, 37, device="cuda")
    xb = x.bfloat16().contiguous(memory_format=torch.channels_last)
    xr = xb.float().requires_grad_(True)
    xq = xb.clone().requires_grad_(True)
    seq.train(); ref.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = fused.run_sequential(seq, xq)
    yr = ref(xr)
    assert float((y.float() - yr).abs().max()) <= 3e-2 * float(yr.abs().max())
    g = torch.randn_like(yr)
    y.backward(g.bfloat16().contiguous(memory_format=torch.channels_last))
    yr.backward(g)
    assert float((seq[1].running_mean - ref[1].running_mean).abs().max()) <= 2e-2      # includes momentum * bias
    assert float((seq[1].running_var - ref[1].running_var).abs().max()) <= 2e-2 * float(ref[1].running_var.abs().max())
    assert float(seq[0].bias.grad.abs().max()) == 0.0 and float(ref[0].bias.grad.abs().max()) <= 1e-3   # identically / numerically zero
    assert float((seq[0].weight.grad - ref[0].weight.grad).norm() / ref[0].weight.grad.norm()) <= 5e-2   # bf16 operands
    seq.eval(); ref.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ye = fused.run_sequential(seq, xb)
    with torch.no_grad():
        ref[1].load_state_dict(seq[1].state_dict())
        yre = ref(xb.float())
    assert float((ye.float() - yre).abs().max()) <= 3e-2 * float(yre.abs().max())
