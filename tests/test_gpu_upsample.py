"""GPU parity: the x4 bilinear up-sampling fused with its consumers (csrc/upsample_ce.cu) against the ATen ops they
replace in train_semi.py:317-324,344-358 (F.interpolate align_corners=True -> softmax/max, -> cross_entropy), all in fp32
on the same inputs.  Floating point: losses 1e-5 relative, gradients 1e-6 absolute (north star: 1e-4), labels exact except
at probability ties."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from u2pl_b200 import ops
    return ops


def _case(seed, B, C, h, w, H, W, frac_ignore=0.2, scale=3.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    low = torch.randn(B, C, h, w, device="cuda", generator=g) * scale
    target = torch.randint(0, C, (B, H, W), device="cuda", generator=g)
    target[torch.rand(B, H, W, device="cuda", generator=g) < frac_ignore] = 255
    return low, target


SHAPES = [(2, 21, 17, 17, 65, 65), (3, 19, 25, 25, 97, 97), (1, 21, 9, 13, 33, 49), (2, 21, 5, 5, 5, 5), (1, 19, 3, 4, 12, 13)]


@pytest.mark.parametrize("B,C,h,w,H,W", SHAPES)
def test_up_softmax_max_matches_torch(B, C, h, w, H, W):
    ops = _ops()
    low, _ = _case(B * 7 + C, B, C, h, w, H, W)
    prob, label = ops.up_softmax_max(low, (H, W))
    up = F.interpolate(low, (H, W), mode="bilinear", align_corners=True)
    p_ref, l_ref = torch.max(F.softmax(up, dim=1), dim=1)
    assert (prob - p_ref).abs().max().item() <= 2e-6
    sp = F.softmax(up, dim=1)
    top2 = sp.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-6                       # arg-max may only differ where two classes tie
    assert torch.equal(label[clear], l_ref[clear])
    assert label.dtype == torch.int64 and prob.dtype == torch.float32


@pytest.mark.parametrize("B,C,h,w,H,W", SHAPES)
@pytest.mark.parametrize("mode", ["mean", "unsup"])
def test_upsampled_ce_forward_backward_match_torch(B, C, h, w, H, W, mode):
    ops = _ops()
    low, target = _case(B * 11 + C + h, B, C, h, w, H, W)
    a = low.clone().requires_grad_(True)
    b = low.clone().requires_grad_(True)
    up = F.interpolate(b, (H, W), mode="bilinear", align_corners=True)
    if mode == "mean":
        mine = ops.upsampled_ce_mean(a, target)
        ref = F.cross_entropy(up, target, ignore_index=255)
    else:
        n_kept = (target != 255).sum()
        mine = ops.upsampled_unsup_ce(a, target, n_kept)
        ref = (target.numel() / n_kept.float()) * F.cross_entropy(up, target, ignore_index=255)     # loss_helper.py:44-46
    (mine * 1.3).backward()
    (ref * 1.3).backward()
    assert abs(mine.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert (a.grad - b.grad).abs().max().item() <= 1e-6 * max(1.0, b.grad.abs().max().item() * 10)


def test_upsampled_ce_all_ignored_and_bf16_input():
    ops = _ops()
    low, target = _case(5, 2, 21, 9, 9, 33, 33)
    a = low.bfloat16().requires_grad_(True)                         # the network hands over bf16 channels-last logits
    b = low.bfloat16().float().requires_grad_(True)
    mine = ops.upsampled_ce_mean(a.contiguous(memory_format=torch.channels_last), target)
    ref = F.cross_entropy(F.interpolate(b, (33, 33), mode="bilinear", align_corners=True), target, ignore_index=255)
    mine.backward()
    ref.backward()
    assert abs(mine.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert a.grad.dtype == torch.bfloat16 and (a.grad.float() - b.grad).abs().max().item() <= 1e-2 * b.grad.abs().max().item()
    target[:] = 255
    z = ops.upsampled_ce_mean(low.clone().requires_grad_(True), target)
    assert torch.isnan(z)                                           # 0 / 0 like nn.CrossEntropyLoss on an all-ignored target


def test_full_size_v16_properties():
    """BASELINE config 2 size (16 x 21 x 129^2 -> 513^2): fused results against the unfused kernels of this library on the
    materialised up-sampled tensor (which the GPU parity tests pin to the oracle)."""
    ops = _ops()
    low, target = _case(1234, 16, 21, 129, 129, 513, 513, frac_ignore=0.1)
    up = F.interpolate(low, (513, 513), mode="bilinear", align_corners=True)
    a = low.clone().requires_grad_(True)
    mine = ops.upsampled_ce_mean(a, target)
    u = up.clone().requires_grad_(True)
    ref = ops.cross_entropy_mean(u, target)
    mine.backward()
    ref.backward()
    g_ref = torch.autograd.grad(F.interpolate(a, (513, 513), mode="bilinear", align_corners=True), a, u.grad)[0]
    assert abs(mine.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert (a.grad - g_ref).abs().max().item() <= 1e-6 * max(1.0, 10 * g_ref.abs().max().item())
    prob, label = ops.up_softmax_max(low, (513, 513))
    p_ref, l_ref = torch.max(F.softmax(up, dim=1), dim=1)
    assert (prob - p_ref).abs().max().item() <= 2e-6
    assert (label != l_ref).float().mean().item() <= 1e-5           # ties only
