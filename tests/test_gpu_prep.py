"""GPU parity: fused low-res prep (train_semi.py:408-465) and label_onehot (utils.py:50-59, quirk Q8)
against the oracle restatement and the fixtures produced by running the reference driver lines."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dtype) if dtype is not None else t).cuda()


def _bits_from_onehot(oh):
    C = oh.shape[1]
    return (oh.astype(np.uint32) * (1 << np.arange(C, dtype=np.uint32))[None, :, None, None]).sum(1).astype(np.uint32)


@pytest.mark.parametrize("name", ["prep_c21", "prep_c19_cutout"])
def test_prep_lowres_golden(golden, name):
    from u2pl_b200 import ops
    g = golden(name)
    C, h, w = int(g["C"]), int(g["h"]), int(g["w"])
    label_l, label_u = g["label_l"].astype(np.int64), g["label_u_aug"].astype(np.int64)
    alpha_t = float(g["alpha_t"])
    ent, thresh, _ = ops.entropy_thresholds(_dev(g["pred_u_large_teacher"]), _dev(label_u), [alpha_t, 100 - alpha_t])
    bits, low, high = ops.contra_prep_lowres(_dev(label_l), _dev(label_u), ent, thresh, 0, 1, (h, w), C)
    out = port.contra_prep(g["pred_u_large_teacher"], label_l, label_u, alpha_t, C, (h, w))
    assert np.array_equal(low.cpu().numpy(), out["low_mask_all"])            # bit-exact vs oracle
    assert np.array_equal(high.cpu().numpy(), out["high_mask_all"])
    want_bits = np.concatenate([_bits_from_onehot(out["label_l_small"]), _bits_from_onehot(out["label_u_small"])])
    assert np.array_equal(bits.cpu().numpy().view(np.uint32).reshape(want_bits.shape), want_bits)
    # the reference's own labels (exact) and masks (exact away from the percentile tie band)
    ref_bits = np.concatenate([_bits_from_onehot(g["label_l_small"]), _bits_from_onehot(g["label_u_small"])])
    assert np.array_equal(want_bits, ref_bits)
    assert (low.cpu().numpy() != g["low_mask_all"]).sum() <= 2 and (high.cpu().numpy() != g["high_mask_all"]).sum() <= 2


def test_label_onehot_quirk(golden):
    from u2pl_b200 import ops
    rng = np.random.default_rng(0)
    lab = rng.integers(0, 21, (3, 17, 19))
    lab[rng.random(lab.shape) < 0.2] = 255
    got = ops.label_onehot(_dev(lab), 21).cpu().numpy()
    assert np.array_equal(got, port.label_onehot(lab, 21))
    assert got[1:].sum() == 0 and got[0].sum(0).max() > 1                    # multi-hot slot 0, empty others


def test_prep_lowres_full_size_consistency():
    """V16 size: the fused kernel equals gather-by-index of the full-res masks computed separately."""
    from u2pl_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    B, C, H, W, h, w = 4, 21, 513, 513, 129, 129
    x = torch.nn.functional.interpolate(torch.randn(B, C, h, w, device="cuda", generator=g) * 3, (H, W),
                                        mode="bilinear", align_corners=True).contiguous()
    label_u = x.argmax(1)
    label_u[:, 100:200, 50:300] = 255
    label_l = torch.randint(0, C, (B, H, W), device="cuda", generator=g)
    label_l[:, :10] = 255
    ent, thresh, _ = ops.entropy_thresholds(x, label_u, [10.0, 90.0])
    bits, low, high = ops.contra_prep_lowres(label_l, label_u, ent, thresh, 0, 1, (h, w), C)
    sy = torch.from_numpy(port.nearest_src_index(h, H)).cuda()
    sx = torch.from_numpy(port.nearest_src_index(w, W)).cuda()
    th = thresh.cpu().numpy()
    valid = label_u != 255
    low_full = ((ent <= float(th[0])) & valid).float()[:, sy][:, :, sx]
    high_full = ((ent >= float(th[1])) & valid).float()[:, sy][:, :, sx]
    assert torch.equal(low[B:, 0], low_full) and torch.equal(high[B:, 0], high_full)
    assert torch.equal(low[:B, 0], (label_l != 255).float()[:, sy][:, :, sx])
    ref = torch.nn.functional.interpolate(ops.label_onehot(label_l, C), size=(h, w), mode="nearest")
    want = (ref.long() << torch.arange(C, device="cuda")[None, :, None, None]).sum(1).reshape(-1)
    assert torch.equal(bits[: B * h * w].long() & 0xFFFFFFFF, want)
