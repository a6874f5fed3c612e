"""GPU: the two routes into the kernels must agree.
Route A = what the UNCHANGED reference driver executes: its inline torch/numpy block (train_semi.py:401-465,
restated here line by line because it has no importable name) feeding the drop-in functions
u2pl.utils.loss_helper.compute_unsupervised_loss / compute_contra_memobank_loss / u2pl.utils.utils.label_onehot.
Route B = the fused block of u2pl_b200.step.SemiStep (one entropy pass, class bitmasks, no one-hots).
Same inputs, same RNG => same key counts and bank contents, losses within 1e-4 (the routes differ only in the
arithmetic of the entropy near the percentile cuts: ATen/numpy in A, the contract in B)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_dropin_route_equals_fused_route():
    import u2pl_b200
    u2pl_b200.install()
    from u2pl.utils.loss_helper import compute_contra_memobank_loss, compute_unsupervised_loss
    from u2pl.utils.utils import label_onehot
    from u2pl_b200 import contra, ops

    g = torch.Generator(device="cuda").manual_seed(11)
    B, C, H, W, h, w, D = 3, 21, 129, 129, 33, 33, 256
    cfg_contra = dict(negative_high_entropy=True, low_rank=3, high_rank=20, current_class_threshold=0.3,
                      current_class_negative_threshold=1, low_entropy_threshold=20, num_negatives=50, num_queries=256,
                      temperature=0.5)
    epoch, epochs = 40, 80
    low_t = torch.randn(B, C, h, w, device="cuda", generator=g) * 4
    pred_all_teacher = torch.cat((torch.randn(B, C, h, w, device="cuda", generator=g) * 4, low_t))
    pred_u_large_teacher = F.interpolate(low_t, (H, W), mode="bilinear", align_corners=True)
    pred_u_large = torch.randn(B, C, H, W, device="cuda", generator=g)
    label_l = torch.randint(0, C, (B, H, W), device="cuda", generator=g)
    label_l[:, :7] = 255
    label_u_aug = pred_u_large_teacher.argmax(1)
    label_u_aug[:, 40:60, 10:50] = 255
    rep_all = torch.randn(2 * B, D, h, w, device="cuda", generator=g)
    rep_all_teacher = torch.randn(2 * B, D, h, w, device="cuda", generator=g)
    prob_all_teacher = F.softmax(pred_all_teacher, dim=1)
    drop_percent = 100 - (100 - 80) * (1 - epoch / epochs)
    alpha_t = cfg_contra["low_entropy_threshold"] * (1 - epoch / epochs)

    def banks():
        return ([[torch.zeros(0, D)] for _ in range(C)], [torch.zeros(1, dtype=torch.long) for _ in range(C)],
                [50000] + [30000] * (C - 1))

    # ---------------- route A: driver-inline code + drop-in functions
    pa = pred_u_large.clone().requires_grad_(True)
    ra = rep_all.clone().requires_grad_(True)
    unsup_a = compute_unsupervised_loss(pa, label_u_aug.clone(), drop_percent, pred_u_large_teacher.detach())
    with torch.no_grad():                                                      # train_semi.py:401-465
        prob = torch.softmax(pred_u_large_teacher, dim=1)
        entropy = -torch.sum(prob * torch.log(prob + 1e-10), dim=1)
        low_thresh = np.percentile(entropy[label_u_aug != 255].cpu().numpy().flatten(), alpha_t)
        low_entropy_mask = entropy.le(low_thresh).float() * (label_u_aug != 255).bool()
        high_thresh = np.percentile(entropy[label_u_aug != 255].cpu().numpy().flatten(), 100 - alpha_t)
        high_entropy_mask = entropy.ge(high_thresh).float() * (label_u_aug != 255).bool()
        low_mask_all = torch.cat(((label_l.unsqueeze(1) != 255).float(), low_entropy_mask.unsqueeze(1)))
        low_mask_all = F.interpolate(low_mask_all, size=(h, w), mode="nearest")
        high_mask_all = torch.cat(((label_l.unsqueeze(1) != 255).float(), high_entropy_mask.unsqueeze(1)))
        high_mask_all = F.interpolate(high_mask_all, size=(h, w), mode="nearest")
        label_l_small = F.interpolate(label_onehot(label_l, C), size=(h, w), mode="nearest")
        label_u_small = F.interpolate(label_onehot(label_u_aug, C), size=(h, w), mode="nearest")
    mb_a, ptr_a, qs = banks()
    torch.manual_seed(5)
    keys_a, contra_a = compute_contra_memobank_loss(ra, label_l_small.long(), label_u_small.long(), prob_all_teacher[:B],
                                                    prob_all_teacher[B:], low_mask_all, high_mask_all, cfg_contra,
                                                    mb_a, ptr_a, qs, rep_all_teacher)
    # second call so that banks are non-empty and the InfoNCE branch runs
    torch.manual_seed(6)
    keys_a2, contra_a2 = compute_contra_memobank_loss(ra, label_l_small.long(), label_u_small.long(), prob_all_teacher[:B],
                                                      prob_all_teacher[B:], low_mask_all, high_mask_all, cfg_contra,
                                                      mb_a, ptr_a, qs, rep_all_teacher)
    (unsup_a + contra_a2).backward()

    # ---------------- route B: fused block of SemiStep (step.py)
    pb = pred_u_large.clone().requires_grad_(True)
    rb = rep_all.clone().requires_grad_(True)
    target = label_u_aug.clone()
    ent, thresh, _ = ops.entropy_thresholds(pred_u_large_teacher, label_u_aug, [drop_percent, alpha_t, 100 - alpha_t])
    n_kept, _ = ops.partition_target_(ent, target, thresh, 0)
    unsup_b = ops.unsup_ce(pb, target, n_kept)
    bits, low_b, high_b = ops.contra_prep_lowres(label_l, label_u_aug, ent, thresh, 1, 2, (h, w), C, True)
    mb_b, ptr_b, _ = banks()
    torch.manual_seed(5)
    keys_b, contra_b = contra.compute_contra_memobank_loss(rb, None, None, prob_all_teacher[:B], prob_all_teacher[B:], low_b,
                                                           high_b, cfg_contra, mb_b, ptr_b, qs, rep_all_teacher, label_bits=bits)
    torch.manual_seed(6)
    keys_b2, contra_b2 = contra.compute_contra_memobank_loss(rb, None, None, prob_all_teacher[:B], prob_all_teacher[B:], low_b,
                                                             high_b, cfg_contra, mb_b, ptr_b, qs, rep_all_teacher, label_bits=bits)
    (unsup_b + contra_b2).backward()

    assert (low_b != low_mask_all).sum().item() <= 2 and (high_b != high_mask_all).sum().item() <= 2   # percentile tie band only
    same_masks = torch.equal(low_b, low_mask_all) and torch.equal(high_b, high_mask_all)
    if same_masks:
        assert keys_a == keys_b and keys_a2 == keys_b2
        for c in range(C):
            a = contra.bank_for(mb_a, qs, D, "cuda").materialize(c)
            b = contra.bank_for(mb_b, qs, D, "cuda").materialize(c)
            assert torch.equal(a, b)
        assert abs(contra_a2.item() - contra_b2.item()) <= 1e-5 and contra_a2.item() > 0
        assert (ra.grad - rb.grad).abs().max().item() <= 1e-6
    assert abs(unsup_a.item() - unsup_b.item()) <= 1e-4 * abs(unsup_a.item())
    assert (pa.grad - pb.grad).abs().max().item() <= 1e-6
    contra.forget_banks()
