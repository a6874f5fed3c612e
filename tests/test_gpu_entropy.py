"""GPU parity: fused entropy / percentile / partition / masked-CE kernels (through the C ABI)
against the oracle.  Entropies, thresholds and index sets must be BIT-EXACT (arithmetic
contract); the CE loss and gradient are floating point: |diff| <= 1e-5 relative (north star: 1e-4).
"""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu


def _ops():
    from u2pl_b200 import ops
    return ops


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _check_entropy_map(e, e_or, th, valid, exact_map):
    """exact_map: every value bit-equal to the oracle.  Two-level path: values within 2.5e-5 of the contract
    (quarter of the kernel's guard kDelta = 1e-4), bit-equal inside the guard band of every threshold, and every
    comparison against every threshold identical to the oracle's."""
    if exact_map:
        assert np.array_equal(e.view(np.uint32), e_or.view(np.uint32))
        return
    assert np.abs(e - e_or).max() <= 2.5e-5
    for t in th[np.isfinite(th)]:
        near = np.abs(e_or - t) <= 1e-4
        assert np.array_equal(e[near & valid].view(np.uint32), e_or[near & valid].view(np.uint32))
        assert np.array_equal((e >= t) & valid, (e_or >= t) & valid) and np.array_equal((e <= t) & valid, (e_or <= t) & valid)


def _rand_case(rng, B, C, H, W, frac_ignore=0.0, scale=3.0):
    h, w = max(2, H // 4), max(2, W // 4)
    x = torch.from_numpy(rng.standard_normal((B, C, h, w)).astype(np.float32)) * scale
    x = torch.nn.functional.interpolate(x, (H, W), mode="bilinear", align_corners=True).numpy()
    target = x.argmax(1).astype(np.int64)
    if frac_ignore > 0:
        target[rng.random(target.shape) < frac_ignore] = 255
    return np.ascontiguousarray(x), target


@pytest.mark.parametrize("name", ["unsup_c21", "unsup_c19_ignore", "unsup_c5_p100"])
def test_unsup_loss_golden(golden, name):
    ops = _ops()
    g = golden(name)
    predict = _dev(g["predict"]).requires_grad_(True)
    target = _dev(g["target_in"], torch.int64)
    loss = ops.unsup_loss(predict, target, float(g["percent"]), _dev(g["pred_teacher"]))
    loss.backward()
    # oracle on the same inputs
    t_or = g["target_in"].astype(np.int64)
    out = port.compute_unsupervised_loss(g["predict"], t_or, float(g["percent"]), g["pred_teacher"])
    assert np.array_equal(target.cpu().numpy(), t_or)                       # bit-exact index set
    assert abs(loss.item() - float(out["loss"])) <= 1e-5 * abs(float(out["loss"]))
    assert np.abs(predict.grad.cpu().numpy() - port.unsup_grad(g["predict"], t_or)).max() <= 1e-6
    # and the reference's own loss value on the same inputs (fixture made by running the reference)
    if np.array_equal(t_or, g["target_out"].astype(np.int64)):
        assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))


@pytest.mark.parametrize("B,C,H,W,ign,percents", [
    (2, 21, 33, 37, 0.0, [90.0]),
    (3, 19, 29, 31, 0.3, [80.0, 12.5, 87.5]),
    (1, 7, 40, 41, 0.1, [100.0, 0.0, 50.0, 99.99]),       # any-C path
    (2, 2, 17, 19, 0.0, [33.3]),
    (1, 21, 5, 3, 0.0, [20.0, 80.0]),                     # tiny, ragged
    (2, 21, 64, 64, 0.95, [95.0]),                        # almost everything ignored
])
@pytest.mark.parametrize("exact_map", [True, False])
def test_entropy_thresholds_bit_exact(B, C, H, W, ign, percents, exact_map):
    ops = _ops()
    rng = np.random.default_rng(B * 1000 + C * 10 + H)
    x, target = _rand_case(rng, B, C, H, W, ign)
    ent, thresh, n_valid = ops.entropy_thresholds(_dev(x), _dev(target), percents, exact_map=exact_map)
    ent_or = port.entropy(x)
    _check_entropy_map(ent.cpu().numpy(), ent_or, thresh.cpu().numpy(), target != 255, exact_map)
    valid = target != 255
    assert n_valid.item() == int(valid.sum())
    th = thresh.cpu().numpy()
    for j, q in enumerate(percents):
        want = port.percentile(ent_or[valid], q)
        assert th[j].view(np.uint32) == np.float32(want).view(np.uint32), (q, th[j], want)
        assert th[j] == np.percentile(ent_or[valid], q)                     # numpy itself agrees


def _fused_vs_two_step(ops, x, target, percents, part_idx):
    """u2pl_entropy_partition_fused (one cooperative launch) must equal thresholds_fast + partition_target, bit for bit."""
    ent, thresh, n_valid = ops.entropy_thresholds(x, target, percents)
    t2 = target.clone()
    n_kept, mask = ops.partition_target_(ent, t2, thresh, part_idx, want_mask=True)
    t_in = target.clone()
    ent_f, thresh_f, n_valid_f, t_f, n_kept_f, mask_f = ops.entropy_partition(x, t_in, percents, part_idx, want_mask=True)
    assert torch.equal(t_in, target)                                            # input untouched
    assert torch.equal(thresh.view(torch.int32), thresh_f.view(torch.int32))    # NaN-safe bit comparison
    assert n_valid.item() == n_valid_f.item() and n_kept.item() == n_kept_f.item()
    assert torch.equal(t2, t_f) and torch.equal(mask, mask_f)
    # The entropy map is an internal quantity, exact (contract arithmetic) only inside each path's own candidate bands
    # around the thresholds and hardware-approximate elsewhere; the bands of the two paths differ (22-bit
    # bins vs 16-bit fine bins).  What both maps must agree on bit for bit is every comparison against every threshold.
    assert (ent - ent_f).abs().max().item() <= 2.5e-5
    for j in range(thresh.numel()):
        if not torch.isnan(thresh[j]):
            assert torch.equal(ent >= thresh[j], ent_f >= thresh[j]) and torch.equal(ent <= thresh[j], ent_f <= thresh[j])
    return ent_f, thresh_f, t_f, n_kept_f


@pytest.mark.parametrize("B,C,H,W,ign,percents,part", [
    (2, 21, 33, 37, 0.0, [90.0], 0),
    (3, 19, 29, 31, 0.3, [80.0, 12.5, 87.5], 0),
    (1, 21, 5, 3, 0.0, [20.0, 80.0], 1),                  # tiny, ragged: one CTA, slice of 16 pixels
    (2, 21, 64, 64, 0.95, [95.0], 0),                     # almost everything ignored
    (4, 21, 129, 131, 0.1, [90.0, 10.0, 90.0, 100.0], 2), # four percentiles, several CTAs, repeated percentile
    (1, 7, 40, 41, 0.1, [50.0], 0),                       # class count without a specialisation: multi-launch path inside
])
def test_fused_chain_equals_two_step(B, C, H, W, ign, percents, part):
    ops = _ops()
    rng = np.random.default_rng(B * 1000 + C * 10 + H)
    x, target = _rand_case(rng, B, C, H, W, ign)
    ent, thresh, t_f, n_kept = _fused_vs_two_step(ops, _dev(x), _dev(target), percents, part)
    ent_or = port.entropy(x)                                                    # and against the oracle
    valid = target != 255
    th = thresh.cpu().numpy()
    for j, q in enumerate(percents):
        assert th[j].view(np.uint32) == np.float32(port.percentile(ent_or[valid], q)).view(np.uint32)
    want = target.copy()
    want[(ent_or >= th[part]) & valid] = 255
    assert np.array_equal(t_f.cpu().numpy(), want)                              # bit-exact reliable / unreliable index set
    assert n_kept.item() == int((want != 255).sum())


def test_fused_chain_edge_cases():
    ops = _ops()
    rng = np.random.default_rng(3)
    x, target = _rand_case(rng, 1, 21, 16, 16)
    target[:] = 255                                                             # empty population -> NaN thresholds, nothing dropped
    _, thresh, t_f, n_kept = _fused_vs_two_step(ops, _dev(x), _dev(target), [80.0], 0)
    assert np.isnan(thresh.cpu().numpy()[0]) and n_kept.item() == 0 and bool((t_f == 255).all())
    x = np.zeros((2, 21, 24, 24), np.float32)                                   # every pixel ties: everything is a candidate
    x[1, 3] = 5.0
    target = np.zeros((2, 24, 24), np.int64)
    _fused_vs_two_step(ops, _dev(x), _dev(target), [50.0, 99.0], 0)


def test_all_ignored_gives_nan():
    ops = _ops()
    rng = np.random.default_rng(0)
    x, target = _rand_case(rng, 1, 21, 16, 16)
    target[:] = 255
    for exact_map in (True, False):
        ent, thresh, n_valid = ops.entropy_thresholds(_dev(x), _dev(target), [80.0], exact_map=exact_map)
        assert n_valid.item() == 0 and np.isnan(thresh.cpu().numpy()[0])
        assert np.abs(ent.cpu().numpy() - port.entropy(x)).max() <= (0 if exact_map else 2.5e-5)


def test_heavy_ties_and_constant_logits():
    ops = _ops()
    x = np.zeros((2, 21, 24, 24), np.float32)                       # every pixel has the same entropy
    x[1, 3] = 5.0
    target = np.zeros((2, 24, 24), np.int64)
    ent, thresh, _ = ops.entropy_thresholds(_dev(x), _dev(target), [50.0, 99.0])       # two-level path, all pixels tie
    ent_or = port.entropy(x)
    assert np.array_equal(ent.cpu().numpy(), ent_or)                                   # everything is a candidate -> exact
    assert thresh.cpu().numpy()[0] == port.percentile(ent_or, 50.0)
    assert thresh.cpu().numpy()[1] == port.percentile(ent_or, 99.0)
    t = _dev(target)
    n_kept, mask = ops.partition_target_(ent, t, thresh, 0, want_mask=True)
    drop = ent_or >= port.percentile(ent_or, 50.0)
    assert np.array_equal(mask.cpu().numpy().astype(bool), drop)
    assert n_kept.item() == int((~drop).sum())


def test_entropy_masks_and_gather():
    ops = _ops()
    rng = np.random.default_rng(7)
    x, target = _rand_case(rng, 2, 19, 41, 41, 0.2)
    ent, thresh, _ = ops.entropy_thresholds(_dev(x), _dev(target), [18.7, 81.3])
    sy, sx = port.nearest_src_index(11, 41), port.nearest_src_index(11, 41)
    flat = (np.arange(2)[:, None, None] * 41 * 41 + sy[None, :, None] * 41 + sx[None, None, :]).astype(np.int64)
    low, high = ops.entropy_masks(ent, _dev(target), thresh, 0, 1, idx=_dev(flat))
    ent_or = port.entropy(x)
    valid = target != 255
    lo_t, hi_t = port.percentile(ent_or[valid], 18.7), port.percentile(ent_or[valid], 81.3)
    want_low = ((ent_or <= lo_t) & valid).reshape(-1)[flat]
    want_high = ((ent_or >= hi_t) & valid).reshape(-1)[flat]
    assert np.array_equal(low.cpu().numpy() > 0, want_low)
    assert np.array_equal(high.cpu().numpy() > 0, want_high)


@pytest.mark.parametrize("C", [21, 19, 6])
def test_cross_entropy_mean_vs_torch(C):
    ops = _ops()
    rng = np.random.default_rng(C)
    x, target = _rand_case(rng, 2, C, 37, 29, 0.25, scale=2.0)
    a = _dev(x).requires_grad_(True)
    b = _dev(x).requires_grad_(True)
    mine = ops.cross_entropy_mean(a, _dev(target))
    ref = torch.nn.functional.cross_entropy(b, _dev(target), ignore_index=255)      # torch fp32 reference
    (mine * 1.7).backward()
    (ref * 1.7).backward()
    assert abs(mine.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert (a.grad - b.grad).abs().max().item() <= 1e-7
    assert abs(mine.item() - float(port.criterion_ce(x, target))) <= 1e-5 * abs(ref.item())


def test_full_size_v16_properties():
    """BASELINE config 2 size (16 x 21 x 513 x 513): size-independent checks.
    (1) np.percentile of the GPU entropies == the on-device threshold, bitwise;
    (2) kept count == #(entropy < thresh); (3) a 50k-pixel sample of entropies is bit-equal to the oracle."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(1234)
    B, C, H, W = 16, 21, 513, 513
    low = torch.randn(B, C, 129, 129, device="cuda", generator=g) * 3
    low = torch.nn.functional.avg_pool2d(low, 5, 1, 2, count_include_pad=False)
    x = torch.nn.functional.interpolate(low, (H, W), mode="bilinear", align_corners=True).contiguous()
    target = x.argmax(1)
    percents = [90.0, 10.0, 90.0000001, 100.0]
    ent, thresh, n_valid = ops.entropy_thresholds(x, target, percents, exact_map=True)
    ent_f, thresh_f, _ = ops.entropy_thresholds(x, target, percents)                   # two-level path
    assert torch.equal(thresh, thresh_f)                                               # bit-identical thresholds
    for j in range(len(percents)):
        assert torch.equal(ent >= thresh[j], ent_f >= thresh[j]) and torch.equal(ent <= thresh[j], ent_f <= thresh[j])
    assert (ent - ent_f).abs().max().item() <= 2.5e-5
    e = ent.cpu().numpy().ravel()
    th = thresh.cpu().numpy()
    for j, q in enumerate(percents):
        assert th[j] == np.percentile(e, q), q
    t2 = target.clone()
    n_kept, _ = ops.partition_target_(ent, t2, thresh, 0)
    assert n_kept.item() == int((e < th[0]).sum()) == int((t2 != 255).sum().item())
    assert abs(n_kept.item() / e.size - 0.9) < 1e-3
    idx = np.random.default_rng(0).choice(B * H * W, 50000, replace=False)
    xs = x.permute(0, 2, 3, 1).reshape(-1, C)[torch.from_numpy(idx).cuda()].cpu().numpy()      # [50000, C]
    ent_s = port.entropy(np.ascontiguousarray(xs.T[None]))[0]                                  # B=1, HW=50000
    assert np.array_equal(ent_s.view(np.uint32), e[idx].view(np.uint32))
    # the single cooperative launch at the benchmark size: bit-identical to the multi-launch chain
    _fused_vs_two_step(ops, x, target, percents, 0)
    _fused_vs_two_step(ops, x, target, [90.0, 10.0, 90.0], 0)                          # the step's own three percentiles


@pytest.mark.parametrize("name", ["ohem_c19", "ohem_c19_kth"])
def test_ohem_golden(golden, name):
    """OHEM (loss_helper.py:502-531): kept set bit-exact vs the oracle, loss/grad vs oracle and reference."""
    ops = _ops()
    g = golden(name)
    pred_np, target_np = g["pred"], g["target"].astype(np.int64)
    pred = _dev(pred_np).requires_grad_(True)
    loss = ops.ohem_cross_entropy(pred, _dev(target_np), float(g["thresh"]), int(g["min_kept"]))
    loss.backward()
    want, kept = port.ohem_ce(pred_np, target_np, float(g["thresh"]), int(g["min_kept"]))
    new_target, kth, n_valid = ops.ohem_select(_dev(pred_np), _dev(target_np), float(g["thresh"]), int(g["min_kept"]))
    assert np.array_equal(new_target.cpu().numpy() != 255, kept)                # bit-exact kept set
    assert n_valid.item() == int((target_np != 255).sum())
    assert abs(loss.item() - float(want)) <= 1e-5 and abs(loss.item() - float(g["loss"])) <= 1e-4
    assert np.abs(pred.grad.cpu().numpy() - g["grad"]).max() <= 1e-6


@pytest.mark.parametrize("C,min_kept", [(7, 50), (21, 100000), (19, 0), (21, 1)])
def test_ohem_edge_cases(C, min_kept):
    ops = _ops()
    rng = np.random.default_rng(C + min_kept)
    x, target = _rand_case(rng, 2, C, 31, 33, 0.3, scale=1.5)
    new_target, _, _ = ops.ohem_select(_dev(x), _dev(target), 0.7, min_kept)
    _, kept = port.ohem_ce(x, target, 0.7, min_kept)
    assert np.array_equal(new_target.cpu().numpy() != 255, kept)


def test_v16_index_set_flips_vs_torch_cuda_eager():
    """north_star: "bit-exact for the integer reliable/unreliable index sets".  Exactness is by construction against the
    arithmetic contract (oracle); this test QUANTIFIES the distance to the reference's own CUDA arithmetic at BASELINE
    config-2 size: loss_helper.py:35-43 / train_semi.py:402-418 executed with torch-CUDA ops + np.percentile on the same
    16 x 21 x 513 x 513 logits, and the number of pixels (of 4 210 704, for each of the three cuts of a step) whose side
    of the cut differs.  Two libm's agree to ~1e-7, so only pixels inside that band of a threshold can flip."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(4321)
    B, C, H, W = 16, 21, 513, 513
    low = torch.randn(B, C, 129, 129, device="cuda", generator=g) * 3
    low = torch.nn.functional.avg_pool2d(low, 5, 1, 2, count_include_pad=False)
    x = torch.nn.functional.interpolate(low, (H, W), mode="bilinear", align_corners=True).contiguous()
    target = x.argmax(1)
    percents = [90.0, 10.0, 90.0]                                     # drop_percent, alpha_t, 100 - alpha_t at epoch 40/80
    ent, thresh, _, new_t, n_kept, _ = ops.entropy_partition(x, target, percents, 0)
    with torch.no_grad():                                             # the reference's arithmetic, on this GPU
        prob = torch.softmax(x, dim=1)
        ent_ref = -torch.sum(prob * torch.log(prob + 1e-10), dim=1)
    e = ent_ref.cpu().numpy().ravel()
    flips = []
    for j, q in enumerate(percents):
        th_ref = np.percentile(e, q)
        ref_side = ent_ref >= float(th_ref) if j != 1 else ent_ref <= float(th_ref)
        our_side = ent >= thresh[j] if j != 1 else ent <= thresh[j]
        flips.append(int((ref_side != our_side).sum().item()))
        assert abs(float(thresh[j]) - float(th_ref)) <= 2e-6
    print(f"\n[flip count vs torch-CUDA eager, {B * H * W} pixels] drop-percent cut: {flips[0]}, low-entropy cut: {flips[1]}, "
          f"high-entropy cut: {flips[2]}; max |entropy - torch| near the cuts <= 2e-6")
    assert max(flips) <= 64                                           # a handful of tie-band pixels out of 4.2 M
    ref_target = target.clone()
    ref_target[ent_ref >= float(np.percentile(e, percents[0]))] = 255
    assert int((ref_target != new_t).sum().item()) == flips[0]


def test_full_size_c2_properties():
    """BASELINE config 3 per-GPU size (2 x 19 x 769 x 769, OHEM thresh 0.7 / min_kept 100000, contrastive prep to 193 x 193):
    size-independent checks at full size.  (1) one-launch chain == multi-launch path (thresholds bitwise, target / mask /
    counts equal, every threshold comparison equal) and np.percentile of the entropies == the on-device thresholds;
    (2) OHEM: the kept set is {valid, prob_of_target <= max(thresh, k-th smallest)} and holds >= min_kept pixels;
    (3) low-resolution masks == nearest-neighbour sampling of the full-resolution comparisons (train_semi.py:417-437)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(769)
    B, C, H, W, h, w = 2, 19, 769, 769, 193, 193
    low = torch.randn(B, C, h, w, device="cuda", generator=g) * 3
    low = torch.nn.functional.avg_pool2d(low, 5, 1, 2, count_include_pad=False)
    x = torch.nn.functional.interpolate(low, (H, W), mode="bilinear", align_corners=True).contiguous()
    target = x.argmax(1)
    target[:, :10] = 255                                               # a border of ignored pixels, as in the synthetic labels
    percents = [84.0, 16.0, 84.0]                                      # drop_percent / alpha_t / 100 - alpha_t at epoch 40 of 200
    ent, thresh, t_f, n_kept = _fused_vs_two_step(ops, x, target, percents, 0)
    e = ent.cpu().numpy().ravel()
    valid = (target != 255).cpu().numpy().ravel()
    th = thresh.cpu().numpy()
    for j, q in enumerate(percents):
        assert th[j] == np.percentile(e[valid], q), q
    assert n_kept.item() == int(((e < th[0]) & valid).sum()) == int((t_f != 255).sum().item())
    # (2) OHEM at this size
    pred = torch.randn(B, C, H, W, device="cuda", generator=g)
    new_target, kth, n_valid = ops.ohem_select(pred, target, 0.7, 100000)
    prob = torch.softmax(pred, 1).gather(1, target.clamp(max=C - 1).unsqueeze(1)).squeeze(1)
    kept = new_target != 255
    cut = max(0.7, float(kth))
    band = (prob - cut).abs() <= 2e-6                                  # torch's softmax vs the contract arithmetic near the cut
    assert int(kept.sum()) >= min(100000, int(n_valid))
    assert bool(((kept == ((prob <= cut) & (target != 255))) | band).all())
    # (3) contrastive prep against nearest sampling of the full-resolution comparisons
    label_l = target.clone()
    bits, low_m, high_m = ops.contra_prep_lowres(label_l, target, ent, thresh, 1, 2, (h, w), C, True)
    iy = torch.clamp((torch.arange(h, device="cuda").float() * np.float32(H / h)).floor().long(), max=H - 1)
    ix = torch.clamp((torch.arange(w, device="cuda").float() * np.float32(W / w)).floor().long(), max=W - 1)
    samp = lambda t: t[:, iy][:, :, ix]                                # noqa: E731  F.interpolate(mode="nearest") source indices
    want_low = ((ent <= thresh[1]) & (target != 255)).float()
    want_high = ((ent >= thresh[2]) & (target != 255)).float()
    assert torch.equal(low_m[B:, 0], samp(want_low)) and torch.equal(high_m[B:, 0], samp(want_high))
    assert torch.equal(low_m[:B, 0], samp((label_l != 255).float()))
