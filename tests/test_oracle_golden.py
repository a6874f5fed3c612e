"""Pins oracle/ (the CPU restatement) against fixtures produced by RUNNING THE REFERENCE
(oracle/make_golden.py, tests/golden/*.npz).  CPU only.

Floating point outputs: within the tolerance written beside each assert.  Index sets: the
reference computes entropies through ATen's libm; the oracle through the arithmetic contract
(DESIGN.md section 3), so a pixel whose entropy lies within a few ulps of the percentile cut may
fall on the other side.  Such pixels must be (a) few and (b) inside the tie band; everything
else must match exactly.
"""
import numpy as np
import pytest
import torch

from oracle import port

TIE_BAND = 2e-6          # |entropy - threshold| below which the two libm's may disagree (~8 ulp at 3.0)


def _check_sets(mine, ref, ref_entropy, ref_thresh, what):
    diff = mine != ref
    n = int(diff.sum())
    if n:
        d = np.abs(ref_entropy[diff] - ref_thresh)
        assert d.max() <= TIE_BAND, f"{what}: {n} pixels differ outside the tie band (max {d.max()})"
    assert n <= max(2, mine.size // 2000), f"{what}: {n} tie-band disagreements"
    return n


@pytest.mark.parametrize("name", ["unsup_c21", "unsup_c19_ignore", "unsup_c5_p100"])
def test_unsup_loss_matches_reference(golden, name):
    g = golden(name)
    target = g["target_in"].astype(np.int64)
    out = port.compute_unsupervised_loss(g["predict"], target, float(g["percent"]), g["pred_teacher"])
    assert np.abs(out["entropy"] - g["ref_entropy"]).max() <= 1e-6          # fp32 entropies ~O(1)
    assert abs(float(out["thresh"]) - float(g["ref_thresh"])) <= 1e-6
    ref_drop = (g["target_out"] == 255) & (g["target_in"] != 255)
    n = _check_sets(out["drop_mask"], ref_drop, g["ref_entropy"], float(g["ref_thresh"]), name)
    if n == 0:
        assert np.array_equal(target, g["target_out"].astype(np.int64))
        assert abs(float(out["loss"]) - float(g["loss"])) <= 1e-4 * max(1.0, abs(float(g["loss"])))
        grad = port.unsup_grad(g["predict"], target)
        assert np.abs(grad - g["grad"]).max() <= 1e-6


def test_percentile_matches_numpy():
    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(1, 5000))
        a = (rng.random(n) * 3).astype(np.float32)
        if rng.random() < 0.2:
            a = np.round(a, 1)                                  # heavy ties
        q = float(rng.choice([0, 100, 80, 90, 20, 10, 99.99, 12.5, rng.random() * 100]))
        assert port.percentile(a, q) == np.percentile(a, q), (n, q)


@pytest.mark.parametrize("name", ["prep_c21", "prep_c19_cutout"])
def test_contra_prep_matches_reference(golden, name):
    g = golden(name)
    out = port.contra_prep(g["pred_u_large_teacher"], g["label_l"].astype(np.int64), g["label_u_aug"].astype(np.int64),
                           float(g["alpha_t"]), int(g["C"]), (int(g["h"]), int(g["w"])))
    assert abs(float(out["low_thresh"]) - float(g["low_thresh"])) <= 1e-6
    assert abs(float(out["high_thresh"]) - float(g["high_thresh"])) <= 1e-6
    assert np.array_equal(out["label_l_small"], g["label_l_small"].astype(np.float32))
    assert np.array_equal(out["label_u_small"], g["label_u_small"].astype(np.float32))
    H, W = g["label_l"].shape[1:]
    sy, sx = port.nearest_src_index(int(g["h"]), H), port.nearest_src_index(int(g["w"]), W)
    ent_small = g["ref_entropy"][:, sy[:, None], sx[None, :]]
    B = g["label_l"].shape[0]
    _check_sets(out["low_mask_all"][B:, 0], g["low_mask_all"][B:, 0].astype(np.float32), ent_small,
                float(g["low_thresh"]), name + " low")
    _check_sets(out["high_mask_all"][B:, 0], g["high_mask_all"][B:, 0].astype(np.float32), ent_small,
                float(g["high_thresh"]), name + " high")
    assert np.array_equal(out["low_mask_all"][:B], g["low_mask_all"][:B].astype(np.float32))


@pytest.mark.parametrize("name", ["contra_c21", "contra_c19_missing", "contra_c21_driver_onehot"])
def test_contra_loss_matches_reference(golden, name):
    g = golden(name)
    cfg = {k: v for k, v in zip(g["cfg_keys"].tolist(), g["cfg_vals"].tolist())}
    for k in ("low_rank", "high_rank", "num_negatives", "num_queries"):
        cfg[k] = int(cfg[k])
    C = g["s0_label_l"].shape[1]
    D = g["s0_rep"].shape[1]
    memobank = [[np.zeros((0, D), np.float32)] for _ in range(C)]
    ptrs = [[0] for _ in range(C)]
    qsize = g["queue_size"].tolist()
    torch.manual_seed(int(g["seed"]))
    for s in range(int(g["steps"])):
        out = port.compute_contra_memobank_loss(
            g[f"s{s}_rep"], g[f"s{s}_label_l"].astype(np.int64), g[f"s{s}_label_u"].astype(np.int64),
            g[f"s{s}_prob_l"], g[f"s{s}_prob_u"], g[f"s{s}_low_mask"].astype(np.float32),
            g[f"s{s}_high_mask"].astype(np.float32), cfg, memobank, ptrs, qsize, g[f"s{s}_rep_teacher"], want_grad=True)
        assert out["new_keys"] == g[f"s{s}_new_keys"].tolist()
        assert [m[0].shape[0] for m in memobank] == g[f"s{s}_bank_len"].tolist()
        assert [int(p[0]) for p in ptrs] == g[f"s{s}_ptr"].tolist()
        assert abs(float(out["loss"]) - float(g[f"s{s}_loss"])) <= 1e-5            # fp32 loss ~1
        assert np.abs(out["rep_grad"] - g[f"s{s}_grad"]).max() <= 1e-6
    for c in range(C):
        assert np.array_equal(memobank[c][0], g[f"bank_{c}"])                      # FIFO content: exact copies


@pytest.mark.parametrize("name", ["ohem_c19", "ohem_c19_kth"])
def test_ohem_matches_reference(golden, name):
    g = golden(name)
    loss, _ = port.ohem_ce(g["pred"], g["target"].astype(np.int64), float(g["thresh"]), int(g["min_kept"]))
    assert abs(float(loss) - float(g["loss"])) <= 1e-5
    assert abs(float(port.criterion_ce(g["pred"], g["target"].astype(np.int64))) - float(g["ce_loss"])) <= 1e-5


@pytest.mark.parametrize("name", ["aug_cutmix", "aug_cutout"])
def test_strong_aug_matches_reference(golden, name):
    g = golden(name)
    np.random.seed(int(g["seed"]))
    nd, nt, nl = port.generate_unsup_data(g["data"], g["target"].astype(np.int64), g["logits"], mode=str(g["mode"]))
    assert np.array_equal(nt, g["new_target"].astype(np.int64))
    assert np.array_equal(nd, g["new_data"])
    assert np.array_equal(nl, g["new_logits"])


@pytest.mark.parametrize("name", ["aug_cutmix", "aug_cutout", "aug_classmix"])
def test_dropin_generate_unsup_data_matches_reference(golden, name):
    """The DROP-IN function itself (u2pl.dataset.augmentation.generate_unsup_data of the mirror package, what
    train_semi.py:331-337 calls) against fixtures made by running the reference's: CutMix, CutOut and ClassMix, same
    numpy / torch RNG streams.  (Tensor ops only, so it runs on CPU tensors here and on device tensors in the step.)"""
    import torch
    import u2pl_b200
    u2pl_b200.install()
    from u2pl.dataset.augmentation import generate_unsup_data
    g = golden(name)
    np.random.seed(int(g["seed"]))
    torch.manual_seed(int(g["seed"]))
    target_in = torch.from_numpy(g["target"].astype(np.int64))
    nd, nt, nl = generate_unsup_data(torch.from_numpy(g["data"]), target_in.clone(), torch.from_numpy(g["logits"]).clone(),
                                     mode=str(g["mode"]))
    assert nt.dtype == torch.int64
    assert np.array_equal(nt.numpy(), g["new_target"].astype(np.int64))
    assert np.array_equal(nd.numpy(), g["new_data"])
    assert np.array_equal(nl.numpy(), g["new_logits"])


def test_rank_matches_torch_sort():
    rng = np.random.default_rng(3)
    p = rng.random((2, 7, 5, 6)).astype(np.float32)
    p[0, 2] = p[0, 4]                                           # exact ties -> stable order
    idx = torch.sort(torch.from_numpy(p), dim=1, descending=True, stable=True)[1].numpy()
    rank = port.rank_desc_stable(p)
    pos = np.empty_like(idx)
    np.put_along_axis(pos, idx, np.arange(7).reshape(1, 7, 1, 1).repeat(2, 0).repeat(5, 2).repeat(6, 3), axis=1)
    assert np.array_equal(rank, pos)


def test_dequeue_newest_kept():
    q, p = [np.zeros((0, 2), np.float32)], [0]
    rows = lambda a, b: np.arange(a, b, dtype=np.float32)[:, None].repeat(2, 1)
    assert port.dequeue_and_enqueue([rows(0, 3)], q, p, 5) == 3 and p[0] == 3
    assert port.dequeue_and_enqueue([rows(3, 5), rows(5, 7)], q, p, 5) == 4 and p[0] == 5
    assert np.array_equal(q[0][:, 0], [2, 3, 4, 5, 6])
    assert port.dequeue_and_enqueue([rows(7, 20)], q, p, 5) == 13
    assert np.array_equal(q[0][:, 0], [15, 16, 17, 18, 19])
