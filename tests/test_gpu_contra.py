"""GPU parity: contrastive memory-bank loss through the C ABI against the oracle (and the reference's
own outputs stored in the fixtures).  Index sets / bank contents / key counts: exact.  Loss: 1e-5
absolute on a loss of O(1) (north star 1e-4).  Gradient rows: 1e-6."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu


def _cfg(g):
    cfg = {k: v for k, v in zip(g["cfg_keys"].tolist(), g["cfg_vals"].tolist())}
    for k in ("low_rank", "high_rank", "num_negatives", "num_queries"):
        cfg[k] = int(cfg[k])
    return cfg


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("name", ["contra_c21", "contra_c19_missing", "contra_c21_driver_onehot"])
def test_contra_loss_golden(golden, name, channels_last):
    from u2pl_b200 import contra
    g = golden(name)
    cfg = _cfg(g)
    C, D = g["s0_label_l"].shape[1], g["s0_rep"].shape[1]
    qsize = g["queue_size"].tolist()
    memobank = [[torch.zeros(0, D)] for _ in range(C)]
    ptrs = [torch.zeros(1, dtype=torch.long) for _ in range(C)]
    o_bank = [[np.zeros((0, D), np.float32)] for _ in range(C)]
    o_ptr = [[0] for _ in range(C)]
    for s in range(int(g["steps"])):
        rep = _dev(g[f"s{s}_rep"])
        rep_t = _dev(g[f"s{s}_rep_teacher"])
        if channels_last:
            rep = rep.contiguous(memory_format=torch.channels_last)
            rep_t = rep_t.contiguous(memory_format=torch.channels_last)
        rep.requires_grad_(True)
        args = (g[f"s{s}_label_l"].astype(np.int64), g[f"s{s}_label_u"].astype(np.int64), g[f"s{s}_prob_l"],
                g[f"s{s}_prob_u"], g[f"s{s}_low_mask"].astype(np.float32), g[f"s{s}_high_mask"].astype(np.float32))
        torch.manual_seed(1000 + s)
        new_keys, loss, plan = contra.compute_contra_memobank_loss(
            rep, *[_dev(a) for a in args], cfg, memobank, ptrs, qsize, rep_t, return_plan=True)
        loss.backward()
        torch.manual_seed(1000 + s)
        out = port.compute_contra_memobank_loss(g[f"s{s}_rep"], *args, cfg, o_bank, o_ptr, qsize,
                                                g[f"s{s}_rep_teacher"], want_grad=True)
        # integer side: exact
        assert new_keys == out["new_keys"] == g[f"s{s}_new_keys"].tolist()
        assert plan["valid_classes"] == out["valid_classes"]
        tot = plan["totals"]
        sel = out["sel"]
        assert [int(x) for x in tot[0]] == [len(x) for x in sel["lowvalid"]]
        assert [int(x) for x in tot[1]] == [len(x) for x in sel["anchors"]]
        assert [int(x) for x in tot[2]] == [len(x) for x in sel["negs"]]
        bits3 = plan["bits3"].cpu().numpy().view(np.uint32)
        for c in range(C):
            assert np.array_equal(np.flatnonzero((bits3[1] >> c) & 1), sel["anchors"][c])
            assert np.array_equal(np.flatnonzero((bits3[2] >> c) & 1), sel["negs"][c])
        assert [int(p[0]) for p in ptrs] == [int(p[0]) for p in o_ptr] == g[f"s{s}_ptr"].tolist()
        assert [m[0].shape[0] for m in memobank] == g[f"s{s}_bank_len"].tolist()
        bank = contra.bank_for(memobank, qsize, D, rep.device)
        for c in range(C):
            assert np.array_equal(bank.materialize(c).cpu().numpy(), o_bank[c][0])       # FIFO content
        if plan["nact"]:
            pix = plan["anchor_pix"].cpu().numpy().reshape(plan["nact"], -1)
            for a, smp in enumerate(out["sampled"]):
                assert np.array_equal(pix[a], sel["anchors"][smp["j"]][smp["a_idx"]])     # same anchors drawn
        # floating point side
        assert abs(loss.item() - float(out["loss"])) <= 1e-5
        grad = rep.grad.cpu().numpy() if rep.grad is not None else np.zeros_like(g[f"s{s}_rep"])
        assert np.abs(grad - out["rep_grad"]).max() <= 1e-6
    for c in range(C):                                                                    # vs the reference's own bank
        assert np.array_equal(contra.bank_for(memobank, qsize, D, "cuda").materialize(c).cpu().numpy(), g[f"bank_{c}"])
    contra.forget_banks()


def test_contra_matches_reference_rng_stream(golden):
    """Same torch seed as the fixture => the reference's own loss values (its RNG stream is consumed
    identically: loss_helper.py:179-181,194-196)."""
    from u2pl_b200 import contra
    g = golden("contra_c21")
    cfg = _cfg(g)
    C, D = g["s0_label_l"].shape[1], g["s0_rep"].shape[1]
    qsize = g["queue_size"].tolist()
    memobank = [[torch.zeros(0, D)] for _ in range(C)]
    ptrs = [torch.zeros(1, dtype=torch.long) for _ in range(C)]
    torch.manual_seed(int(g["seed"]))
    for s in range(int(g["steps"])):
        rep = _dev(g[f"s{s}_rep"]).requires_grad_(True)
        new_keys, loss = contra.compute_contra_memobank_loss(
            rep, _dev(g[f"s{s}_label_l"], torch.int64), _dev(g[f"s{s}_label_u"], torch.int64), _dev(g[f"s{s}_prob_l"]),
            _dev(g[f"s{s}_prob_u"]), _dev(g[f"s{s}_low_mask"], torch.float32), _dev(g[f"s{s}_high_mask"], torch.float32),
            cfg, memobank, ptrs, qsize, _dev(g[f"s{s}_rep_teacher"]))
        loss.backward()
        assert abs(loss.item() - float(g[f"s{s}_loss"])) <= 1e-4          # north-star tolerance vs the reference
        assert np.abs(rep.grad.cpu().numpy() - g[f"s{s}_grad"]).max() <= 1e-5
    contra.forget_banks()


@pytest.mark.parametrize("channels_last", [False, True])
def test_contra_large_random_vs_oracle(channels_last):
    """Bigger, D=256, genuine one-hot labels on every image, bank wrap-around, 3 steps.  channels_last: the network's
    own layout (register-accumulating prototype kernel, strided key / anchor gathers)."""
    from u2pl_b200 import contra
    rng = np.random.default_rng(77)
    Bl = Bu = 3
    C, D, h, w = 21, 256, 33, 29
    cfg = dict(negative_high_entropy=True, low_rank=3, high_rank=20, current_class_threshold=0.3,
               current_class_negative_threshold=1, num_negatives=50, num_queries=256, temperature=0.5)
    qsize = [150] * C
    qsize[0] = 200
    memobank = [[torch.zeros(0, D)] for _ in range(C)]
    ptrs = [torch.zeros(1, dtype=torch.long) for _ in range(C)]
    o_bank = [[np.zeros((0, D), np.float32)] for _ in range(C)]
    o_ptr = [[0] for _ in range(C)]
    for s in range(3):
        logit = rng.standard_normal((Bl + Bu, C, h, w)).astype(np.float32) * 4
        prob = port.softmax(logit)
        order = np.argsort(-prob, axis=1, kind="stable")
        pick = rng.integers(0, 8, (Bl + Bu, 1, h, w))
        lab = np.where(rng.random((Bl + Bu, h, w)) < 0.5, np.take_along_axis(order, pick, 1)[:, 0], prob.argmax(1))
        lab[:Bl, :3] = 255
        onehot = (np.arange(C)[None, :, None, None] == lab[:, None]).astype(np.int64)
        low = np.concatenate([(lab[:Bl] != 255), rng.random((Bu, h, w)) < 0.6]).astype(np.float32)[:, None]
        high = np.concatenate([(lab[:Bl] != 255), rng.random((Bu, h, w)) < 0.7]).astype(np.float32)[:, None]
        rep_np = rng.standard_normal((Bl + Bu, D, h, w)).astype(np.float32)
        rept_np = rng.standard_normal((Bl + Bu, D, h, w)).astype(np.float32)
        args = (onehot[:Bl], onehot[Bl:], prob[:Bl], prob[Bl:], low, high)
        rep, rep_t = _dev(rep_np), _dev(rept_np)
        if channels_last:
            rep, rep_t = rep.contiguous(memory_format=torch.channels_last), rep_t.contiguous(memory_format=torch.channels_last)
        rep.requires_grad_(True)
        torch.manual_seed(5 + s)
        new_keys, loss, plan = contra.compute_contra_memobank_loss(
            rep, *[_dev(a) for a in args], cfg, memobank, ptrs, qsize, rep_t, return_plan=True)
        loss.backward()
        torch.manual_seed(5 + s)
        out = port.compute_contra_memobank_loss(rep_np, *args, cfg, o_bank, o_ptr, qsize, rept_np, want_grad=True)
        assert new_keys == out["new_keys"]
        bank = contra.bank_for(memobank, qsize, D, rep.device)
        for c in range(C):
            assert np.array_equal(bank.materialize(c).cpu().numpy(), o_bank[c][0])
        assert abs(loss.item() - float(out["loss"])) <= 2e-5
        assert np.abs(rep.grad.cpu().numpy() - out["rep_grad"]).max() <= 1e-6
        proto = plan["proto"].cpu().numpy()
        for c in plan["valid_classes"]:
            rows = rept_np.transpose(0, 2, 3, 1).reshape(-1, D)[out["sel"]["lowvalid"][c]]
            assert np.abs(proto[c] - rows.mean(0)).max() <= 1e-5
    assert sum(new_keys) > 0 and plan["nact"] > 0
    contra.forget_banks()
