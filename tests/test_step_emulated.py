"""CPU: the WHOLE fused training step (u2pl_b200/step.py: teacher T1, augmentation, student forward, supervised CE,
teacher T2, fused entropy / percentiles / partition, masked CE, contrastive memory-bank loss, backward, SGD, EMA) over
the emulated C ABI (tests/emulated_abi.py), against the CPU restatement of the reference driver step
(oracle/step_port.py) from identical weights, inputs and RNG seeds.  fp32 network on both sides (same conv library
here), so the three losses agree to 1e-4; the GPU counterpart is tests/test_gpu_step.py.  This is the product's host
logic end to end -- ops.py, contra.py, bank.py, step.py, the loss_helper mirror -- running where there is no GPU."""
import copy

import numpy as np
import pytest
import torch

import bench
import emulated_abi
from oracle import model_port, step_port


@pytest.fixture
def emulated(monkeypatch):
    from u2pl_b200 import contra
    fake = emulated_abi.install(monkeypatch)
    yield fake
    contra.forget_banks()


def test_semi_step_on_emulated_abi_matches_reference_step(emulated):
    import u2pl_b200
    u2pl_b200.install()
    from u2pl.models.model_helper import ModelBuilder
    from u2pl.utils.loss_helper import get_criterion
    from u2pl.utils.lr_helper import get_optimizer
    from u2pl_b200 import contra
    from u2pl_b200.step import SemiStep

    C, crop, bl, bu = 21, 65, 2, 2
    cfg = bench.make_cfg("tiny")
    cfg["net"]["encoder"]["type"] = "u2pl.models.resnet.resnet50"
    torch.manual_seed(3)
    model = ModelBuilder(copy.deepcopy(cfg["net"]))
    with torch.no_grad():
        model.decoder.classifier[-1].weight.mul_(8.0)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0
    teacher = copy.deepcopy(model)
    ref = step_port.ReferenceStep(model_port.state_from_module(model), model_port.state_from_module(teacher), cfg, "resnet50")
    ref.student.dropout_p = ref.teacher.dropout_p = 0.0
    for p in teacher.parameters():
        p.requires_grad = False
    lr = 0.001
    opt = get_optimizer([dict(params=model.encoder.parameters(), lr=lr),
                         dict(params=model.decoder.parameters(), lr=lr * 10)], cfg["trainer"]["optimizer"])
    memobank = [[torch.zeros(0, 256)] for _ in range(C)]
    ptrs = [torch.zeros(1, dtype=torch.long) for _ in range(C)]
    qsize = [30000] * C
    qsize[0] = 50000
    step = SemiStep(model, teacher, opt, get_criterion(cfg), cfg, memobank, ptrs, qsize, amp=False, channels_last=False)

    got, want = [], []
    for rnd, (seed_np, seed_t) in enumerate([(5, 6), (7, 8)]):
        image_l, label_l, image_u = bench.synth_batch(100 + rnd, bl, bu, crop, C)
        np.random.seed(seed_np)
        torch.manual_seed(seed_t)
        want.append(ref.step(image_l, label_l, image_u, 40, 4000 + rnd, 100))
        np.random.seed(seed_np)
        torch.manual_seed(seed_t)
        losses = step(image_l, label_l, image_u, 40, 4000 + rnd, 100)
        got.append([float(x) for x in losses])
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert abs(a - b) <= 1e-4 * max(1.0, abs(b)), (got, want)
    assert want[1][2] > 0 and got[1][2] > 0                      # the contrastive branch really ran in step 2
    bank = contra.bank_for(memobank, qsize, 256, torch.device("cpu"))
    for c in range(C):                                           # same keys in the same FIFO order (to fp32 noise of step-2 weights)
        a, b = bank.materialize(c).numpy(), ref.memobank[c][0]
        assert a.shape == b.shape and np.allclose(a, b, atol=1e-4), c
    k = "decoder.classifier.8.weight"
    assert (dict(teacher.named_parameters())[k].detach() - ref.teacher.s[k].detach()).abs().max() <= 1e-5     # EMA
    assert (dict(model.named_parameters())[k].detach() - ref.student.s[k].detach()).abs().max() <= 1e-4       # SGD


@pytest.mark.parametrize("min_kept", [50, 100000])
def test_ohem_criterion_with_aux_on_emulated_abi(emulated, min_kept):
    """get_criterion -> CriterionOhem (Cityscapes configs: ohem + aux head, loss_helper.py:323-360, 451-531) through
    ops.ohem_select / cross_entropy_mean, forward and gradient, against the oracle's ohem_ce."""
    import u2pl_b200
    from oracle import port
    u2pl_b200.install()
    from u2pl.utils.loss_helper import get_criterion
    crit = get_criterion({"criterion": {"type": "ohem", "kwargs": {"thresh": 0.7, "min_kept": min_kept}},
                          "net": {"aux_loss": {"loss_weight": 0.4, "aux_plane": 1024}}, "dataset": {"ignore_label": 255}})
    g = torch.Generator().manual_seed(min_kept)
    main = (torch.randn(2, 19, 13, 11, generator=g) * 3).requires_grad_(True)
    aux = (torch.randn(2, 19, 13, 11, generator=g) * 3).requires_grad_(True)
    target = torch.randint(0, 19, (2, 13, 11), generator=g)
    target[:, :2] = 255
    loss = crit([main, aux], target)
    loss.backward()
    l_main, kept_main = port.ohem_ce(main.detach().numpy(), target.numpy(), 0.7, min_kept)
    l_aux, _ = port.ohem_ce(aux.detach().numpy(), target.numpy(), 0.7, min_kept)
    assert abs(float(loss.detach()) - (float(l_main) + 0.4 * float(l_aux))) <= 1e-5
    mr = main.detach().clone().requires_grad_(True)
    t_kept = torch.where(torch.from_numpy(kept_main), target, torch.full_like(target, 255))
    torch.nn.functional.cross_entropy(mr, t_kept, ignore_index=255).backward()
    assert (main.grad - mr.grad).abs().max() <= 1e-6


def test_sup_only_branch_on_emulated_abi(emulated):
    """epoch < sup_only_epoch (train_semi.py:288-307): supervised CE on the labelled crops, teacher forward in train mode
    for its BatchNorm statistics, zero unsupervised / contrastive losses, one SGD step, no EMA -- against plain torch."""
    import u2pl_b200
    u2pl_b200.install()
    from u2pl.models.model_helper import ModelBuilder
    from u2pl.utils.loss_helper import get_criterion
    from u2pl.utils.lr_helper import get_optimizer
    from u2pl_b200.step import SemiStep
    C, crop = 21, 49
    cfg = bench.make_cfg("tiny")
    cfg["trainer"]["sup_only_epoch"] = 1
    torch.manual_seed(11)
    model = ModelBuilder(copy.deepcopy(cfg["net"]))
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0
    teacher, ref_model = copy.deepcopy(model), copy.deepcopy(model)
    t_before = copy.deepcopy(teacher.state_dict())
    opt = get_optimizer([dict(params=model.parameters(), lr=0.01)], cfg["trainer"]["optimizer"])
    ref_opt = torch.optim.SGD(ref_model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    step = SemiStep(model, teacher, opt, get_criterion(cfg), cfg, [[torch.zeros(0, 256)] for _ in range(C)],
                    [torch.zeros(1, dtype=torch.long) for _ in range(C)], [30000] * C, amp=False, channels_last=False)
    image_l, label_l, image_u = bench.synth_batch(7, 2, 2, crop, C)
    losses = step(image_l, label_l, image_u, 0, 0, 100)
    ref_model.train()
    pred = torch.nn.functional.interpolate(ref_model(image_l)["pred"], label_l.shape[1:], mode="bilinear", align_corners=True)
    ref_loss = torch.nn.functional.cross_entropy(pred, label_l, ignore_index=255)
    ref_opt.zero_grad()
    ref_loss.backward()
    ref_opt.step()
    assert abs(float(losses[0]) - float(ref_loss.detach())) <= 1e-5 and float(losses[1]) == 0.0 and float(losses[2]) == 0.0
    for (n, a), (_, b) in zip(model.named_parameters(), ref_model.named_parameters()):
        assert (a - b).abs().max() <= 1e-5, n
    t_after = teacher.state_dict()
    assert torch.equal(t_after["encoder.bn1.weight"], t_before["encoder.bn1.weight"])                 # no EMA in this branch
    assert not torch.equal(t_after["encoder.bn1.running_mean"], t_before["encoder.bn1.running_mean"])  # but BN statistics moved
