"""GPU parity: the fused SGD + EMA multi-tensor kernel (csrc/sgd_ema.cu, u2pl_b200.optim.FusedSGDEMA) against
torch.optim.SGD (what lr_helper.get_optimizer builds, lr_helper.py:12-27) followed by the reference's EMA loop
(train_semi.py:531-548), over several steps with a changing learning rate and two parameter groups."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    shapes = [(64, 3, 3, 3), (64,), (256, 64, 1, 1), (1000,), (7,), (3, 5, 7), (512, 512, 3, 3), (1,), (8193,), (2048, 21)]
    return [torch.randn(s, device="cuda", generator=g) for s in shapes]


def test_fused_sgd_ema_matches_torch():
    from u2pl_b200.optim import FusedSGDEMA
    init = _make(0)
    ref_p = [torch.nn.Parameter(t.clone()) for t in init]
    my_p = [torch.nn.Parameter(t.clone()) for t in init]
    ref_t = [t.clone() * 0.9 for t in init]
    my_t = [t.clone() * 0.9 for t in init]

    def groups(ps):
        return [dict(params=ps[:4], lr=0.01), dict(params=ps[4:], lr=0.1)]

    kw = dict(momentum=0.9, weight_decay=1e-4)
    ref_opt = torch.optim.SGD(groups(ref_p), lr=0.01, **kw)
    my_opt = torch.optim.SGD(groups(my_p), lr=0.01, **kw)
    fused = FusedSGDEMA(my_opt, my_p, my_t)
    for step in range(4):
        grads = _make(100 + step)
        for p, q, g in zip(ref_p, my_p, grads):
            p.grad = g.clone()
            q.grad = g.clone()
        for opt in (ref_opt, my_opt):                                   # the poly schedule rewrites group["lr"] every iteration
            for gi, grp in enumerate(opt.param_groups):
                grp["lr"] = (0.01 if gi == 0 else 0.1) * (1 - step / 10) ** 0.9
        decay = min(1 - 1 / (step + 2), 0.99)
        ref_opt.step()
        with torch.no_grad():
            for t, s in zip(ref_t, ref_p):                              # train_semi.py:543-548
                t.copy_(decay * t + (1 - decay) * s)
        fused.step(decay if step != 2 else None)                        # step 2: plain SGD, no EMA
        if step == 2:
            with torch.no_grad():
                for t, s in zip(my_t, my_p):
                    t.copy_(decay * t + (1 - decay) * s)
        for a, b in zip(ref_p, my_p):
            assert torch.allclose(a, b, rtol=2e-6, atol=5e-7), step
        for a, b in zip(ref_t, my_t):
            assert torch.allclose(a, b, rtol=2e-6, atol=5e-7), step
        for a, b in zip(ref_p, my_p):
            assert torch.allclose(ref_opt.state[a]["momentum_buffer"], my_opt.state[b]["momentum_buffer"], rtol=2e-6, atol=5e-7)
    # torch's own step keeps working on the state the fused kernel maintained (checkpoint / fallback compatibility)
    for p, q, g in zip(ref_p, my_p, _make(999)):
        p.grad, q.grad = g.clone(), g.clone()
    ref_opt.step()
    my_opt.step()
    for a, b in zip(ref_p, my_p):
        assert torch.allclose(a, b, rtol=2e-6, atol=5e-7)
