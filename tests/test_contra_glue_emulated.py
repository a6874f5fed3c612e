"""CPU: u2pl_b200/contra.py END TO END (host control flow, ring-bank bookkeeping, sampling order, plan construction,
autograd wrapper) with the contrastive C-ABI entry points EMULATED inside this test from their header contracts, against
the oracle and the reference's own fixtures (tests/golden/contra_*).  The real kernels are compared with the same oracle
on the GPU (tests/test_gpu_contra.py); this file keeps the layer above them under test where there is no GPU.
Test-only emulation: the product has no CPU path."""
import numpy as np
import pytest
import torch

from oracle import port
import emulated_abi
from u2pl_b200 import contra

@pytest.fixture
def emulated(monkeypatch):
    fake = emulated_abi.install(monkeypatch)
    yield fake
    contra.forget_banks()


def _cfg(g):
    cfg = {k: v for k, v in zip(g["cfg_keys"].tolist(), g["cfg_vals"].tolist())}
    for k in ("low_rank", "high_rank", "num_negatives", "num_queries"):
        cfg[k] = int(cfg[k])
    return cfg


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("name", ["contra_c21", "contra_c19_missing", "contra_c21_driver_onehot"])
def test_contra_host_logic_matches_oracle(golden, emulated, name, channels_last):
    g = golden(name)
    cfg = _cfg(g)
    C, D = g["s0_label_l"].shape[1], g["s0_rep"].shape[1]
    qsize = g["queue_size"].tolist()
    memobank = [[torch.zeros(0, D)] for _ in range(C)]
    ptrs = [torch.zeros(1, dtype=torch.long) for _ in range(C)]
    o_bank = [[np.zeros((0, D), np.float32)] for _ in range(C)]
    o_ptr = [[0] for _ in range(C)]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))                                 # noqa: E731
    for s in range(int(g["steps"])):
        rep, rep_t = t(g[f"s{s}_rep"]), t(g[f"s{s}_rep_teacher"])
        if channels_last:
            rep, rep_t = rep.contiguous(memory_format=torch.channels_last), rep_t.contiguous(memory_format=torch.channels_last)
        rep.requires_grad_(True)
        args = (g[f"s{s}_label_l"].astype(np.int64), g[f"s{s}_label_u"].astype(np.int64), g[f"s{s}_prob_l"],
                g[f"s{s}_prob_u"], g[f"s{s}_low_mask"].astype(np.float32), g[f"s{s}_high_mask"].astype(np.float32))
        torch.manual_seed(1000 + s)
        new_keys, loss, plan = contra.compute_contra_memobank_loss(rep, *[t(a) for a in args], cfg, memobank, ptrs, qsize, rep_t,
                                                                   return_plan=True)
        loss.backward()
        torch.manual_seed(1000 + s)
        out = port.compute_contra_memobank_loss(g[f"s{s}_rep"], *args, cfg, o_bank, o_ptr, qsize, g[f"s{s}_rep_teacher"], want_grad=True)
        assert new_keys == out["new_keys"] == g[f"s{s}_new_keys"].tolist()
        assert plan["valid_classes"] == out["valid_classes"]
        assert [int(p[0]) for p in ptrs] == [int(p[0]) for p in o_ptr] == g[f"s{s}_ptr"].tolist()
        assert [m[0].shape[0] for m in memobank] == g[f"s{s}_bank_len"].tolist()
        bank = contra.bank_for(memobank, qsize, D, rep.device)
        for c in range(C):
            assert np.array_equal(bank.materialize(c).numpy(), o_bank[c][0])                 # FIFO content
        if plan["nact"]:
            pix = plan["anchor_pix"].numpy().reshape(plan["nact"], -1)
            for a, smp in enumerate(out["sampled"]):
                assert np.array_equal(pix[a], out["sel"]["anchors"][smp["j"]][smp["a_idx"]])  # same anchors, same order of draws
        assert abs(float(loss.detach()) - float(out["loss"])) <= 1e-5
        grad = rep.grad.numpy() if rep.grad is not None else np.zeros_like(g[f"s{s}_rep"])
        assert np.abs(grad - out["rep_grad"]).max() <= 1e-6
