"""CPU: the drop-in model mirror (u2pl_b200/u2pl/models) and the oracle's functional restatement
(oracle/model_port.py) against outputs of the REFERENCE's ModelBuilder stored by oracle/make_golden.py.
Same torch seed => same parameters (construction order and init sequence mirror the reference)."""
import numpy as np
import pytest
import torch

import u2pl_b200


def _net_cfg(g):
    net = {"num_classes": int(g["C"]), "sync_bn": False, "ema_decay": 0.99,
           "encoder": {"type": "u2pl.models.resnet." + str(g["arch"]),
                       "kwargs": {"multi_grid": True, "zero_init_residual": True, "fpn": True,
                                  "replace_stride_with_dilation": [False, True, True], "pretrained": False}},
           "decoder": {"type": "u2pl.models.decoder.dec_deeplabv3_plus",
                       "kwargs": {"inner_planes": 256, "dilations": [12, 24, 36]}}}
    if bool(g["aux"]):
        net["aux_loss"] = {"aux_plane": 1024, "loss_weight": 0.4}
    return net


@pytest.mark.parametrize("name", ["model_r50_c21", "model_r50_c19_aux"])
def test_model_mirror_and_oracle_port(golden, name):
    u2pl_b200.install()
    from u2pl.models.model_helper import ModelBuilder
    from oracle import model_port
    g = golden(name)
    torch.manual_seed(int(g["seed"]))
    m = ModelBuilder(_net_cfg(g))
    assert [n for n, _ in m.named_parameters()] == g["param_names"].tolist()      # EMA zips parameters() in order
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"])
    assert np.array_equal(np.array([float(p.detach().double().sum()) for p in m.parameters()]), g["param_sums"])
    x = torch.from_numpy(g["x"])
    port = model_port.Net(model_port.state_from_module(m), str(g["arch"]), int(g["C"]), bool(g["aux"]))
    m.eval()
    port.training = False
    with torch.no_grad():
        out_m, out_p = m(x), port.forward(x)
    for k in out_m:
        ref = torch.from_numpy(g["eval_" + k])
        assert (out_m[k] - ref).abs().max() <= 1e-5 and (out_p[k] - ref).abs().max() <= 1e-5
    m.train()
    port.training = True
    torch.manual_seed(int(g["seed"]) + 1)
    out_m = m(x)
    torch.manual_seed(int(g["seed"]) + 1)
    out_p = port.forward(x)
    for k in out_m:
        ref = torch.from_numpy(g["train_" + k])
        assert (out_m[k].detach() - ref).abs().max() <= 1e-5 and (out_p[k].detach() - ref).abs().max() <= 1e-5


def test_dotted_type_plugin_api_and_helpers():
    u2pl_b200.install()
    import u2pl.models.resnet as R
    from u2pl.utils.lr_helper import get_optimizer, get_scheduler
    from u2pl.utils.utils import AverageMeter, intersectionAndUnion
    enc = R.resnet50(pretrained=False, fpn=True, replace_stride_with_dilation=[False, True, True], multi_grid=True)
    assert enc.get_outplanes() == 2048 and enc.get_auxplanes() == 1024
    opt = get_optimizer([dict(params=enc.parameters(), lr=0.01)], {"type": "SGD", "kwargs": {"lr": 0.01, "momentum": 0.9}})
    sch = get_scheduler({"epochs": 2, "lr_scheduler": {"mode": "poly", "kwargs": {"power": 0.9}}}, 10, opt)
    sch.step()
    assert sch.get_lr()[0] == 0.01
    sch.step()
    assert abs(sch.get_lr()[0] - 0.01 * (1 - 1 / 20) ** 0.9) < 1e-12 and opt.param_groups[0]["lr"] == sch.get_lr()[0]
    m = AverageMeter(2)
    for v in (1.0, 2.0, 4.0):
        m.update(v)
    assert m.val == 4.0 and m.avg == 3.0
    i, u, t = intersectionAndUnion(np.array([0, 1, 1, 2]), np.array([0, 1, 2, 255]), 3)
    assert i.tolist() == [1, 1, 0] and t.tolist() == [1, 1, 1]
