"""DeepLabv3 / DeepLabv3+ decoders and the auxiliary head for the drop-in `u2pl.models.decoder.*` types.
Module names, parameter order and PyTorch-default initialisation follow u2pl/models/decoder.py:8-142, so
checkpoints and seeded constructions are interchangeable with the reference; every (BN, ReLU) pair runs
through the fused kernels (`run_sequential`)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from u2pl_b200.fused import DilatedConv2d, run_sequential

from .base import ASPP, _norm


def _unit(cin, cout, norm, bias):
    """3x3 conv -> norm -> ReLU -> Dropout2d(0.1): the building unit of every head."""
    conv_type = nn.Conv2d if bias else DilatedConv2d          # bias-free 1280->256 head: GEMM-based weight gradient
    return [conv_type(cin, cout, 3, 1, 1, bias=bias), norm(cout), nn.ReLU(inplace=True), nn.Dropout2d(0.1)]


def _project(cin, cout):
    return nn.Conv2d(cin, cout, 1, 1, 0, bias=True)


class dec_deeplabv3(nn.Module):
    def __init__(self, in_planes, num_classes=19, inner_planes=256, sync_bn=False, dilations=(12, 24, 36)):
        super().__init__()
        self.aspp = ASPP(in_planes, inner_planes=inner_planes, sync_bn=sync_bn, dilations=dilations)
        self.head = nn.Sequential(*_unit(self.aspp.get_outplanes(), 256, _norm(sync_bn), False), _project(256, num_classes))

    def forward(self, x):
        return run_sequential(self.head, self.aspp(x))


class dec_deeplabv3_plus(nn.Module):
    """ASPP on the last stage, fused with a 1x1 projection of stage 1; `classifier` -> "pred" and the parallel
    `representation` head -> "rep" (256-d pixel embeddings for the contrastive loss)."""

    def __init__(self, in_planes, num_classes=19, inner_planes=256, sync_bn=False, dilations=(12, 24, 36),
                 rep_head=True):
        super().__init__()
        norm = _norm(sync_bn)
        self.rep_head = rep_head
        self.low_conv = nn.Sequential(nn.Conv2d(256, 256, 1), norm(256), nn.ReLU(inplace=True))
        self.aspp = ASPP(in_planes, inner_planes=inner_planes, sync_bn=sync_bn, dilations=dilations)
        self.head = nn.Sequential(*_unit(self.aspp.get_outplanes(), 256, norm, False))
        heads = [("classifier", num_classes)] + ([("representation", 256)] if rep_head else [])
        for name, width in heads:                                  # registration order: classifier, representation
            setattr(self, name, nn.Sequential(*_unit(512, 256, norm, True), *_unit(256, 256, norm, True), _project(256, width)))

    def forward(self, x):
        stage1, stage4 = x[0], x[3]
        deep = run_sequential(self.head, self.aspp(stage4))
        low = run_sequential(self.low_conv, stage1)
        fused = torch.cat((low, F.interpolate(deep, size=low.shape[-2:], mode="bilinear", align_corners=True)), dim=1)
        out = {"pred": run_sequential(self.classifier, fused)}
        # `skip_rep` (set by u2pl_b200.step for the teacher's pseudo-label forward, train_semi.py:318-319, which only reads
        # "pred"): the reference computes the 61 GFLOP/image representation head there and drops the result
        if self.rep_head and not getattr(self, "skip_rep", False):
            out["rep"] = run_sequential(self.representation, fused)
        return out


class Aux_Module(nn.Module):
    def __init__(self, in_planes, num_classes=19, sync_bn=False):
        super().__init__()
        self.aux = nn.Sequential(*_unit(in_planes, 256, _norm(sync_bn), True), _project(256, num_classes))

    def forward(self, x):
        return run_sequential(self.aux, x)
