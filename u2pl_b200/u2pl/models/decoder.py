"""DeepLabv3 / DeepLabv3+ decoders and the auxiliary head.  Mirrors u2pl/models/decoder.py:8-142
(module names, parameter order and default initialisation identical to the reference)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from u2pl_b200.fused import DilatedConv2d, run_sequential

from .base import ASPP, _norm


def _head3x3(cin, cout, norm, bias):
    conv = nn.Conv2d if bias else DilatedConv2d          # bias-free 1280->256 head: GEMM-based weight gradient
    return [conv(cin, cout, kernel_size=3, stride=1, padding=1, bias=bias), norm(cout),
            nn.ReLU(inplace=True), nn.Dropout2d(0.1)]


class dec_deeplabv3(nn.Module):
    def __init__(self, in_planes, num_classes=19, inner_planes=256, sync_bn=False, dilations=(12, 24, 36)):
        super().__init__()
        norm = _norm(sync_bn)
        self.aspp = ASPP(in_planes, inner_planes=inner_planes, sync_bn=sync_bn, dilations=dilations)
        self.head = nn.Sequential(*_head3x3(self.aspp.get_outplanes(), 256, norm, False),
                                  nn.Conv2d(256, num_classes, kernel_size=1, stride=1, padding=0, bias=True))

    def forward(self, x):
        return run_sequential(self.head, self.aspp(x))


class dec_deeplabv3_plus(nn.Module):
    """ASPP on x4, fused with a 1x1 projection of x1; two parallel heads: classifier -> `pred`,
    representation -> `rep` (reference decoder.py:45-124)."""

    def __init__(self, in_planes, num_classes=19, inner_planes=256, sync_bn=False, dilations=(12, 24, 36),
                 rep_head=True):
        super().__init__()
        norm = _norm(sync_bn)
        self.rep_head = rep_head
        self.low_conv = nn.Sequential(nn.Conv2d(256, 256, kernel_size=1), norm(256), nn.ReLU(inplace=True))
        self.aspp = ASPP(in_planes, inner_planes=inner_planes, sync_bn=sync_bn, dilations=dilations)
        self.head = nn.Sequential(*_head3x3(self.aspp.get_outplanes(), 256, norm, False))
        self.classifier = nn.Sequential(*_head3x3(512, 256, norm, True), *_head3x3(256, 256, norm, True),
                                        nn.Conv2d(256, num_classes, kernel_size=1, stride=1, padding=0, bias=True))
        if self.rep_head:
            self.representation = nn.Sequential(*_head3x3(512, 256, norm, True), *_head3x3(256, 256, norm, True),
                                                nn.Conv2d(256, 256, kernel_size=1, stride=1, padding=0, bias=True))

    def forward(self, x):
        x1, _, _, x4 = x
        deep = run_sequential(self.head, self.aspp(x4))
        low = run_sequential(self.low_conv, x1)
        deep = F.interpolate(deep, size=low.shape[-2:], mode="bilinear", align_corners=True)
        feat = torch.cat((low, deep), dim=1)
        out = {"pred": run_sequential(self.classifier, feat)}
        if self.rep_head:
            out["rep"] = run_sequential(self.representation, feat)
        return out


class Aux_Module(nn.Module):
    def __init__(self, in_planes, num_classes=19, sync_bn=False):
        super().__init__()
        norm = _norm(sync_bn)
        self.aux = nn.Sequential(*_head3x3(in_planes, 256, norm, True),
                                 nn.Conv2d(256, num_classes, kernel_size=1, stride=1, padding=0, bias=True))

    def forward(self, x):
        return run_sequential(self.aux, x)
