"""ASPP + norm-layer selection.  Mirrors u2pl/models/base.py:6-100 of the reference (same module
names / parameter order, so state_dicts and seeded initialisation are interchangeable)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from u2pl_b200.fused import DilatedConv2d, run_sequential


def get_syncbn():
    # reference base.py:6-8
    return nn.SyncBatchNorm


def _norm(sync_bn):
    return get_syncbn() if sync_bn else nn.BatchNorm2d


def _conv_bn_relu(cin, cout, k, dilation, norm, pool=False):
    pad = 0 if k == 1 else dilation
    layers = [nn.AdaptiveAvgPool2d((1, 1))] if pool else []
    conv = DilatedConv2d if (k == 3 and dilation > 1) else nn.Conv2d      # same parameters, faster weight gradient
    layers += [conv(cin, cout, kernel_size=k, padding=pad, dilation=dilation, bias=False),
               norm(cout), nn.ReLU(inplace=True)]
    return nn.Sequential(*layers)


class ASPP(nn.Module):
    """Image-pool branch + 1x1 + three dilated 3x3 branches, concatenated (reference base.py:11-100)."""

    def __init__(self, in_planes, inner_planes=256, sync_bn=False, dilations=(12, 24, 36)):
        super().__init__()
        norm = _norm(sync_bn)
        self.conv1 = _conv_bn_relu(in_planes, inner_planes, 1, 1, norm, pool=True)
        self.conv2 = _conv_bn_relu(in_planes, inner_planes, 1, 1, norm)
        self.conv3 = _conv_bn_relu(in_planes, inner_planes, 3, dilations[0], norm)
        self.conv4 = _conv_bn_relu(in_planes, inner_planes, 3, dilations[1], norm)
        self.conv5 = _conv_bn_relu(in_planes, inner_planes, 3, dilations[2], norm)
        self.out_planes = (len(dilations) + 2) * inner_planes

    def get_outplanes(self):
        return self.out_planes

    def forward(self, x):
        h, w = x.shape[-2:]
        pooled = F.interpolate(self.conv1(x), size=(h, w), mode="bilinear", align_corners=True)   # 1x1 map: BN stays ATen
        branches = [run_sequential(b, x) for b in (self.conv2, self.conv3, self.conv4, self.conv5)]
        return torch.cat([pooled] + branches, 1)
