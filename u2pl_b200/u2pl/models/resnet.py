"""ResNet-v1c encoders (deep 3x3 stem, ceil-mode max-pool, stride->dilation, multi-grid) for the
drop-in `u2pl.models.resnet.*` dotted types.

Everything a checkpoint or a seeded construction can observe is kept identical to the reference
(u2pl/models/resnet.py:25-402): module names and registration order, parameter order, the order in
which convolutions are constructed (it fixes the RNG stream) and the re-initialisation pass.  The
code itself is organised around two small tables (`_BLOCK_SPECS`, `_ARCHS`) instead of hand-written
classes per variant, and the forward passes go through the fused BN(+ReLU+residual) kernels.
"""
import torch
import torch.nn as nn

from u2pl_b200.fused import (DilatedConv2d, StemConv2d, bn_act, chain_ok, conv_bn_act, conv_bn_relu_chain,
                             max_pool, run_sequential)

from .base import _norm

model_urls = {f"resnet{d}": f"/path/to/resnet{d}.pth" for d in (18, 34, 50, 101, 152)}


def conv3x3(cin, cout, stride=1, groups=1, dilation=1):
    # strided or strongly dilated 3x3 (layer2.0, layer4 d=4/8/16): GEMM-based weight gradient (fused.py);
    # cuDNN 9 falls back to wgrad_alg0_engine for some of them on B200 (18 ms for one 512->512 d=16 layer)
    kind = DilatedConv2d if (stride != 1 or dilation >= 4) else nn.Conv2d
    return kind(cin, cout, 3, stride, dilation, dilation, groups, False)


def conv1x1(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 1, stride, bias=False)


# (kernel, uses the block's stride/dilation?) for conv1..convN of each residual block type
_BLOCK_SPECS = {"basic": (1, [(3, True), (3, False)]), "bottleneck": (4, [(1, False), (3, True), (1, False)])}


class _Residual(nn.Module):
    """conv/bn pairs named conv1, bn1, ... in the reference's registration order, then relu, downsample."""
    kind = None

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 norm_layer=None):
        super().__init__()
        norm_layer = norm_layer or nn.BatchNorm2d
        expansion, convs = _BLOCK_SPECS[self.kind]
        if self.kind == "basic":
            if groups != 1 or base_width != 64:
                raise ValueError("BasicBlock only supports groups=1 and base_width=64")
            if dilation > 1:
                raise NotImplementedError("Dilation > 1 not supported in BasicBlock")
            widths = [planes, planes]
        else:
            inner = int(planes * (base_width / 64.0)) * groups
            widths = [inner, inner, planes * expansion]
        cin = inplanes
        for i, ((k, strided), cout) in enumerate(zip(convs, widths), start=1):
            if k == 1:
                conv = conv1x1(cin, cout)
            elif self.kind == "basic":
                conv = conv3x3(cin, cout, stride if strided else 1)
            else:
                conv = conv3x3(cin, cout, stride, groups, dilation)
            setattr(self, f"conv{i}", conv)
            setattr(self, f"bn{i}", norm_layer(cout))
            cin = cout
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        self._depth = len(convs)

    def forward(self, x):
        shortcut = x if self.downsample is None else run_sequential(self.downsample, x)
        convs = [getattr(self, f"conv{i}") for i in range(1, self._depth + 1)]
        bns = [getattr(self, f"bn{i}") for i in range(1, self._depth + 1)]
        if chain_ok(x, convs, bns, shortcut):                   # no-grad train-mode forward (teacher T2), opt-in
            return conv_bn_relu_chain(x, convs, bns, True, residual=shortcut)
        y = x
        for i in range(1, self._depth):
            y = conv_bn_act(y, getattr(self, f"conv{i}"), getattr(self, f"bn{i}"), self.relu)
        last = self._depth
        return conv_bn_act(y, getattr(self, f"conv{last}"), getattr(self, f"bn{last}"), self.relu, residual=shortcut)


class BasicBlock(_Residual):
    kind, expansion = "basic", 1


class Bottleneck(_Residual):
    kind, expansion = "bottleneck", 4


class ResNet(nn.Module):
    def __init__(self, block, layers, zero_init_residual=False, groups=1, width_per_group=64,
                 replace_stride_with_dilation=(False, False, False), sync_bn=False, multi_grid=False, fpn=False):
        super().__init__()
        if replace_stride_with_dilation is None:
            replace_stride_with_dilation = (False, False, False)
        if len(replace_stride_with_dilation) != 3:
            raise ValueError(f"replace_stride_with_dilation should be None or a 3-element tuple, got "
                             f"{replace_stride_with_dilation}")
        norm = self._norm_layer = _norm(sync_bn)
        self.inplanes, self.dilation, self.groups, self.base_width, self.fpn = 128, 1, groups, width_per_group, fpn
        stem = [StemConv2d(3, 64, 3, 2, 1, 1, 1, False), norm(64), nn.ReLU(inplace=True),       # == conv3x3(3, 64, 2)
                conv3x3(64, 64), norm(64), nn.ReLU(inplace=True), conv3x3(64, self.inplanes)]
        self.conv1 = nn.Sequential(*stem)
        self.bn1 = norm(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1, ceil_mode=True)
        stage_cfg = [(64, 1, False, False)] + [(128 << i, 2, bool(replace_stride_with_dilation[i]), multi_grid and i == 2)
                                               for i in range(3)]
        for idx, ((planes, stride, dilate, grid), depth) in enumerate(zip(stage_cfg, layers), start=1):
            setattr(self, f"layer{idx}", self._make_layer(block, planes, depth, stride, dilate, grid))
        self._reset_parameters(zero_init_residual)

    def _reset_parameters(self, zero_init_residual):
        for m in self.modules():                       # same traversal order as the reference (RNG stream)
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm, nn.SyncBatchNorm)):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, _Residual):
                    nn.init.zeros_(getattr(m, f"bn{m._depth}").weight)

    def get_outplanes(self):
        return self.inplanes

    def get_auxplanes(self):
        return self.inplanes // 2

    def _make_layer(self, block, planes, blocks, stride=1, dilate=False, multi_grid=False):
        first_dilation = self.dilation
        if dilate:
            self.dilation, stride = self.dilation * stride, 1
        out_planes = planes * block.expansion
        shortcut = None
        if stride != 1 or self.inplanes != out_planes:
            shortcut = nn.Sequential(conv1x1(self.inplanes, out_planes, stride), self._norm_layer(out_planes))
        grids = [2, 2, 4] if multi_grid else [1] * blocks
        common = dict(groups=self.groups, base_width=self.base_width, norm_layer=self._norm_layer)
        stack = [block(self.inplanes, planes, stride, shortcut, dilation=first_dilation * grids[0], **common)]
        self.inplanes = out_planes
        stack += [block(out_planes, planes, dilation=self.dilation * grids[i], **common) for i in range(1, blocks)]
        return nn.Sequential(*stack)

    def forward(self, x):
        x = max_pool(bn_act(run_sequential(self.conv1, x), self.bn1, self.relu), self.maxpool)
        feats = []
        for idx in range(1, 5):
            x = getattr(self, f"layer{idx}")(x)
            feats.append(x)
        return feats if self.fpn else feats[2:]


_ARCHS = {"resnet18": (BasicBlock, [2, 2, 2, 2], False), "resnet34": (BasicBlock, [3, 4, 6, 3], False),
          "resnet50": (Bottleneck, [3, 4, 6, 3], True), "resnet101": (Bottleneck, [3, 4, 23, 3], True),
          "resnet152": (Bottleneck, [3, 8, 36, 3], True)}


def _factory(name):
    block, depths, pretrained_default = _ARCHS[name]

    def build(pretrained=pretrained_default, **kwargs):
        net = ResNet(block, depths, **kwargs)
        if pretrained:
            missing, unexpected = net.load_state_dict(torch.load(model_urls[name]), strict=False)
            print(f"[Info] Load ImageNet pretrain from '{model_urls[name]}'", "\nmissing_keys: ", missing,
                  "\nunexpected_keys: ", unexpected)
        return net

    build.__name__ = name
    build.__doc__ = f"{name} encoder (reference u2pl/models/resnet.py); pretrained defaults to {pretrained_default}."
    return build


resnet18, resnet34, resnet50, resnet101, resnet152 = (_factory(n) for n in _ARCHS)
__all__ = ["ResNet"] + list(_ARCHS)
