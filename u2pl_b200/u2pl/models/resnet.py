"""ResNet-v1c encoders (deep 3x3 stem, ceil-mode max-pool, stride->dilation, multi-grid).
Mirrors u2pl/models/resnet.py:25-402: same module tree, registration order and initialisation
sequence, so a state_dict or a torch seed produces the same weights as the reference."""
import torch
import torch.nn as nn

from u2pl_b200.fused import DilatedConv2d, StemConv2d, bn_act, run_sequential

from .base import _norm

__all__ = ["ResNet", "resnet18", "resnet34", "resnet50", "resnet101", "resnet152"]

model_urls = {name: f"/path/to/{name}.pth" for name in ("resnet18", "resnet34", "resnet50", "resnet101", "resnet152")}


def conv3x3(cin, cout, stride=1, groups=1, dilation=1):
    # strided or strongly dilated 3x3 (layer2.0, layer4 d=4/8/16): GEMM-based weight gradient (fused.py);
    # cuDNN 9 falls back to wgrad_alg0_engine for some of them on B200 (18 ms for one 512->512 d=16 layer)
    conv = DilatedConv2d if (stride != 1 or dilation >= 4) else nn.Conv2d
    return conv(cin, cout, kernel_size=3, stride=stride, padding=dilation, groups=groups, bias=False,
                dilation=dilation)


def conv1x1(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=1, stride=stride, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 norm_layer=None):
        super().__init__()
        norm_layer = norm_layer or nn.BatchNorm2d
        if groups != 1 or base_width != 64:
            raise ValueError("BasicBlock only supports groups=1 and base_width=64")
        if dilation > 1:
            raise NotImplementedError("Dilation > 1 not supported in BasicBlock")
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = norm_layer(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = norm_layer(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = bn_act(self.conv1(x), self.bn1, self.relu)
        identity = x if self.downsample is None else run_sequential(self.downsample, x)
        return bn_act(self.conv2(y), self.bn2, self.relu, residual=identity)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 norm_layer=nn.BatchNorm2d):
        super().__init__()
        width = int(planes * (base_width / 64.0)) * groups
        self.conv1 = conv1x1(inplanes, width)
        self.bn1 = norm_layer(width)
        self.conv2 = conv3x3(width, width, stride, groups, dilation)
        self.bn2 = norm_layer(width)
        self.conv3 = conv1x1(width, planes * self.expansion)
        self.bn3 = norm_layer(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = bn_act(self.conv1(x), self.bn1, self.relu)
        y = bn_act(self.conv2(y), self.bn2, self.relu)
        identity = x if self.downsample is None else run_sequential(self.downsample, x)
        return bn_act(self.conv3(y), self.bn3, self.relu, residual=identity)      # relu(bn3(.) + identity)


class ResNet(nn.Module):
    def __init__(self, block, layers, zero_init_residual=False, groups=1, width_per_group=64,
                 replace_stride_with_dilation=(False, False, False), sync_bn=False, multi_grid=False, fpn=False):
        super().__init__()
        norm_layer = _norm(sync_bn)
        self._norm_layer = norm_layer
        self.inplanes, self.dilation = 128, 1
        if replace_stride_with_dilation is None:
            replace_stride_with_dilation = [False, False, False]
        if len(replace_stride_with_dilation) != 3:
            raise ValueError("replace_stride_with_dilation should be None or a 3-element tuple, "
                             f"got {replace_stride_with_dilation}")
        self.groups, self.base_width, self.fpn = groups, width_per_group, fpn
        stem = StemConv2d(3, 64, kernel_size=3, stride=2, padding=1, bias=False, dilation=1)   # = conv3x3(3, 64, 2)
        self.conv1 = nn.Sequential(stem, norm_layer(64), nn.ReLU(inplace=True),
                                   conv3x3(64, 64), norm_layer(64), nn.ReLU(inplace=True),
                                   conv3x3(64, self.inplanes))
        self.bn1 = norm_layer(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1, ceil_mode=True)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2, dilate=replace_stride_with_dilation[0])
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2, dilate=replace_stride_with_dilation[1])
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2, dilate=replace_stride_with_dilation[2],
                                       multi_grid=multi_grid)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm, nn.SyncBatchNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.constant_(m.bn3.weight, 0)
                elif isinstance(m, BasicBlock):
                    nn.init.constant_(m.bn2.weight, 0)

    def get_outplanes(self):
        return self.inplanes

    def get_auxplanes(self):
        return self.inplanes // 2

    def _make_layer(self, block, planes, blocks, stride=1, dilate=False, multi_grid=False):
        norm_layer = self._norm_layer
        prev_dilation = self.dilation
        if dilate:
            self.dilation *= stride
            stride = 1
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(conv1x1(self.inplanes, planes * block.expansion, stride),
                                       norm_layer(planes * block.expansion))
        grids = [2, 2, 4] if multi_grid else [1] * blocks
        stack = [block(self.inplanes, planes, stride, downsample, self.groups, self.base_width,
                       prev_dilation * grids[0], norm_layer)]
        self.inplanes = planes * block.expansion
        for i in range(1, blocks):
            stack.append(block(self.inplanes, planes, groups=self.groups, base_width=self.base_width,
                               dilation=self.dilation * grids[i], norm_layer=norm_layer))
        return nn.Sequential(*stack)

    def forward(self, x):
        x = self.maxpool(bn_act(run_sequential(self.conv1, x), self.bn1, self.relu))
        x1 = self.layer1(x)
        x2 = self.layer2(x1)
        x3 = self.layer3(x2)
        x4 = self.layer4(x3)
        return [x1, x2, x3, x4] if self.fpn else [x3, x4]


def _build(name, block, layers, pretrained, **kwargs):
    model = ResNet(block, layers, **kwargs)
    if pretrained:
        state = torch.load(model_urls[name])
        missing, unexpected = model.load_state_dict(state, strict=False)
        print(f"[Info] Load ImageNet pretrain from '{model_urls[name]}'", "\nmissing_keys: ", missing,
              "\nunexpected_keys: ", unexpected)
    return model


def resnet18(pretrained=False, **kwargs):
    return _build("resnet18", BasicBlock, [2, 2, 2, 2], pretrained, **kwargs)


def resnet34(pretrained=False, **kwargs):
    return _build("resnet34", BasicBlock, [3, 4, 6, 3], pretrained, **kwargs)


def resnet50(pretrained=True, **kwargs):
    return _build("resnet50", Bottleneck, [3, 4, 6, 3], pretrained, **kwargs)


def resnet101(pretrained=True, **kwargs):
    return _build("resnet101", Bottleneck, [3, 4, 23, 3], pretrained, **kwargs)


def resnet152(pretrained=True, **kwargs):
    return _build("resnet152", Bottleneck, [3, 8, 36, 3], pretrained, **kwargs)
