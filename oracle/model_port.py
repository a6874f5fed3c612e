"""oracle/model_port.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

Plain-PyTorch fp32 functional restatement of the reference network (ResNet-v1c encoder,
ASPP, DeepLabv3+ decoder with classifier + representation heads, optional aux head):
reference u2pl/models/resnet.py:93-292, base.py:11-100, decoder.py:45-142,
model_helper.py:49-66.  It walks a flat {name: tensor} state that uses the reference's own
state_dict keys, so any checkpoint / seeded ModelBuilder of the reference (or of the product
mirror) can be fed through it.  Used (a) as the fp32 checker for the product model in tests and
(b) as the network of the CPU reference arm in bench.py (--impl reference / cpu_baseline).
"""
import math

import torch
import torch.nn.functional as F

LAYERS = {"resnet50": [3, 4, 6, 3], "resnet101": [3, 4, 23, 3], "resnet152": [3, 8, 36, 3]}


class Net:
    """state: dict name -> tensor (parameters require grad, BN buffers do not)."""

    def __init__(self, state, arch="resnet101", num_classes=21, aux=False, dilations=(12, 24, 36),
                 replace_stride_with_dilation=(False, True, True), multi_grid=True, momentum=0.1, eps=1e-5):
        self.s, self.arch, self.C, self.aux = state, arch, num_classes, aux
        self.dil, self.rswd, self.mg = dilations, replace_stride_with_dilation, multi_grid
        self.momentum, self.eps = momentum, eps
        self.training = True
        self.dropout_p = 0.1

    # -- primitives ----------------------------------------------------------------
    def conv(self, x, name, stride=1, dilation=1):
        w = self.s[name + ".weight"]
        pad = dilation * (w.shape[-1] // 2)
        return F.conv2d(x, w, self.s.get(name + ".bias"), stride=stride, padding=pad, dilation=dilation)

    def bn(self, x, name):
        s = self.s
        if self.training and name + ".num_batches_tracked" in s:
            s[name + ".num_batches_tracked"] += 1
        return F.batch_norm(x, s[name + ".running_mean"], s[name + ".running_var"], s[name + ".weight"],
                            s[name + ".bias"], self.training, self.momentum, self.eps)

    def cbr(self, x, conv, bn, **kw):
        return F.relu(self.bn(self.conv(x, conv, **kw), bn))

    def drop(self, x):
        return F.dropout2d(x, self.dropout_p, self.training) if self.dropout_p > 0 else x

    # -- encoder (resnet.py:279-292) -----------------------------------------------
    def bottleneck(self, x, p, stride, dilation, has_down):
        y = self.cbr(x, p + ".conv1", p + ".bn1")
        y = self.cbr(y, p + ".conv2", p + ".bn2", stride=stride, dilation=dilation)
        y = self.bn(self.conv(y, p + ".conv3"), p + ".bn3")
        if has_down:
            x = self.bn(self.conv(x, p + ".downsample.0", stride=stride), p + ".downsample.1")
        return F.relu(y + x)

    def encoder(self, x):
        e = "encoder."
        x = self.cbr(x, e + "conv1.0", e + "conv1.1", stride=2)
        x = self.cbr(x, e + "conv1.3", e + "conv1.4")
        x = self.cbr(x, e + "conv1.6", e + "bn1")
        x = F.max_pool2d(x, 3, 2, 1, ceil_mode=True)
        feats, dilation = [], 1
        for li, blocks in enumerate(LAYERS[self.arch]):
            stride = 1 if li == 0 else 2
            prev = dilation
            if li > 0 and self.rswd[li - 1]:
                dilation *= stride
                stride = 1
            grids = [2, 2, 4] if (li == 3 and self.mg) else [1] * blocks
            for bi in range(blocks):
                p = f"{e}layer{li + 1}.{bi}"
                if bi == 0:
                    x = self.bottleneck(x, p, stride, prev * grids[0], True)
                else:
                    x = self.bottleneck(x, p, 1, dilation * grids[bi], False)
            feats.append(x)
        return feats

    # -- decoder (base.py:90-100, decoder.py:108-124) --------------------------------
    def head2(self, x, p):
        x = self.drop(self.cbr(x, p + ".0", p + ".1"))
        x = self.drop(self.cbr(x, p + ".4", p + ".5"))
        return self.conv(x, p + ".8")

    def forward(self, x):
        x1, x2, x3, x4 = self.encoder(x)
        d = "decoder."
        h, w = x4.shape[-2:]
        pooled = self.cbr(F.adaptive_avg_pool2d(x4, 1), d + "aspp.conv1.1", d + "aspp.conv1.2")
        branches = [F.interpolate(pooled, (h, w), mode="bilinear", align_corners=True),
                    self.cbr(x4, d + "aspp.conv2.0", d + "aspp.conv2.1")]
        for i, dl in enumerate(self.dil):
            branches.append(self.cbr(x4, f"{d}aspp.conv{3 + i}.0", f"{d}aspp.conv{3 + i}.1", dilation=dl))
        deep = self.drop(self.cbr(torch.cat(branches, 1), d + "head.0", d + "head.1"))
        low = self.cbr(x1, d + "low_conv.0", d + "low_conv.1")
        deep = F.interpolate(deep, low.shape[-2:], mode="bilinear", align_corners=True)
        feat = torch.cat((low, deep), 1)
        out = {"pred": self.head2(feat, d + "classifier"), "rep": self.head2(feat, d + "representation")}
        if self.aux:
            a = self.drop(self.cbr(x3, "auxor.aux.0", "auxor.aux.1"))
            out["aux"] = self.conv(a, "auxor.aux.4")
        return out

    def parameters(self):
        return [v for k, v in self.s.items() if v.requires_grad]


def init_state(arch="resnet101", num_classes=21, aux=False, seed=0, peak=1.0):
    """Random initialisation with the reference's schemes (kaiming-normal fan_out for the encoder,
    resnet.py:209-224; PyTorch defaults for the decoder) -- NOT the reference's RNG stream."""
    g = torch.Generator().manual_seed(seed)
    s = {}

    def conv(name, cin, cout, k, bias=False, encoder=True):
        fan_out, fan_in = cout * k * k, cin * k * k
        if encoder:
            w = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / fan_out)
        else:
            bound = 1.0 / math.sqrt(fan_in)
            w = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        s[name + ".weight"] = w.requires_grad_(True)
        if bias:
            s[name + ".bias"] = ((torch.rand(cout, generator=g) * 2 - 1) / math.sqrt(fan_in)).requires_grad_(True)

    def bn(name, c, zero=False):
        s[name + ".weight"] = (torch.zeros(c) if zero else torch.ones(c)).requires_grad_(True)
        s[name + ".bias"] = torch.zeros(c).requires_grad_(True)
        s[name + ".running_mean"] = torch.zeros(c)
        s[name + ".running_var"] = torch.ones(c)
        s[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    e = "encoder."
    conv(e + "conv1.0", 3, 64, 3); bn(e + "conv1.1", 64)
    conv(e + "conv1.3", 64, 64, 3); bn(e + "conv1.4", 64)
    conv(e + "conv1.6", 64, 128, 3); bn(e + "bn1", 128)
    inpl = 128
    for li, blocks in enumerate(LAYERS[arch]):
        planes = 64 * 2 ** li
        for bi in range(blocks):
            p = f"{e}layer{li + 1}.{bi}"
            conv(p + ".conv1", inpl, planes, 1); bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3); bn(p + ".bn2", planes)
            conv(p + ".conv3", planes, planes * 4, 1); bn(p + ".bn3", planes * 4, zero=True)
            if bi == 0:
                conv(p + ".downsample.0", inpl, planes * 4, 1); bn(p + ".downsample.1", planes * 4)
            inpl = planes * 4
    d = "decoder."
    conv(d + "low_conv.0", 256, 256, 1, True, False); bn(d + "low_conv.1", 256)
    conv(d + "aspp.conv1.1", 2048, 256, 1, False, False); bn(d + "aspp.conv1.2", 256)
    conv(d + "aspp.conv2.0", 2048, 256, 1, False, False); bn(d + "aspp.conv2.1", 256)
    for i in range(3):
        conv(f"{d}aspp.conv{3 + i}.0", 2048, 256, 3, False, False); bn(f"{d}aspp.conv{3 + i}.1", 256)
    conv(d + "head.0", 1280, 256, 3, False, False); bn(d + "head.1", 256)
    for head, cout in (("classifier", num_classes), ("representation", 256)):
        conv(f"{d}{head}.0", 512, 256, 3, True, False); bn(f"{d}{head}.1", 256)
        conv(f"{d}{head}.4", 256, 256, 3, True, False); bn(f"{d}{head}.5", 256)
        conv(f"{d}{head}.8", 256, cout, 1, True, False)
    if aux:
        conv("auxor.aux.0", 1024, 256, 3, True, False); bn("auxor.aux.1", 256)
        conv("auxor.aux.4", 256, num_classes, 1, True, False)
    if peak != 1.0:
        with torch.no_grad():
            s[d + "classifier.8.weight"] *= peak
    return s


def state_from_module(module):
    """Flat state (reference key names) from an nn.Module built by the reference or by the mirror."""
    params = dict(module.named_parameters())
    s = {}
    for k, v in module.state_dict().items():
        s[k] = v.detach().clone().requires_grad_(k in params)
    return s
