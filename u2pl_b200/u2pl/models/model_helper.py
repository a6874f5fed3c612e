"""`ModelBuilder`: the reference's plugin API (u2pl/models/model_helper.py:9-66) -- encoder and decoder
classes are named by dotted strings in the config and resolved with importlib; `forward(x)` returns
{"pred", "rep"[, "aux"]}.  Attribute names (`encoder`, `decoder`, `auxor`) are part of the contract:
train_semi.py:82-110 builds optimiser parameter groups from them."""
from importlib import import_module

import torch.nn as nn

from .decoder import Aux_Module


def _instantiate(dotted, kwargs):
    module_path, _, attr = dotted.rpartition(".")
    return getattr(import_module(module_path), attr)(**kwargs)


class ModelBuilder(nn.Module):
    def __init__(self, net_cfg):
        super().__init__()
        sync_bn, classes = net_cfg["sync_bn"], net_cfg["num_classes"]
        self._sync_bn, self._num_classes = sync_bn, classes

        enc = net_cfg["encoder"]
        enc["kwargs"]["sync_bn"] = sync_bn                      # the reference mutates the config dict in place too
        self.encoder = _instantiate(enc["type"], enc["kwargs"])

        dec = net_cfg["decoder"]
        dec["kwargs"].update(in_planes=self.encoder.get_outplanes(), sync_bn=sync_bn, num_classes=classes)
        self.decoder = _instantiate(dec["type"], dec["kwargs"])

        aux_cfg = net_cfg.get("aux_loss", False)
        self._use_auxloss = bool(aux_cfg)
        self.fpn = bool(enc["kwargs"].get("fpn", False))
        if aux_cfg:
            self.loss_weight = aux_cfg["loss_weight"]
            self.auxor = Aux_Module(aux_cfg["aux_plane"], classes, sync_bn)

    def forward(self, x):
        feats = self.encoder(x)
        if not self._use_auxloss:
            return self.decoder(feats)
        # with the aux head the decoder sees either the 4-level pyramid (fpn) or only the last stage
        outs = self.decoder(list(feats) if self.fpn else feats[1])
        outs["aux"] = self.auxor(feats[2] if self.fpn else feats[0])
        return outs
