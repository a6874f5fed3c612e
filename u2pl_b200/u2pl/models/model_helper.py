"""ModelBuilder: encoder / decoder / aux head resolved from dotted type strings (the reference's
plugin API, u2pl/models/model_helper.py:9-66).  forward(x) -> {"pred", "rep"[, "aux"]}."""
import importlib

import torch.nn as nn

from .decoder import Aux_Module


class ModelBuilder(nn.Module):
    def __init__(self, net_cfg):
        super().__init__()
        self._sync_bn = net_cfg["sync_bn"]
        self._num_classes = net_cfg["num_classes"]
        self.encoder = self._build_encoder(net_cfg["encoder"])
        self.decoder = self._build_decoder(net_cfg["decoder"])
        self._use_auxloss = bool(net_cfg.get("aux_loss", False))
        self.fpn = bool(net_cfg["encoder"]["kwargs"].get("fpn", False))
        if self._use_auxloss:
            cfg_aux = net_cfg["aux_loss"]
            self.loss_weight = cfg_aux["loss_weight"]
            self.auxor = Aux_Module(cfg_aux["aux_plane"], self._num_classes, self._sync_bn)

    def _build_encoder(self, enc_cfg):
        enc_cfg["kwargs"].update({"sync_bn": self._sync_bn})
        return self._build_module(enc_cfg["type"], enc_cfg["kwargs"])

    def _build_decoder(self, dec_cfg):
        dec_cfg["kwargs"].update({"in_planes": self.encoder.get_outplanes(), "sync_bn": self._sync_bn,
                                  "num_classes": self._num_classes})
        return self._build_module(dec_cfg["type"], dec_cfg["kwargs"])

    @staticmethod
    def _build_module(mtype, kwargs):
        module_name, class_name = mtype.rsplit(".", 1)
        return getattr(importlib.import_module(module_name), class_name)(**kwargs)

    def forward(self, x):
        if not self._use_auxloss:
            return self.decoder(self.encoder(x))
        if self.fpn:
            f1, f2, feat1, feat2 = self.encoder(x)
            outs = self.decoder([f1, f2, feat1, feat2])
        else:
            feat1, feat2 = self.encoder(x)
            outs = self.decoder(feat2)
        outs.update({"aux": self.auxor(feat1)})
        return outs
