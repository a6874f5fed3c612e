"""Drop-in for u2pl/utils/lr_helper.py:12-113 (SGD/Adam factory + per-iteration poly/cosine LR)."""
from math import cos, pi

import torch.optim as optim


def get_optimizer(parms, cfg_optim):
    kind, kwargs = cfg_optim["type"], cfg_optim["kwargs"]
    optimizer = {"SGD": optim.SGD, "adam": optim.Adam}.get(kind)
    assert optimizer is not None, "optimizer type is not supported by LightSeg"
    return optimizer(parms, **kwargs)


def get_scheduler(cfg_trainer, len_data, optimizer, start_epoch=0, use_iteration=False):
    epochs = cfg_trainer["epochs"] if not use_iteration else 1
    sched = cfg_trainer["lr_scheduler"]
    return LRScheduler(sched["mode"], sched["kwargs"], len_data, optimizer, epochs, start_epoch)


class LRScheduler(object):
    def __init__(self, mode, lr_args, data_size, optimizer, num_epochs, start_epochs):
        assert mode in ["multistep", "poly", "cosine"]
        self.mode, self.optimizer, self.data_size = mode, optimizer, data_size
        self.cur_iter = start_epochs * data_size
        self.max_iter = num_epochs * data_size
        self.base_lr = [g["lr"] for g in optimizer.param_groups]
        self.cur_lr = list(self.base_lr)
        if mode == "poly":
            self.power = lr_args["power"] if lr_args.get("power", False) else 0.9
        if mode == "cosine":
            self.targetlr = lr_args["targetlr"]

    def step(self):
        frac = float(self.cur_iter) / self.max_iter
        if self.mode == "poly":
            self.cur_lr = [lr * (1 - frac) ** self.power for lr in self.base_lr]
        elif self.mode == "cosine":
            self.cur_lr = [self.targetlr + (lr - self.targetlr) * (1 + cos(pi * frac)) / 2 for lr in self.base_lr]
        else:
            raise NotImplementedError
        for group, lr in zip(self.optimizer.param_groups, self.cur_lr):
            group["lr"] = lr
        self.cur_iter += 1

    def get_lr(self):
        return self.cur_lr
