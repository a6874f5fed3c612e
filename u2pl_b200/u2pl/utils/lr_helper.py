"""Optimiser factory and the per-iteration LR schedule for the drop-in `u2pl.utils.lr_helper`
(reference :12-113).  The schedule is stepped once per training iteration BEFORE the batch
(train_semi.py:277-279) and rewrites `param_group["lr"]` of the optimiser it was given."""
import math

import torch.optim as optim

_OPTIMIZERS = {"SGD": optim.SGD, "adam": optim.Adam}


def get_optimizer(parms, cfg_optim):
    cls = _OPTIMIZERS.get(cfg_optim["type"])
    assert cls is not None, "optimizer type is not supported by LightSeg"
    return cls(parms, **cfg_optim["kwargs"])


def get_scheduler(cfg_trainer, len_data, optimizer, start_epoch=0, use_iteration=False):
    spec = cfg_trainer["lr_scheduler"]
    return LRScheduler(spec["mode"], spec["kwargs"], len_data, optimizer, 1 if use_iteration else cfg_trainer["epochs"],
                       start_epoch)


class LRScheduler(object):
    """lr_i(t) = base_i * (1 - t/T)^power            (poly)
               = target + (base_i - target) * (1 + cos(pi t/T)) / 2   (cosine),   t = iterations done, T = total."""

    def __init__(self, mode, lr_args, data_size, optimizer, num_epochs, start_epochs):
        assert mode in ["multistep", "poly", "cosine"]
        self.mode, self.optimizer, self.data_size = mode, optimizer, data_size
        self.cur_iter, self.max_iter = start_epochs * data_size, num_epochs * data_size
        self.base_lr = [group["lr"] for group in optimizer.param_groups]
        self.cur_lr = list(self.base_lr)
        if mode == "poly":
            self.power = lr_args.get("power") or 0.9
        elif mode == "cosine":
            self.targetlr = lr_args["targetlr"]

    def _factor(self, lr):
        t = float(self.cur_iter) / self.max_iter
        if self.mode == "poly":
            return lr * (1 - t) ** self.power
        if self.mode == "cosine":
            return self.targetlr + (lr - self.targetlr) * (1 + math.cos(math.pi * t)) / 2
        raise NotImplementedError(self.mode)

    def step(self):
        self.cur_lr = [self._factor(lr) for lr in self.base_lr]
        for group, lr in zip(self.optimizer.param_groups, self.cur_lr):
            group["lr"] = lr
        self.cur_iter += 1

    def get_lr(self):
        return self.cur_lr
