"""Drop-in for the parts of u2pl/utils/utils.py the drivers import (train_semi.py:29-38)."""
import logging
import os
import random

import numpy as np
import torch
import torch.distributed as dist


def get_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def get_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def label_onehot(inputs, num_segments):
    """Reference utils.py:50-59, including its scatter quirk: batch slot 0 receives the
    union of every image's class (ignored pixels count as class 0), slots b > 0 stay zero, and slot 0
    is cleared where image 0 is ignored.  Returns [B, C, H, W] float32."""
    from u2pl_b200 import ops
    return ops.label_onehot(inputs, num_segments)


def set_random_seed(seed, deterministic=False):
    """Reference utils.py:378-386."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    if deterministic:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False


class AverageMeter(object):
    """Windowed (length > 0) or cumulative mean -- reference utils.py:438-468."""

    def __init__(self, length=0):
        self.length = length
        self.reset()

    def reset(self):
        self.history, self.count, self.sum = [], 0, 0.0
        self.val = self.avg = 0.0

    def update(self, val, num=1):
        self.val = val
        if self.length > 0:
            assert num == 1
            self.history = (self.history + [val])[-self.length:]
            self.avg = np.mean(self.history)
        else:
            self.sum += val * num
            self.count += num
            self.avg = self.sum / self.count


_logs = set()


def init_log(name, level=logging.INFO):
    """Reference utils.py:474-491."""
    if (name, level) in _logs:
        return
    _logs.add((name, level))
    logger = logging.getLogger(name)
    logger.setLevel(level)
    ch = logging.StreamHandler()
    ch.setLevel(level)
    if "SLURM_PROCID" in os.environ:
        rank = int(os.environ["SLURM_PROCID"])
        logger.addFilter(lambda record: rank == 0)
    ch.setFormatter(logging.Formatter("[%(asctime)s][%(levelname)8s] %(message)s"))
    logger.addHandler(ch)
    return logger


def intersectionAndUnion(output, target, K, ignore_index=255):
    """Reference utils.py:568-580 (numpy histograms on the host; validation only)."""
    assert output.ndim in [1, 2, 3] and output.shape == target.shape
    output = output.reshape(output.size).copy()
    target = target.reshape(target.size)
    output[target == ignore_index] = ignore_index
    inter = output[output == target]
    bins = np.arange(K + 1)
    area_i = np.histogram(inter, bins=bins)[0]
    area_o = np.histogram(output, bins=bins)[0]
    area_t = np.histogram(target, bins=bins)[0]
    return area_i, area_o + area_t - area_i, area_t


def load_state(path, model, optimizer=None, key="state_dict"):
    """Reference utils.py:583-636: shape-mismatched keys are skipped, strict=False."""
    rank = get_rank()
    if not os.path.isfile(path):
        if rank == 0:
            print("=> no checkpoint found at '{}'".format(path))
        return None
    checkpoint = torch.load(path, map_location=lambda storage, loc: storage.cuda())
    state_dict = checkpoint[key]
    own = model.state_dict()
    for k in [k for k, v in state_dict.items() if k in own and v.shape != own[k].shape]:
        if rank == 0:
            print("caution: size-mismatch key: {} size: {} -> {}".format(k, state_dict[k].shape, own[k].shape))
        checkpoint.pop(k, None)
    model.load_state_dict(state_dict, strict=False)
    if rank == 0:
        for k in set(own.keys()) - set(state_dict.keys()):
            print("caution: missing keys from checkpoint {}: {}".format(path, k))
    if optimizer is not None:
        optimizer.load_state_dict(checkpoint["optimizer_state"])
        return checkpoint["best_miou"], checkpoint["epoch"]
