"""The helpers of `u2pl.utils.utils` that the drivers import (train_semi.py:29-38, train_sup.py): same names and
observable behaviour as the reference (u2pl/utils/utils.py), host-side bookkeeping only -- except `label_onehot`,
which runs on the device and reproduces the reference's scatter quirk."""
import logging
import os
import random

import numpy as np
import torch
import torch.distributed as dist


def _dist_ready():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if _dist_ready() else 1


def get_rank():
    return dist.get_rank() if _dist_ready() else 0


def label_onehot(inputs, num_segments):
    """[B,H,W] int64 -> [B,C,H,W] float32 as the reference computes it (utils.py:50-59): batch slot 0 receives the
    union of every image's class (ignored pixels count as class 0) and is cleared where image 0 is ignored;
    slots b > 0 stay zero."""
    from u2pl_b200 import ops
    return ops.label_onehot(inputs, num_segments)


def set_random_seed(seed, deterministic=False):
    for seeder in (random.seed, np.random.seed, torch.manual_seed, torch.cuda.manual_seed_all):
        seeder(seed)
    if deterministic:
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False


class AverageMeter(object):
    """`length > 0`: mean over the last `length` values; otherwise a running (weighted) mean.  Exposes .val / .avg."""

    def __init__(self, length=0):
        self.length = length
        self.reset()

    def reset(self):
        self.history, self.count, self.sum, self.val, self.avg = [], 0, 0.0, 0.0, 0.0

    def update(self, val, num=1):
        self.val = val
        if self.length > 0:
            assert num == 1, "windowed meters take one value at a time"
            self.history.append(val)
            del self.history[:-self.length]
            self.avg = float(np.mean(self.history))
        else:
            self.sum, self.count = self.sum + val * num, self.count + num
            self.avg = self.sum / self.count


_configured_loggers = set()


def init_log(name, level=logging.INFO):
    """Console logger, configured once per (name, level); under SLURM only rank 0 emits records."""
    key = (name, level)
    if key in _configured_loggers:
        return None
    _configured_loggers.add(key)
    logger = logging.getLogger(name)
    logger.setLevel(level)
    if "SLURM_PROCID" in os.environ:
        procid = int(os.environ["SLURM_PROCID"])
        logger.addFilter(lambda record: procid == 0)
    handler = logging.StreamHandler()
    handler.setLevel(level)
    handler.setFormatter(logging.Formatter("[%(asctime)s][%(levelname)8s] %(message)s"))
    logger.addHandler(handler)
    return logger


def intersectionAndUnion(output, target, K, ignore_index=255):
    """Per-class intersection / union / target pixel counts (numpy, validation loop of train_semi.py:625-632)."""
    assert output.ndim in (1, 2, 3) and output.shape == target.shape
    pred = output.reshape(-1).copy()
    gt = target.reshape(-1)
    pred[gt == ignore_index] = ignore_index
    edges = np.arange(K + 1)
    hit = np.histogram(pred[pred == gt], bins=edges)[0]
    n_pred, n_gt = np.histogram(pred, bins=edges)[0], np.histogram(gt, bins=edges)[0]
    return hit, n_pred + n_gt - hit, n_gt


def load_state(path, model, optimizer=None, key="state_dict"):
    """Load `checkpoint[key]` non-strictly (shape-mismatched entries are reported and skipped by strict=False);
    with an optimizer also restore it and return (best_miou, epoch) like the reference (utils.py:583-636)."""
    verbose = get_rank() == 0
    if not os.path.isfile(path):
        if verbose:
            print(f"=> no checkpoint found at '{path}'")
        return None
    if verbose:
        print(f"=> loading checkpoint '{path}'")
    checkpoint = torch.load(path, map_location=lambda storage, location: storage.cuda())
    state, own = checkpoint[key], model.state_dict()
    for name, value in state.items():
        if name in own and value.shape != own[name].shape:
            checkpoint.pop(name, None)
            if verbose:
                print(f"caution: size-mismatch key: {name} size: {value.shape} -> {own[name].shape}")
    model.load_state_dict(state, strict=False)
    if verbose:
        for name in sorted(set(own) - set(state)):
            print(f"caution: missing keys from checkpoint {path}: {name}")
    if optimizer is None:
        return None
    optimizer.load_state_dict(checkpoint["optimizer_state"])
    if verbose:
        print(f"=> also loaded optimizer from checkpoint '{path}' (epoch {checkpoint['epoch']})")
    return checkpoint["best_miou"], checkpoint["epoch"]
