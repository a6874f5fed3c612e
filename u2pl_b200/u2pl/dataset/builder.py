"""`get_loader` for the drop-in `u2pl.dataset.builder` (reference builder.py:9-45).

The real VOC / Cityscapes pipelines (PIL decode, tensor-space augmentation, split lists) are outside this
project's scope: the benchmarked hot path starts at the batch.  What the drivers
need from this module is its *shape*: `get_loader(cfg, seed)` returning (sup, unsup, val) loaders for the
`*_semi` dataset types or (sup, val) otherwise, each with `len()`, `.sampler.set_epoch(e)` and iterators that
answer both `next(it)` and the Python-2 style `it.next()` the reference still calls (train_semi.py:281,285).
This implementation serves SYNTHETIC crops of the configured size (images ~ N(0,1), 8x8-blocky labels with a
10-pixel ignore border -- the generator of SURVEY.md section 8d), sharded across ranks like DistributedSampler.
Synthetic data is served ONLY on request (`dataset.synthetic: True` in the config, or U2PL_SYNTHETIC_DATA=1): any other
config -- a real `data_root`, a mistyped or missing one, none at all -- raises, so an unchanged train_semi.py started from
the wrong directory cannot silently train and validate on noise."""
import os

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler


class SyntheticCrops(Dataset):
    def __init__(self, n, size, num_classes, ignore_label=255, seed=0, with_label=True):
        self.n, self.size, self.C, self.ignore, self.seed, self.with_label = n, tuple(size), num_classes, ignore_label, seed, with_label

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + idx)
        H, W = self.size
        image = torch.randn(3, H, W, generator=g)
        blocks = torch.randint(0, self.C, ((H + 7) // 8, (W + 7) // 8), generator=g)
        label = blocks.repeat_interleave(8, 0).repeat_interleave(8, 1)[:H, :W].contiguous()
        b = min(10, H // 8)
        label[:b] = label[-b:] = self.ignore
        label[:, :b] = label[:, -b:] = self.ignore
        return image, label.long()


class _Iter:
    """DataLoader iterator that also answers `.next()`."""

    def __init__(self, it):
        self._it = it

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._it)

    next = __next__


class _Loader(DataLoader):
    def __iter__(self):
        return _Iter(super().__iter__())


def _world():
    ready = dist.is_available() and dist.is_initialized()
    return (dist.get_rank(), dist.get_world_size()) if ready else (0, 1)


def _make(cfg, split, n, seed, with_label=True):
    ds_cfg = cfg["dataset"]
    root = ds_cfg.get(split, {}).get("data_root", "")
    if not (ds_cfg.get("synthetic", False) or os.environ.get("U2PL_SYNTHETIC_DATA", "0") == "1"):
        raise NotImplementedError(
            "u2pl_b200 ships no real-data pipeline (out of scope, SURVEY.md C11) and serves synthetic crops only on "
            "request: data_root={!r} ({}).  Use the reference's u2pl.dataset package for real VOC/Cityscapes data, or set "
            "`dataset.synthetic: True` / U2PL_SYNTHETIC_DATA=1 to train on synthetic crops".format(
                root, "exists" if root and os.path.isdir(root) else "missing"))
    size = ds_cfg.get(split, {}).get("crop", {}).get("size", [513, 513])
    data = SyntheticCrops(n, size, cfg["net"]["num_classes"], ds_cfg.get("ignore_label", 255), seed, with_label)
    rank, world = _world()
    sampler = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=(split == "train"), seed=seed)
    batch = ds_cfg.get("batch_size", 1) if split == "train" else 1
    return _Loader(data, batch_size=batch, num_workers=0, sampler=sampler, shuffle=False, pin_memory=True,
                   drop_last=(split == "train"))


def get_loader(cfg, seed=0):
    kind = cfg["dataset"]["type"]
    if kind not in ("cityscapes_semi", "cityscapes", "pascal_semi", "pascal"):
        raise NotImplementedError("dataset type {} is not supported".format(cfg["dataset"]))
    n_sup = int(cfg["dataset"].get("n_sup", 64))
    rank, world = _world()
    per_epoch = max(n_sup, cfg["dataset"].get("batch_size", 1) * world)
    sup = _make(cfg, "train", per_epoch, seed)
    val = _make(cfg, "val", max(world, 8), seed + 2)
    if kind.endswith("_semi"):
        unsup = _make(cfg, "train", per_epoch, seed + 1)      # same length as `sup` (train_semi.py:258-260 asserts it)
        return sup, unsup, val
    return sup, val
