"""CPU: every name the reference drivers import (train_semi.py:19-38, train_sup.py:13-31) resolves to the drop-in
after `u2pl_b200.install()`, and the loader objects have the shape the drivers rely on."""
import torch

import u2pl_b200


def test_driver_imports_resolve():
    u2pl_b200.install()
    from u2pl.dataset.augmentation import generate_unsup_data                      # noqa: F401
    from u2pl.dataset.builder import get_loader                                    # noqa: F401
    from u2pl.models.model_helper import ModelBuilder                              # noqa: F401
    from u2pl.utils.dist_helper import setup_distributed                           # noqa: F401
    from u2pl.utils.loss_helper import (compute_contra_memobank_loss, compute_unsupervised_loss,  # noqa: F401
                                        get_criterion)
    from u2pl.utils.lr_helper import get_optimizer, get_scheduler                  # noqa: F401
    from u2pl.utils.utils import (AverageMeter, get_rank, get_world_size, init_log, intersectionAndUnion,  # noqa: F401
                                  label_onehot, load_state, set_random_seed)
    import u2pl
    assert u2pl.__file__.startswith(u2pl_b200._HERE)
    crit = get_criterion({"criterion": {"type": "CELoss", "kwargs": {"use_weight": False}}, "net": {},
                          "dataset": {"ignore_label": 255}})
    assert type(crit).__name__ == "Criterion"
    ohem = get_criterion({"criterion": {"type": "ohem", "kwargs": {"thresh": 0.7, "min_kept": 100000}},
                          "net": {"aux_loss": {"loss_weight": 0.4, "aux_plane": 1024}}, "dataset": {"ignore_label": 255}})
    assert type(ohem).__name__ == "CriterionOhem" and ohem._aux_weight == 0.4


def test_synthetic_loaders_have_the_driver_shape():
    u2pl_b200.install()
    from u2pl.dataset.builder import get_loader
    cfg = {"dataset": {"type": "pascal_semi", "batch_size": 2, "n_sup": 6, "ignore_label": 255, "synthetic": True,
                       "train": {"crop": {"size": [33, 41]}}, "val": {"crop": {"size": [33, 41]}}},
           "net": {"num_classes": 21}}
    sup, unsup, val = get_loader(cfg, seed=3)
    assert len(sup) == len(unsup) == 3
    sup.sampler.set_epoch(1)
    it = iter(sup)
    image, label = it.next()                                                       # train_semi.py:281
    assert image.shape == (2, 3, 33, 41) and label.shape == (2, 33, 41) and label.dtype == torch.int64
    assert set(label.unique().tolist()) <= set(range(21)) | {255} and (label == 255).any()
    image_u, _ = iter(unsup).next()
    assert not torch.equal(image_u, image)
    assert len(get_loader({**cfg, "dataset": {**cfg["dataset"], "type": "pascal"}})) == 2
    # synthetic crops are served on request only: a config without the flag (real, mistyped or missing data_root) raises
    import pytest
    real = {**cfg, "dataset": {**cfg["dataset"], "synthetic": False, "train": {"data_root": "../../no/such/VOC2012", "crop": {"size": [33, 41]}}}}
    with pytest.raises(NotImplementedError, match="synthetic"):
        get_loader(real)


def test_run_semi_runner_check():
    """The shipped runner for the unchanged reference drivers (python -m u2pl_b200.run_semi) resolves every name the drivers
    import to the mirror package."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "u2pl_b200.run_semi", "--check"], cwd=root, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "drop-in surface complete" in out.stdout, out.stderr[-2000:]
