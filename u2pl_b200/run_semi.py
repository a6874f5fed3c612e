"""python -m u2pl_b200.run_semi /path/to/U2PL/train_semi.py --config=cfg.yaml [--seed N --port P ...]

Runs the reference's UNCHANGED driver (train_semi.py or train_sup.py, byte-identical) on the drop-in package: it registers
the `u2pl` mirror (u2pl_b200.install()), supplies the three things the driver needs around it on a current PyTorch image
 and then executes the file with runpy -- nothing in the driver is edited or monkey-patched:

  * `tensorboardX.SummaryWriter`: used for logging only (train_semi.py:16,151); a no-op stand-in is registered when the
    package is not installed;
  * `LOCAL_RANK` / `RANK` / `WORLD_SIZE` / `MASTER_ADDR` / `MASTER_PORT`: defaulted for a single-process run (the driver
    reads os.environ["LOCAL_RANK"], train_semi.py:114); under torchrun they are already set;
  * `pretrained: True` needs the ImageNet weights the reference downloads; keep `pretrained: False` or a local path.

The parts of the loop body that are inline in the driver (second entropy / percentile block train_semi.py:397-465, EMA
:531-548) stay torch-eager + numpy on this route; `u2pl_b200.step.SemiStep` is the fused restatement of the whole body
(what bench.py times).  tests/test_reference_driver_dropin.py is the executable check of this route (CPU, emulated ABI,
against the reference's own package)."""
import os
import runpy
import sys
import types


def _shims():
    try:
        import tensorboardX  # noqa: F401
    except Exception:
        tb = types.ModuleType("tensorboardX")
        tb.SummaryWriter = type("SummaryWriter", (), {"__init__": lambda self, *a, **k: None,
                                                      "add_scalar": lambda self, *a, **k: None,
                                                      "close": lambda self, *a, **k: None})
        sys.modules["tensorboardX"] = tb
    for k, v in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0"), ("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", "29533")):
        os.environ.setdefault(k, v)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0
    check_only = argv[0] == "--check"
    if check_only:
        argv = argv[1:]
    import u2pl_b200
    u2pl_b200.install()
    _shims()
    if check_only:
        # every name the drivers import (train_semi.py:19-38, train_sup.py:16-30) resolves to the mirror
        import importlib
        names = {"u2pl.dataset.augmentation": ["generate_unsup_data"], "u2pl.dataset.builder": ["get_loader"],
                 "u2pl.models.model_helper": ["ModelBuilder"], "u2pl.utils.dist_helper": ["setup_distributed"],
                 "u2pl.utils.loss_helper": ["compute_contra_memobank_loss", "compute_unsupervised_loss", "get_criterion"],
                 "u2pl.utils.lr_helper": ["get_optimizer", "get_scheduler"],
                 "u2pl.utils.utils": ["AverageMeter", "get_rank", "get_world_size", "init_log", "intersectionAndUnion",
                                      "label_onehot", "load_state", "set_random_seed"]}
        for mod, attrs in names.items():
            m = importlib.import_module(mod)
            assert m.__file__.startswith(u2pl_b200._HERE), (mod, m.__file__)
            for a in attrs:
                assert hasattr(m, a), (mod, a)
        print("u2pl_b200.run_semi: drop-in surface complete")
        return 0
    driver = os.path.abspath(argv[0])
    sys.argv = [driver] + argv[1:]
    runpy.run_path(driver, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
