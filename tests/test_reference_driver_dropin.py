"""CPU, build container only (needs /root/reference): the reference's BYTE-IDENTICAL train_semi.py runs to completion --
one supervised-only epoch, one semi-supervised epoch (teacher aliasing, pseudo labels, CutMix, unsupervised loss, the
inline contrastive preparation, compute_contra_memobank_loss, EMA), validation and checkpointing -- against the drop-in
`u2pl` package of this repository (u2pl_b200.install()), with the CUDA library emulated (tests/emulated_abi.py) and the
environment shims SURVEY.md 8(b) lists for running the driver without a GPU (stub tensorboardX, `.cuda()` -> identity,
gloo process group, DDP without device ids).  The same driver is first run on the reference's OWN package (control) with the same
synthetic loader and seeds: the recorded unsupervised / contrastive loss values, the final teacher weights and the best mIoU
of the two runs must agree.  It is the executable form of INTEGRATION.md.  Skipped where the reference is absent (the GPU box)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = "/root/reference/train_semi.py"
DRIVER_SUP = "/root/reference/train_sup.py"

CONFIG = """
dataset:
  type: pascal_semi
  synthetic: True
  train: {crop: {type: rand, size: [33, 33]}}
  val: {crop: {type: center, size: [33, 33]}}
  batch_size: 2
  n_sup: 4
  workers: 0
  ignore_label: 255
trainer:
  epochs: 2
  eval_on: True
  optimizer: {type: SGD, kwargs: {lr: 0.001, momentum: 0.9, weight_decay: 0.0001}}
  lr_scheduler: {mode: poly, kwargs: {power: 0.9}}
  unsupervised: {TTA: False, drop_percent: 80, apply_aug: cutmix}
  contrastive: {negative_high_entropy: True, low_rank: 3, high_rank: 20, current_class_threshold: 0.3,
                current_class_negative_threshold: 1, unsupervised_entropy_ignore: 80, low_entropy_threshold: 20,
                num_negatives: 50, num_queries: 256, temperature: 0.5}
saver: {snapshot_dir: checkpoints, pretrain: ''}
criterion: {type: CELoss, kwargs: {use_weight: False}}
net:
  num_classes: 21
  sync_bn: False
  ema_decay: 0.99
  encoder:
    type: u2pl.models.resnet.resnet50
    kwargs: {multi_grid: True, zero_init_residual: True, fpn: True, replace_stride_with_dilation: [False, True, True],
             pretrained: False}
  decoder:
    type: u2pl.models.decoder.dec_deeplabv3_plus
    kwargs: {inner_planes: 256, dilations: [12, 24, 36]}
"""

RUNNER = """
import os, runpy, sys, types
sys.path[:0] = [{root!r}, {tests!r}]
import torch, torch.distributed as dist
# ---- environment shims (no GPU here): SURVEY.md 8(b)
tb = types.ModuleType("tensorboardX")
tb.SummaryWriter = type("SummaryWriter", (), {{"__init__": lambda self, *a, **k: None, "add_scalar": lambda self, *a, **k: None}})
sys.modules["tensorboardX"] = tb
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.device_count = lambda: 1
_init = dist.init_process_group
dist.init_process_group = lambda backend=None, **k: _init(backend="gloo", **k)
_DDP = torch.nn.parallel.DistributedDataParallel
class DDP(_DDP):
    def __init__(self, module, device_ids=None, output_device=None, **k):
        super().__init__(module, **k)
torch.nn.parallel.DistributedDataParallel = DDP
os.environ.update(RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT={port!r})
import u2pl_b200
if {dropin!r} == "1":                                       # ---- the drop-in, over the emulated C ABI
    import emulated_abi
    emulated_abi.install(emulated_abi.DirectPatcher())
    u2pl_b200.install()
else:                                                       # ---- the reference's own package (control run)
    sk = types.ModuleType("skimage"); sk.measure = types.ModuleType("skimage.measure"); sk.measure.label = sk.measure.regionprops = None
    sys.modules["skimage"], sys.modules["skimage.measure"] = sk, sk.measure
    sys.path.insert(0, "/root/reference")
    import importlib.util, u2pl.dataset.builder as RB
    spec = importlib.util.spec_from_file_location("synthetic_builder", os.path.join(u2pl_b200._HERE, "u2pl", "dataset", "builder.py"))
    sb = importlib.util.module_from_spec(spec); spec.loader.exec_module(sb)
    RB.get_loader = sb.get_loader                           # same synthetic crops in both runs
    _next = torch.utils.data.dataloader._BaseDataLoaderIter
    _next.next = _next.__next__
import u2pl.models.model_helper as MH                       # sharper initial predictions in BOTH runs, so that the contrastive
_mb_init = MH.ModelBuilder.__init__                         # branch finds anchors (prob > 0.3) with an untrained network
def _peaked(self, net_cfg):
    _mb_init(self, net_cfg)
    with torch.no_grad():
        self.decoder.classifier[-1].weight.mul_(8.0)
MH.ModelBuilder.__init__ = _peaked
calls = {{"unsup": [], "contra": [], "sup": []}}
import u2pl.utils.loss_helper as LH
_gc = LH.get_criterion
def _get_criterion(cfg):                                    # record what the supervised criterion returns (train_sup.py:218-220)
    crit = _gc(cfg)
    fwd = crit.forward
    def rec(*a, **kw):
        out = fwd(*a, **kw)
        calls["sup"].append(float(out.detach()))
        return out
    crit.forward = rec
    return crit
LH.get_criterion = _get_criterion
def _rec(f, k):
    def g(*a, **kw):
        out = f(*a, **kw)
        calls[k].append(float((out[-1] if isinstance(out, tuple) else out).detach()))
        return out
    return g
for name, key in (("compute_unsupervised_loss", "unsup"), ("compute_contra_memobank_loss", "contra")):
    setattr(LH, name, _rec(getattr(LH, name), key))
sys.argv = [os.path.basename({driver!r}), "--config", {config!r}, "--seed", "2", "--port", {port!r}]
os.chdir(os.path.dirname({config!r}))
runpy.run_path({driver!r}, run_name="__main__")
import u2pl, json
assert u2pl.__file__.startswith(u2pl_b200._HERE) == ({dropin!r} == "1"), u2pl.__file__
ck = torch.load(os.path.join(os.path.dirname({config!r}), "checkpoints", "ckpt.pth"), map_location="cpu", weights_only=False)
w = ck["teacher_state" if "teacher_state" in ck else "model_state"]["module.decoder.classifier.8.weight"]
print("DRIVER_DONE", json.dumps(dict(calls=calls, teacher_sum=float(w.double().abs().sum()), best=float(ck.get("best_miou", -1)))))
"""


def _run_both(tmp_path, driver, config_text):
    import json
    res = {}
    for dropin in ("0", "1"):                                              # control run on the reference's package, then the drop-in
        work = tmp_path / ("dropin" if dropin == "1" else "reference")
        work.mkdir()
        cfg = work / "config.yaml"
        cfg.write_text(textwrap.dedent(config_text))
        script = RUNNER.format(root=ROOT, tests=os.path.join(ROOT, "tests"), config=str(cfg), driver=driver, dropin=dropin,
                               port=str(36000 + (os.getpid() + int(dropin)) % 2000))
        out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=900, cwd=str(work))
        tail = (out.stdout + out.stderr)[-3000:]
        assert out.returncode == 0 and "DRIVER_DONE" in out.stdout, tail
        res[dropin] = json.loads(out.stdout[out.stdout.index("DRIVER_DONE") + len("DRIVER_DONE"):].strip().splitlines()[0])
        assert (work / "checkpoints" / "ckpt.pth").exists()
    return res["0"], res["1"]


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="the reference is only mounted in the build container")
def test_unchanged_reference_driver_runs_on_the_dropin(tmp_path):
    a, b = _run_both(tmp_path, DRIVER, CONFIG)
    assert len(b["calls"]["unsup"]) == len(b["calls"]["contra"]) == 2      # the semi-supervised epoch ran both drop-in losses twice
    for k in ("unsup", "contra"):                                          # same driver, two packages: the same loss trajectory
        for x, y in zip(a["calls"][k], b["calls"][k]):
            assert abs(x - y) <= 1e-3 * max(1.0, abs(x)), (k, a["calls"], b["calls"])   # fp32 noise, amplified by one SGD step
    assert max(a["calls"]["contra"]) > 0 and max(b["calls"]["contra"]) > 0   # the contrastive loss was really non-trivial
    assert abs(a["teacher_sum"] - b["teacher_sum"]) <= 1e-4 * a["teacher_sum"]
    assert abs(a["best"] - b["best"]) <= 2e-3                                  # mIoU: a few arg-max flips on near-tie pixels


@pytest.mark.skipif(not os.path.exists(DRIVER_SUP), reason="the reference is only mounted in the build container")
def test_unchanged_train_sup_runs_on_the_dropin(tmp_path):
    """BASELINE.json configs[0]: train_sup.py SupOnly, ResNet50-DeepLabv3+, two synthetic VOC crops, CPU -- the unchanged
    driver on the reference's package (control) and on the drop-in: same supervised losses, weights and mIoU."""
    cfg = CONFIG.replace("type: pascal_semi", "type: pascal").replace("n_sup: 4", "n_sup: 6") \
        .replace("kwargs: {inner_planes: 256, dilations: [12, 24, 36]}",
                 "kwargs: {inner_planes: 256, dilations: [12, 24, 36], rep_head: False}")      # as experiments/*/suponly/config.yaml
    a, b = _run_both(tmp_path, DRIVER_SUP, cfg)
    assert len(a["calls"]["sup"]) == len(b["calls"]["sup"]) == 6           # 3 iterations x 2 epochs
    for x, y in zip(a["calls"]["sup"], b["calls"]["sup"]):
        assert abs(x - y) <= 1e-3 * max(1.0, abs(x)), (a["calls"]["sup"], b["calls"]["sup"])
    assert abs(a["teacher_sum"] - b["teacher_sum"]) <= 1e-4 * a["teacher_sum"] and abs(a["best"] - b["best"]) <= 2e-3
