#!/usr/bin/env python
"""bench.py -- train-step images/sec of the U2PL semi-supervised step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|eager] [--workload v16|c2|tiny]

Workload at N=1: BASELINE.json configs[1] -- train_semi.py U2PL VOC, ResNet101-DeepLabv3+,
513x513 crops, C=21, batch 16 labelled + 16 unlabelled, mid-training epoch (40/80) so that
drop_percent = 90 and alpha_t = 10, memory banks pre-filled to capacity.  N>1: the same per-GPU
batch on every rank (weak scaling), DDP + SyncBN over NCCL, one rank per GPU (torchrun).

A "step" is train_semi.py:272-561 for one batch: teacher eval forward, strong augmentation,
student forward, supervised loss, teacher train forward, fused entropy / percentile / partition,
masked CE, contrastive memory-bank loss, backward, SGD, EMA.

`--impl reference` times the CPU restatement of the same step (oracle/step_port.py: plain torch
fp32 network + oracle losses) on the host cores, on a bounded sample (1 labelled + 1 unlabelled
crop per step).  The reference itself is Python and cannot travel to the GPU box (/root/reference
does not exist there); the oracle is pinned to it by tests/golden.

`--impl eager` (one GPU) times oracle/eager_step.py: the same step with the reference's own device placement
(torch-eager fp32 network and per-pixel math on the GPU, numpy percentiles and memory banks on the host) --
the "reference PyTorch-eager step" BASELINE.json's north_star compares against.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "train-step images/sec (513^2 ResNet101-DeepLabv3+)"
FWD_GFLOP_PER_IMG = {513: 651.1, 769: 1496.1, 129: 651.1 * (129 / 513) ** 2}      # SURVEY.md 8d [FlopCounterMode]

WORKLOADS = {
    # name: (arch, C, crop, per-GPU labelled batch, per-GPU unlabelled batch, criterion, aux)
    "v16": ("resnet101", 21, 513, 16, 16, "CELoss", False),
    # BASELINE configs[2]: experiments/cityscapes/744/ours/config.yaml (OHEM thresh 0.7 / min_kept 100000, aux head 0.4, SyncBN)
    "c2": ("resnet101", 19, 769, 2, 2, "ohem", True),
    "tiny": ("resnet50", 21, 129, 2, 2, "CELoss", False),
}


def make_cfg(workload):
    arch, C, crop, bl, bu, crit, aux = WORKLOADS[workload]
    cfg = {
        "dataset": {"ignore_label": 255, "type": "pascal_semi" if C == 21 else "cityscapes_semi"},
        "trainer": {"epochs": 80, "sup_only_epoch": 0,
                    "optimizer": {"type": "SGD", "kwargs": {"lr": 0.001, "momentum": 0.9, "weight_decay": 0.0001}},
                    "unsupervised": {"drop_percent": 80, "apply_aug": "cutmix"},
                    "contrastive": {"negative_high_entropy": True, "low_rank": 3, "high_rank": 20,
                                    "current_class_threshold": 0.3, "current_class_negative_threshold": 1,
                                    "low_entropy_threshold": 20, "num_negatives": 50, "num_queries": 256,
                                    "temperature": 0.5}},
        "criterion": ({"type": "ohem", "kwargs": {"thresh": 0.7, "min_kept": 100000}} if crit == "ohem"
                      else {"type": crit, "kwargs": {"use_weight": False}}),
        "net": {"num_classes": C, "sync_bn": False, "ema_decay": 0.99,
                "encoder": {"type": f"u2pl.models.resnet.{arch}",
                            "kwargs": {"multi_grid": True, "zero_init_residual": True, "fpn": True,
                                       "replace_stride_with_dilation": [False, True, True], "pretrained": False}},
                "decoder": {"type": "u2pl.models.decoder.dec_deeplabv3_plus",
                            "kwargs": {"inner_planes": 256, "dilations": [12, 24, 36]}}},
    }
    if aux:
        cfg["net"]["aux_loss"] = {"aux_plane": 1024, "loss_weight": 0.4}
    if C == 19:                                                # cityscapes/744/ours/config.yaml:31-38
        cfg["trainer"]["epochs"] = 200
        cfg["trainer"]["optimizer"]["kwargs"].update(lr=0.01, weight_decay=0.0005)
    return cfg


def synth_batch(seed, bl, bu, crop, C):
    """SURVEY.md 8d synthetic inputs: images ~N(0,1); labels 8x8-blocky uniform over classes with a
    10-pixel border of 255."""
    g = torch.Generator().manual_seed(seed)
    image_l = torch.randn(bl, 3, crop, crop, generator=g)
    image_u = torch.randn(bu, 3, crop, crop, generator=g)
    nb = (crop + 7) // 8
    lab = torch.randint(0, C, (bl, nb, nb), generator=g)
    lab = lab.repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :crop, :crop].contiguous()
    lab[:, :10] = 255
    lab[:, -10:] = 255
    lab[:, :, :10] = 255
    lab[:, :, -10:] = 255
    return image_l, lab, image_u


NETWORK_NOTE = ("channels-last bf16 autocast.  Own kernels (libu2pl_b200.so): flat-tile tcgen05 implicit-GEMM convolution "
                "(im2col TMA, BN + bias + residual + ReLU in the epilogue) for every stride-1 convolution of the teacher's "
                "eval forward (T1: 106 of 116 convolutions, 97 % of its FLOPs) and for the dilation >= 18 forwards of all "
                "passes; tcgen05 weight gradient (MN-major operands) for the dilated 3x3 layers; BN statistics / apply "
                "(+ReLU, +residual) / backward; stem max-pool; every loss kernel; SGD + EMA.  Library (cuDNN / cuBLAS): "
                "train-mode forward of the other convolutions, data gradients, remaining weight gradients -- see "
                "config.routing (U2PL_TC_TRAIN=1 moves those onto the own kernels too, measured 5 % slower per step)")
EPOCH, LEN_LOADER = 40, 100                 # V16 mid-training (epoch 40/80): drop_percent 90, alpha_t 10
                                            # (c2 has 200 epochs: epoch 40 -> drop_percent 84, alpha_t 16)
PEAK = 8.0                                  # scale of the last classifier conv so that random-init teacher
                                            # probabilities are peaked enough to produce anchors / keys


# =========================================================================== clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
                for n, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =========================================================================== reference arm (CPU)
def usable_cores():
    """Host threads this process may really use: min(affinity, cgroup CPU quota).  (On the GPU box
    nproc says 128 but the container's cpu.max is 16 CPUs; 128 torch threads there run 60x slower.)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def reference_arm(args, quiet=False):
    from oracle import model_port, step_port
    arch, C, crop, _, _, _, aux = WORKLOADS[args.workload]
    cores = usable_cores()
    torch.set_num_threads(cores)
    cfg = make_cfg(args.workload)
    s_state = model_port.init_state(arch, C, aux, seed=1, peak=PEAK)
    t_state = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in s_state.items()}
    ref = step_port.ReferenceStep(s_state, t_state, cfg, arch)
    g = np.random.default_rng(0)
    for c in range(C):                                        # banks pre-filled (reduced: CPU sample)
        ref.memobank[c][0] = g.standard_normal((2000, 256)).astype(np.float32)
    bl = bu = 1
    image_l, label_l, image_u = synth_batch(1234, bl, bu, crop, C)
    np.random.seed(1234)
    torch.manual_seed(1234)
    warm, steps = (1, 2) if quiet else (max(1, min(args.warmup, 1)), max(1, min(args.steps, 3)))
    for i in range(warm):
        ref.step(image_l, label_l, image_u, EPOCH, EPOCH * LEN_LOADER + i, LEN_LOADER)
    t0 = time.perf_counter()
    for i in range(steps):
        losses = ref.step(image_l, label_l, image_u, EPOCH, EPOCH * LEN_LOADER + warm + i, LEN_LOADER)
    dt = (time.perf_counter() - t0) / steps
    value = (bl + bu) / dt
    sample = (f"{bl}+{bu} crops of {crop}x{crop} per step ({arch}, C={C}), {warm} warm-up + {steps} timed steps, "
              f"banks 2000 rows/class, torch fp32 on {cores} threads")
    base = {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample,
            "ms_per_step": dt * 1e3, "losses": [float(x) for x in losses], "steps_run": steps, "warmup_run": warm,
            "batch": bl + bu, "bank_rows_per_class": 2000}
    return base


def print_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = reference_arm(args)
    arch, C, crop, bl, bu, crit, aux = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": base["steps_run"], "warmup": base["warmup_run"], "steps_requested": args.steps,
            "warmup_requested": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"train_semi.py U2PL step, {arch}-DeepLabv3+ {crop}x{crop} C={C} (CPU sample)",
                       "same_config": False, "global_batch": base["batch"], "bank_rows_per_class": base["bank_rows_per_class"],
                       "normalisation": "images/s is per-image throughput of a bounded sample (1+1 crops per step, 2000-row banks); "
                                        "the GPU arm runs 16+16 crops per GPU and 30k/50k-row banks -- a per-image ratio, not a "
                                        "same-config speed-up",
                       "note": "the reference is Python and cannot travel; this is oracle/step_port.py, pinned to it"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# =========================================================================== torch-eager comparator (GPU)
def eager_arm(args, dev=None, quiet=False):
    """the reference step as the reference runs it on a GPU -- torch-eager fp32 (cuDNN TF32
    allowed, torch's default), host percentiles, CPU banks, per-class Python loops -- via oracle/eager_step.py.
    Single GPU only (rank 0); halves the batch on a CUDA OOM and says so in `config`."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import eager_step, model_port
    arch, C, crop, bl, bu, _, aux = WORKLOADS[args.workload]
    if dev is None:                                           # `dev="cpu"` is for the plumbing test only
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl eager needs a CUDA device (use --impl reference for the CPU arm)")
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        torch.cuda.set_device(dev)
        torch.backends.cudnn.benchmark = True
    dev = torch.device(dev)
    on_gpu = dev.type == "cuda"
    cfg = make_cfg(args.workload)
    base = model_port.init_state(arch, C, aux, seed=1, peak=PEAK)

    def on_dev():
        return {k: v.detach().to(dev).requires_grad_(v.requires_grad) for k, v in base.items()}

    while True:
        if on_gpu:
            torch.cuda.empty_cache()                          # (after an OOM retry: the failed attempt's frames are gone here)
        try:
            ref = eager_step.EagerStep(on_dev(), on_dev(), cfg, arch)
            g = torch.Generator().manual_seed(7)
            for c in range(C):                                # banks pre-filled to capacity, on the host as in the reference
                ref.memobank[c][0] = torch.randn(ref.queue_size[c], 256, generator=g)
            image_l, label_l, image_u = synth_batch(1234, bl, bu, crop, C)
            np.random.seed(1234)
            torch.manual_seed(1234)
            it = EPOCH * LEN_LOADER
            for _ in range(max(args.warmup, 3) if on_gpu else args.warmup):
                ref.step(image_l, label_l, image_u, EPOCH, it, LEN_LOADER)
                it += 1
            clocks = ClockSampler(dev.index or 0)
            if on_gpu:
                clocks.start()
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                losses = ref.step(image_l, label_l, image_u, EPOCH, it, LEN_LOADER)
                it += 1
            if on_gpu:
                e.record()
                torch.cuda.synchronize()
            wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
            ms = s.elapsed_time(e) / args.steps if on_gpu else wall_ms
            clk = clocks.stop() if on_gpu else None
            break
        except torch.cuda.OutOfMemoryError:
            if bl == 1:
                raise
            ref = None
            bl, bu = bl // 2, bu // 2
    value = (bl + bu) / (ms * 1e-3)
    h2d = image_l.numel() * 4 + label_l.numel() * 8 + image_u.numel() * 4
    line = {
        "impl": "eager", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "wall_ms_per_step": wall_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (cuDNN TF32 allowed: torch default)", "data": "synthetic",
        "config": {"workload": f"train_semi.py U2PL step as the reference executes it (oracle/eager_step.py), {arch}-DeepLabv3+ "
                               f"{crop}x{crop} C={C}, {bl}+{bu} crops, epoch {EPOCH}/80, CPU banks full (30k/50k x 256)",
                   "global_batch": bl + bu, "requested_batch": WORKLOADS[args.workload][3] * 2, "l2": "inputs_exceed_L2"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                "note": "host inputs are copied inside every step (train_semi.py:283,287), so value == e2e"},
        "gpu_launches": 0, "clocks": clk, "losses": [float(x) for x in losses]}
    if not quiet:
        print(json.dumps(line))
    return line


# =========================================================================== our arm (GPU)
def our_arm(args):
    import torch.distributed as dist
    import u2pl_b200
    u2pl_b200.install()
    from u2pl.models.model_helper import ModelBuilder
    from u2pl.utils.loss_helper import get_criterion
    from u2pl.utils.lr_helper import get_optimizer
    from u2pl_b200 import _lib
    from u2pl_b200.step import SemiStep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    _lib.load(build_if_missing=False)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"
    torch.backends.cudnn.benchmark = True

    arch, C, crop, bl, bu, crit, aux = WORKLOADS[args.workload]
    cfg = make_cfg(args.workload)
    cfg["net"]["sync_bn"] = world > 1
    torch.manual_seed(1)
    model = ModelBuilder(cfg["net"])
    with torch.no_grad():
        model.decoder.classifier[-1].weight.mul_(PEAK)
    teacher = ModelBuilder(cfg["net"])
    teacher.load_state_dict(model.state_dict())
    model.cuda().to(memory_format=torch.channels_last)
    teacher.cuda().to(memory_format=torch.channels_last)
    times = 10 if "pascal" in cfg["dataset"]["type"] else 1
    lr = cfg["trainer"]["optimizer"]["kwargs"]["lr"]
    params = [dict(params=model.encoder.parameters(), lr=lr)]        # train_semi.py:82-110: backbone, then heads at lr x times
    for head in ([model.auxor, model.decoder] if aux else [model.decoder]):
        params.append(dict(params=head.parameters(), lr=lr * times))
    optimizer = get_optimizer(params, cfg["trainer"]["optimizer"])
    if world > 1:                                             # same order as train_semi.py:114-133
        ddp = torch.nn.parallel.DistributedDataParallel
        model = ddp(model, device_ids=[local_rank], output_device=local_rank, find_unused_parameters=False)
        teacher = ddp(teacher, device_ids=[local_rank], output_device=local_rank, find_unused_parameters=False)
    for p in teacher.parameters():
        p.requires_grad = False
    sup_loss_fn = get_criterion(cfg)
    memobank, queue_ptrlis, queue_size = [], [], []
    g = torch.Generator().manual_seed(7)
    for c in range(C):                                        # train_semi.py:161-169, pre-filled to capacity
        queue_size.append(50000 if c == 0 else 30000)
        memobank.append([torch.randn(queue_size[-1], 256, generator=g)])
        queue_ptrlis.append(torch.zeros(1, dtype=torch.long))
    step = SemiStep(model, teacher, optimizer, sup_loss_fn, cfg, memobank, queue_ptrlis, queue_size, amp=not args.fp32)
    step.timers = {}

    image_l, label_l, image_u = [t.pin_memory() for t in synth_batch(1234 + rank, bl, bu, crop, C)]
    np.random.seed(1234 + rank)
    torch.manual_seed(1234 + rank)
    d_l, d_lab, d_u = image_l.to(dev), label_l.to(dev), image_u.to(dev)
    h2d = image_l.numel() * 4 + label_l.numel() * 8 + image_u.numel() * 4
    it = [EPOCH * LEN_LOADER]

    def run_resident():
        out = step(d_l, d_lab, d_u, EPOCH, it[0], LEN_LOADER)
        it[0] += 1
        return out

    def run_e2e():
        a = image_l.to(dev, non_blocking=True)
        b = label_l.to(dev, non_blocking=True)
        c = image_u.to(dev, non_blocking=True)
        out = step(a, b, c, EPOCH, it[0], LEN_LOADER)
        it[0] += 1
        return out.cpu()                                      # device->host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            out = fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)         # max over ranks, timed on the device
        return ms.item() / steps, out

    fast = bool(os.environ.get("U2PL_BENCH_FAST"))           # profiling runs only (ncu): fewer untimed steps
    for _ in range(args.warmup if fast else max(args.warmup, 3)):
        run_resident()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = _lib.launch_count()
    step.timers = {"entropy_partition": []}
    if args.phases:                                          # per-phase CUDA-event brackets (analysis runs)
        step.timers.update({name: [] for name in SemiStep.PHASES})
    ms_step, losses = timed(run_resident, args.steps)
    launches = (_lib.launch_count() - launches0) // args.steps
    torch.cuda.synchronize()
    ep_us = [a.elapsed_time(b) * 1e3 for a, b in step.timers.pop("entropy_partition")]
    phases_ms = {k: float(np.median([a.elapsed_time(b) for a, b in v])) for k, v in step.timers.items() if v}
    step.timers = {}
    if fast:
        ms_e2e = float("nan")
    else:
        run_e2e()
        ms_e2e, _ = timed(run_e2e, args.steps)
    clk = clocks.stop() if rank == 0 else None
    imgs = (bl + bu) * world
    value = imgs / (ms_step * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            mp = json.load(fh)
        peaks = {"hbm_gbs": mp["hbm_gbs"], "bf16_tflops_sustained": mp["bf16_tflops_sustained"], "src": "measured"}
    except Exception:
        pass
    N = bu * crop * crop
    alg_bytes = (4 * C + 25) * N                                # (4C+25) B/pixel, entropy-partition
    ep = float(np.median(ep_us)) if ep_us else None
    traffic = None                                              # dram bytes of the dominant kernel from one `ncu --set full`
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            traffic = json.load(fh).get("entropy_chain_kernel_dram_bytes")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "entropy_chain_kernel -- ONE cooperative launch for softmax-entropy + three percentile "
                "thresholds (np.percentile float32 semantics) + reliable/unreliable partition (u2pl_entropy_partition_fused; "
                "two tiny memsets precede it), CUDA events on the launching stream inside the timed step",
                "achieved": (alg_bytes / (ep * 1e-6) / 1e9) if ep else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": (alg_bytes / (ep * 1e-6) / 1e9 / peaks["hbm_gbs"]) if ep else None, "traffic": traffic,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the same kernel, one launch, ncu --set full "
                                "(profiles/r02_ncu_summaries.txt); same scope as algorithmic_bytes",
                "us_per_call": ep, "algorithmic_bytes": alg_bytes, "peak_source": peaks["src"]}
    flop_step = FWD_GFLOP_PER_IMG.get(crop, 651.1) * 1e9 * (bu + 3 * (bl + bu) + (bl + bu))   # T1 + S fwd+bwd(2x) + T2
    tensor = {"bound": "tensor", "scope": "whole step (network = 3 passes)", "achieved": flop_step / (ms_step * 1e-3) / 1e12,
              "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
              "frac": flop_step / (ms_step * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"], "model_tflop_per_step": flop_step / 1e12}
    from u2pl_b200 import fused as _fused
    bn_fallbacks = _fused.FALLBACKS["bn_module"]
    new_keys_last = int(sum(step.last.get("new_keys", [0])))
    eager = None
    if world == 1 and not args.no_eager_baseline and not fast:
        # the north star's ">= 10x the reference PyTorch-eager step" comparator, same GPU, same run: free our step first
        del step, model, teacher, optimizer, run_resident, run_e2e
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        try:
            ea = argparse.Namespace(**vars(args))
            ea.steps, ea.warmup = min(args.steps, 5), 3
            el = eager_arm(ea, quiet=True)
            eager = {"value": el["value"], "unit": "images/s", "ms_per_step": el["ms_per_step"], "dtype": el["dtype"],
                     "global_batch": el["config"]["global_batch"], "steps": ea.steps, "warmup": 3,
                     "impl": "oracle/eager_step.py (the reference step with the reference's device placement)",
                     "speedup_value": value / el["value"], "speedup_e2e": (imgs / (ms_e2e * 1e-3)) / el["value"]}
        except Exception as ex:                                  # reported, never fatal
            eager = {"error": repr(ex)}
        torch.cuda.empty_cache()
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = {k: v for k, v in reference_arm(args, quiet=True).items() if k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:                                  # reported, never fatal
            cpu = {"error": repr(ex)}
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.fp32 else "bf16", "data": "synthetic",
            "config": {"workload": f"train_semi.py U2PL step (VOC-style), {arch}-DeepLabv3+ {crop}x{crop} C={C}, "
                                   f"{bl}+{bu} crops per GPU, epoch {EPOCH}/80, banks full (30k/50k x 256)",
                       "global_batch": imgs, "parallelism": f"dp{world}", "l2": "inputs_exceed_L2",
                       "network": NETWORK_NOTE,
                       "routing": {k: v for k, v in _fused.ENABLED.items()},
                       "bn_module_fallbacks_total": bn_fallbacks,
                       "classifier_peak_scale": PEAK,
                       "bank": "class-sharded, peer-mapped (U2PL_BANK_SHARDED=1)"
                               if world > 1 and os.environ.get("U2PL_BANK_SHARDED", "0") == "1" else "replicated per GPU",
                       "infonce_depth": int(os.environ.get("U2PL_INFONCE_DEPTH", "2"))},
            "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": 12 * world, "ms_per_step": ms_e2e},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "tensor_roofline": tensor,
            "cpu_baseline": cpu, "eager_baseline": eager, "phases_ms": phases_ms or None, "losses": [float(x) / world for x in losses.cpu()],
            "new_keys_last_step": new_keys_last}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager"])
    ap.add_argument("--workload", default="v16", choices=sorted(WORKLOADS))
    ap.add_argument("--fp32", action="store_true", help="network in fp32 (TF32 off) instead of bf16 autocast")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the torch-eager comparator leg (N=1 only)")
    ap.add_argument("--phases", action="store_true", help="also report median CUDA-event time per phase of the step")
    args = ap.parse_args()
    if args.impl == "reference":
        print_reference(args)
    elif args.impl == "eager":
        eager_arm(args)
    else:
        our_arm(args)


if __name__ == "__main__":
    main()
