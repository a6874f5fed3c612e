"""Per-kernel device time of the network passes of one V16 step in isolation (torch.profiler): T1 = teacher eval
forward, no grad (16 images); T2 = teacher train-mode forward, no grad (32); S = student forward + backward (32, dummy
loss).  Scratch tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import u2pl_b200
u2pl_b200.install()
from u2pl.models.model_helper import ModelBuilder
from torch.profiler import profile, ProfilerActivity

cfg = bench.make_cfg("v16")
torch.manual_seed(1)
torch.backends.cudnn.benchmark = True
model = ModelBuilder(cfg["net"]).cuda().to(memory_format=torch.channels_last)
x16 = torch.randn(16, 3, 513, 513, device="cuda").contiguous(memory_format=torch.channels_last)
x32 = torch.randn(32, 3, 513, 513, device="cuda").contiguous(memory_format=torch.channels_last)


def t1():
    model.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return model(x16)["pred"]


def t2():
    model.train()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return model(x32)["pred"]


def s():
    model.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o = model(x32)
    (o["pred"].float().mean() + o["rep"].float().mean()).backward()
    model.zero_grad(set_to_none=True)


for name, fn in (("T1", t1), ("T2", t2), ("S", s)):
    if os.environ.get("ONLY") and os.environ["ONLY"] != name:
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
    rows = sorted(((e.device_time_total / 2, e.count // 2, e.key) for e in prof.key_averages() if e.device_time_total > 0), reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"== {name}: {tot/1e3:.2f} ms device time per pass, {sum(r[1] for r in rows)} launches")
    for t, n, k in rows[:int(os.environ.get("TOP", "28"))]:
        print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  n={n:5d}  {k[:130]}")
