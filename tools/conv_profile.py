"""Which convolutions are slow?  One student forward/backward under torch.profiler with shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, u2pl_b200
u2pl_b200.install()
from u2pl.models.model_helper import ModelBuilder
from torch.profiler import profile, ProfilerActivity
cfg = bench.make_cfg("v16")
torch.backends.cudnn.benchmark = True
m = ModelBuilder(cfg["net"]).cuda().to(memory_format=torch.channels_last)
x = torch.randn(32, 3, 513, 513, device="cuda").contiguous(memory_format=torch.channels_last)
def run():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o = m(x)
    (o["pred"].float().mean() + o["rep"].float().mean()).backward()
for _ in range(2): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    run(); torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.key in ("aten::convolution_backward", "aten::cudnn_convolution", "aten::convolution", "aten::mm", "aten::conv2d"):
        rows.append((e.device_time_total, e.count, e.key, str(e.input_shapes)[:150]))
rows.sort(reverse=True)
for t, n, k, s in rows[:30]:
    print(f"{t/1e3:8.2f} ms n={n:3d} {k:28s} {s}")
