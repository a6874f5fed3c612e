"""torch.profiler breakdown of the loss-side kernels at V16 size (scratch tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from u2pl_b200 import ops
B, C, H, W = 16, 21, 513, 513
g = torch.Generator(device="cuda").manual_seed(0)
low = torch.randn(B, C, 129, 129, device="cuda", generator=g) * 3
x = torch.nn.functional.interpolate(low, (H, W), mode="bilinear", align_corners=True).contiguous()
student = torch.randn(B, C, H, W, device="cuda", generator=g)
target = x.argmax(1)
def run():
    ent, th, _ = ops.entropy_thresholds(x, target, [90.0, 10.0, 90.0])
    ent2, th2, _ = ops.entropy_thresholds(x, target, [90.0], exact_map=True)
    t2 = target.clone()
    n_kept, _ = ops.partition_target_(ent, t2, th, 0)
    p = student.detach().requires_grad_(True)
    l = ops.unsup_ce(p, t2, n_kept); l.backward()
for _ in range(3): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): run()
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:22]:
    print(f"{e.device_time_total/e.count:9.1f} us x{e.count:3d}  {e.key[:110]}")
