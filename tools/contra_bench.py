"""BASELINE config 4: contrastive memobank micro-bench -- 19 banks x 30 000 x 256 fp32 resident, ~50 000
anchor candidates (uniform classes, prob > 0.3), 256 queries x 50 negatives per class, tau = 0.5.
Prints per-kernel device times (torch.profiler) and the InfoNCE gather bandwidth."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from u2pl_b200 import contra
C, D, h, w = 19, 256, 158, 158
g = torch.Generator(device="cuda").manual_seed(0)
cfg = dict(negative_high_entropy=True, low_rank=3, high_rank=20, current_class_threshold=0.3,
           current_class_negative_threshold=1, num_negatives=50, num_queries=256, temperature=0.5)
lab = torch.randint(0, C, (2, h, w), device="cuda", generator=g)
onehot = torch.nn.functional.one_hot(lab, C).permute(0, 3, 1, 2).contiguous()
prob = (onehot.float() * 0.5 + 0.5 / C).contiguous()
low = torch.ones(2, 1, h, w, device="cuda"); high = torch.zeros(2, 1, h, w, device="cuda")
rep = torch.randn(2, D, h, w, device="cuda", generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
rep_t = torch.randn(2, D, h, w, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
cpu_g = torch.Generator().manual_seed(1)
memobank = [[torch.randn(30000, D, generator=cpu_g)] for _ in range(C)]
ptrs = [torch.zeros(1, dtype=torch.long) for _ in range(C)]
qsize = [30000] * C
def run():
    rep.grad = None
    nk, loss = contra.compute_contra_memobank_loss(rep, onehot[:1], onehot[1:], prob[:1], prob[1:], low, high, cfg,
                                                   memobank, ptrs, qsize, rep_t)
    loss.backward()
    return loss
for _ in range(3): l = run()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        flush.zero_(); run()
    torch.cuda.synchronize()
res = {}
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]:
    print(f"{e.device_time_total/e.count:9.1f} us x{e.count:3d}  {e.key[:100]}")
    if "infonce_fwd" in e.key: res["infonce_fwd_us"] = e.device_time_total / e.count
nbytes = C * (256 + 256 * 50 + 1) * D * 4
res.update(loss=float(l), algorithmic_MB=nbytes / 1e6, infonce_fwd_GBps=nbytes / (res["infonce_fwd_us"] * 1e-6) / 1e9)
print(json.dumps(res))
