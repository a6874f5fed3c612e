"""tools/chain_time.py -- fused entropy chain at V16 size: CUDA-event time of the one-launch path vs the multi-launch
path (+ clone + partition), and the per-phase %globaltimer stamps of CTA 0 (U2PL_CHAIN_TIMING=1)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2pl_b200 import ops

B, C, H, W = 16, int(os.environ.get("C", "21")), 513, 513
g = torch.Generator(device="cuda").manual_seed(0)
low = torch.randn(B, C, 129, 129, device="cuda", generator=g) * 3
x = torch.nn.functional.interpolate(low, (H, W), mode="bilinear", align_corners=True).contiguous()
target = x.argmax(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
P = [90.0, 10.0, 90.0]


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()                       # L2 flush between timed iterations
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def two_step():
    ent, th, _ = ops.entropy_thresholds(x, target, P)
    t2 = target.clone()
    ops.partition_target_(ent, t2, th, 0)


res = {"fused_us": timeit(lambda: ops.entropy_partition(x, target, P, 0)), "two_step_us": timeit(two_step)}
alg = (4 * C + 25) * B * H * W
res["algorithmic_MB"] = alg / 1e6
res["fused_GBps"] = alg / (res["fused_us"] * 1e-6) / 1e9
print(json.dumps(res))
