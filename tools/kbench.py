"""Kernel-level timing of the fused loss path at BASELINE config-2 size (scratch tool; bench.py is the contract)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2pl_b200 import ops

def timeit(fn, iters=20, warm=5):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

def main():
    B, C, H, W = 16, 21, 513, 513
    g = torch.Generator(device="cuda").manual_seed(0)
    low = torch.randn(B, C, 129, 129, device="cuda", generator=g) * 3
    x = torch.nn.functional.interpolate(low, (H, W), mode="bilinear", align_corners=True).contiguous()
    student = torch.randn(B, C, H, W, device="cuda", generator=g)
    target = x.argmax(1)
    N = B * H * W
    res = {}
    res["entropy_thresholds_exact_1q_us"] = timeit(lambda: ops.entropy_thresholds(x, target, [90.0], exact_map=True))
    res["entropy_thresholds_1q_us"] = timeit(lambda: ops.entropy_thresholds(x, target, [90.0]))
    res["entropy_thresholds_3q_us"] = timeit(lambda: ops.entropy_thresholds(x, target, [90.0, 10.0, 90.0]))
    ent, th, _ = ops.entropy_thresholds(x, target, [90.0])
    t2 = target.clone()
    res["partition_us"] = timeit(lambda: ops.partition_target_(ent, t2, th, 0))
    res["ce_fwd_us"] = timeit(lambda: ops._ce_forward(student, t2, 255))
    sc = torch.ones(1, device="cuda")
    res["ce_bwd_us"] = timeit(lambda: ops._ce_backward(student, t2, 255, sc))
    def full():
        p = student.detach().requires_grad_(True)
        l = ops.unsup_loss(p, target.clone(), 90.0, x)
        l.backward()
    res["unsup_loss_fwd_bwd_us"] = timeit(full)
    alg = (4 * C + 25) * N
    res["entropy_partition_GBps_algorithmic"] = alg / ((res["entropy_thresholds_1q_us"] + res["partition_us"]) * 1e-6) / 1e9
    res["ce_fwd_bwd_GBps_algorithmic"] = 8 * C * N / ((res["ce_fwd_us"] + res["ce_bwd_us"]) * 1e-6) / 1e9
    # torch eager reference of the same math, for scale
    def eager():
        prob = torch.softmax(x, 1)
        e = -torch.sum(prob * torch.log(prob + 1e-10), 1)
        return e
    res["torch_eager_entropy_only_us"] = timeit(eager)
    print(json.dumps(res, indent=1))

if __name__ == "__main__":
    main()
