"""tools/launch_list_summary.py <ncu gpu__time_duration csv> -- share of one training step per kernel from the ncu launch
list of `bench.py --steps 1 --warmup 1` (cold-cache, serialised times: compare SHARES, not absolutes).  One step = the
launches after the fused SGD+EMA kernel of the warm-up step up to and including that of the timed step."""
import csv, re, sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as fh:
    rd = csv.reader(l for l in fh if l.startswith('"'))
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        us = v / 1e3 if u in ("nsecond", "ns") else (v if u in ("usecond", "us") else v * 1e3 if u in ("msecond", "ms") else v / 1e3)
        rows.append((r[ki], us))
marks = [i for i, (k, _) in enumerate(rows) if "sgd_ema_kernel" in k]           # the last launch of every step
assert len(marks) >= 2, f"need two steps in the capture, found {len(marks)} optimizer launches"
step = rows[marks[-2] + 1:marks[-1] + 1]                                        # the timed step (the one before it warmed cuDNN's autotuner)


def short(k):
    k = re.sub(r"\(.*", "", k)
    k = re.sub(r"^void ", "", k)
    k = re.sub(r"at::native::(\(anonymous namespace\)::)?", "at::", k)
    return k[:110]


agg = defaultdict(lambda: [0.0, 0])
for k, us in step:
    a = agg[short(k)]
    a[0] += us
    a[1] += 1
tot = sum(a[0] for a in agg.values())
own = sum(a[0] for k, a in agg.items() if "u2pl::" in k)
conv_own = sum(a[0] for k, a in agg.items() if "conv_flat" in k or "wgrad_tc" in k or "conv_tc" in k)
lib_conv = sum(a[0] for k, a in agg.items() if any(t in k for t in ("cutlass", "cudnn", "nvjet", "xmma", "implicit_gemm")))
bn = sum(a[0] for k, a in agg.items() if "u2pl::bn_" in k)
print(f"launches in the capture: {len(rows)}; one step = {len(step)} launches, {tot/1e3:.2f} ms of serialised kernel time")
print(f"own kernels (u2pl::*): {100*own/tot:.1f} %   of which tcgen05 convolution / weight-gradient kernels: {100*conv_own/tot:.1f} %")
print(f"library GEMM / convolution kernels (cuDNN, cuBLAS): {100*lib_conv/tot:.1f} %   own BatchNorm kernels: {100*bn/tot:.1f} %")
print(f"{'us':>10s} {'share':>6s} {'n':>5s}  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{a[0]:10.1f} {100*a[0]/tot:5.1f}% {a[1]:5d}  {k}")
