"""tcgen05 GEMM (csrc/gemm_tc.cu) vs cuBLAS (torch.matmul) on the network's 1x1-convolution shapes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2pl_b200 import ops
def t(fn, iters=20):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
rows = []
for name, M, K, N in [("layer3 conv1 (32 img)", 135200, 1024, 256), ("layer3 conv3 (32 img)", 135200, 256, 1024),
                      ("layer4 conv1", 135200, 2048, 512), ("layer4 conv3", 135200, 512, 2048),
                      ("layer1 conv3 129^2", 532512, 64, 256), ("aspp 1x1", 135200, 2048, 256), ("square 8192", 8192, 8192, 8192)]:
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
    sc = torch.ones(N, device="cuda"); sh = torch.zeros(N, device="cuda")
    us = t(lambda: ops.gemm_bf16_tn(a, b)); use = t(lambda: ops.gemm_bf16_tn(a, b, sc, sh, True)); ub = t(lambda: a @ b.t())
    fl = 2.0 * M * N * K
    rows.append(dict(shape=name, M=M, K=K, N=N, ours_us=round(us, 1), ours_bn_relu_us=round(use, 1), cublas_us=round(ub, 1),
                     ours_tflops=round(fl / us / 1e6, 1), cublas_tflops=round(fl / ub / 1e6, 1)))
    print(rows[-1], flush=True)
json.dump(rows, open("gpurun_out/gemm_bench.json", "w"), indent=1)
