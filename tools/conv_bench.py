"""tools/conv_bench.py -- tcgen05 implicit-GEMM conv (+folded BN, ReLU) vs cuDNN conv + u2pl bn_apply on the network's
stride-1 layer shapes at the teacher's T1 batch (16 x 65x65 / 129x129).  Run on the GPU box:
    python tools/conv_bench.py > gpurun_out/conv_bench.json"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from u2pl_b200 import ops  # noqa: E402

SHAPES = [  # name, N, Cin, H, W, Cout, k, d
    ("layer3.conv1 1x1 1024->256", 16, 1024, 65, 65, 256, 1, 1),
    ("layer3.conv2 3x3 d2 256->256", 16, 256, 65, 65, 256, 3, 2),
    ("layer3.conv3 1x1 256->1024", 16, 256, 65, 65, 1024, 1, 1),
    ("layer4.conv2 3x3 d8 512->512", 16, 512, 65, 65, 512, 3, 8),
    ("layer4.conv3 1x1 512->2048", 16, 512, 65, 65, 2048, 1, 1),
    ("aspp 3x3 d12 2048->256", 16, 2048, 65, 65, 256, 3, 12),
    ("aspp 3x3 d36 2048->256", 16, 2048, 65, 65, 256, 3, 36),
    ("head 3x3 1280->256", 16, 1280, 65, 65, 256, 3, 1),
    ("decoder 3x3 512->256 @129", 16, 512, 129, 129, 256, 3, 1),
    ("decoder 3x3 256->256 @129", 16, 256, 129, 129, 256, 3, 1),
    ("8192^3 GEMM as a 1x1 conv (128x256 tile; compare tools/gemm_bench.py: 128x128 tile, cuBLAS)", 1, 8192, 8192, 1, 8192, 1, 1),
]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    torch.backends.cudnn.benchmark = True
    rows = []
    for name, N, Cin, H, W, Cout, k, d in SHAPES:
        x = torch.randn(N, Cin, H, W, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
        w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).bfloat16().contiguous(memory_format=torch.channels_last)
        scale, shift = torch.rand(Cout, device="cuda") + 0.5, torch.randn(Cout, device="cuda")
        ours = timeit(lambda: ops.conv_bf16_nhwc(x, w, d, scale, shift, None, True))
        lib = timeit(lambda: F.relu(F.conv2d(x, w, None, 1, d * (k // 2), d) * scale.bfloat16()[None, :, None, None]
                                    + shift.bfloat16()[None, :, None, None]))
        conv_only = timeit(lambda: F.conv2d(x, w, None, 1, d * (k // 2), d))
        a = ops.conv_bf16_nhwc(x, w, d, scale, shift, None, True).float()
        b = F.relu(F.conv2d(x.float(), w.float(), None, 1, d * (k // 2), d) * scale[None, :, None, None] + shift[None, :, None, None])
        flop = 2.0 * N * H * W * Cout * Cin * k * k
        rows.append({"layer": name, "tc_ms": ours, "cudnn_conv_ms": conv_only, "cudnn_conv_bn_relu_eager_ms": lib,
                     "tc_tflops": flop / ours / 1e9, "cudnn_tflops": flop / conv_only / 1e9,
                     "max_err": float((a - b).abs().max()), "ref_max": float(b.abs().max())})
        print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
