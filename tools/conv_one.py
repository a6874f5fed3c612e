"""tools/conv_one.py -- one layer through the tcgen05 implicit-GEMM conv, for `ncu --set full -k regex:conv_tc -c 1`:
    ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 1 -o gpurun_out/r02_conv_tc \\
        python tools/conv_one.py [aspp|layer3|1x1]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from u2pl_b200 import ops  # noqa: E402

SHAPES = {"aspp": (16, 2048, 65, 65, 256, 3, 12), "layer3": (16, 256, 65, 65, 256, 3, 2), "1x1": (16, 1024, 65, 65, 256, 1, 1)}
N, Cin, H, W, Cout, k, d = SHAPES[sys.argv[1] if len(sys.argv) > 1 else "layer3"]
x = torch.randn(N, Cin, H, W, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).bfloat16().contiguous(memory_format=torch.channels_last)
scale, shift = torch.rand(Cout, device="cuda") + 0.5, torch.randn(Cout, device="cuda")
for _ in range(4):
    y = ops.conv_bf16_nhwc(x, w, d, scale, shift, None, True)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
