import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2pl_b200 import ops
M = N = K = 8192
a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
for _ in range(4): d = ops.gemm_bf16_tn(a, b)
torch.cuda.synchronize(); print(float(d.float().abs().mean()))
