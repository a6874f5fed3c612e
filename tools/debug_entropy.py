import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import port
from u2pl_b200 import ops
g = np.load("tests/golden/unsup_c21.npz")
x = g["pred_teacher"]; t = g["target_in"].astype(np.int64)
ent, th, nv = ops.entropy_thresholds(torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda(), [float(g["percent"])])
e = ent.cpu().numpy(); eo = port.entropy(x)
bad = (e.view(np.uint32) != eo.view(np.uint32))
print("entropy mismatches", bad.sum(), "of", bad.size, "max abs diff", np.abs(e - eo).max())
if bad.any():
    idx = np.argwhere(bad)[:5]
    for i in idx: print(i, e[tuple(i)], eo[tuple(i)])
print("thresh", th.cpu().numpy(), port.percentile(eo[t != 255], float(g["percent"])), "np on gpu ent", np.percentile(e[t != 255], float(g["percent"])), "n", nv.item())
for shape in [(1, 21, 8, 8), (1, 21, 16, 32), (1, 21, 32, 32), (2, 21, 33, 33)]:
    rng = np.random.default_rng(0)
    xx = rng.standard_normal(shape).astype(np.float32) * 3
    tt = np.zeros((shape[0],) + shape[2:], np.int64)
    ee, _, _ = ops.entropy_thresholds(torch.from_numpy(xx).cuda(), torch.from_numpy(tt).cuda(), [50.0])
    b = ee.cpu().numpy().view(np.uint32) != port.entropy(xx).view(np.uint32)
    print(shape, "mismatch", b.sum(), np.argwhere(b)[:6].tolist())
