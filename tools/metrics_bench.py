"""Chamfer + approximate EMD throughput on cloud pairs of 2048 points (the shape of the reference's pairwise evaluation,
metrics/evaluation_metrics.py:68-101: one generated cloud against a batch of reference clouds)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spvd_b200 import metrics as gm

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator().manual_seed(0)
a = (torch.rand(1, 2048, 3, generator=g) - 0.5).expand(B, -1, -1).contiguous().cuda()
b = (torch.rand(B, 2048, 3, generator=g) - 0.5).cuda()
cham = gm.chamfer_3DDist()
for name, fn in (("chamfer", lambda: cham(a, b)), ("emd", lambda: gm.EMD(a, b, transpose=False))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: B={B} pairs of 2048x2048 points: {ms:.2f} ms -> {B / ms * 1e3:.0f} pairs/s")
