"""Kernel-time breakdown of one training step (SPVD, B=32x2048): torch profiler, CUDA activity only.
usage: python tools/train_profile.py"""
import os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, spvd_b200
from spvd_b200.parallel import GradientAllReduce, train_step
from spvd_b200.optim import FlatAdam
from spvd_b200.sampler import torch2sparse
from torch.profiler import ProfilerActivity, profile

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = spvd_b200.build_model("SPVD").to(dev).train()
red = GradientAllReduce(model.parameters())
opt = FlatAdam(red, lr=1e-4)          # flat-buffer Adam + clip + average on the library's kernels
strat, xs, ts = bench.make_workload(6, 32, 2048, 7)
def step(j):
    st = torch2sparse(xs[j].to(dev))
    t = torch.randint(0, 1000, (32,), device=dev)
    eps = torch.randn(st.F.shape, device=dev)
    train_step(model, opt, (st, t, eps), red)
for j in range(3):
    step(j)
torch.cuda.synchronize()
import time
t0 = time.perf_counter(); step(3); torch.cuda.synchronize(); print("wall ms", (time.perf_counter() - t0) * 1e3)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(4)
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if "cuda" in str(getattr(e, "device_type", "")).lower():
        k = e.name.split("(")[0][:70]
        tot[k][0] += 1
        tot[k][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
all_us = sum(v[1] for v in tot.values())
print(f"GPU kernel time {all_us/1e3:.2f} ms in {sum(v[0] for v in tot.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1]/1e3:8.3f} ms {v[0]:5d}  {k}")
