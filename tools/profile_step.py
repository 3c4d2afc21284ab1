"""Run a few denoising steps (SPVD, B=32x2048) for ncu: python tools/profile_step.py [steps] [warmup]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import spvd_b200
from spvd_b200.sampler import DDPMSparseSchedulerGPU

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
preset = os.environ.get("SPVD_PRESET", "SPVD")
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = spvd_b200.build_model(preset)
model = model.to(dev).eval()
sched = DDPMSparseSchedulerGPU()
strat, xs, ts = bench.make_workload(steps + warm, 32, 2048, 1234)
shape = (32, 2048, 3)
with torch.no_grad():
    for j in range(steps + warm):
        if j == warm:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        st = sched.torch2sparse(xs[j].to(dev), shape)
        eps = model((st, torch.full((32,), ts[j], device=dev, dtype=torch.long)))
        new = strat.update(st.F, eps, ts[j], j, lambda s: torch.randn(s, device=dev))
        sched.torch2sparse(new, shape)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
