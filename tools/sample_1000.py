"""Closed-loop DDPM sampling (BASELINE config 2): 1000 steps, B=32 x 2048, random-init SPVD.  Random weights make the
trajectory meaningless (and it may leave the packed coordinate range, which raises), so this is a robustness / wall-clock
check of the loop, not a quality number."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spvd_b200
from spvd_b200.sampler import DDPMSparseSchedulerGPU

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
torch.manual_seed(0)
model = spvd_b200.build_model("SPVD")
model = model.cuda().eval()
sched = DDPMSparseSchedulerGPU(n_steps=steps)
torch.manual_seed(0)
sched.sample(model, 32, 2048)            # warm-up (graph capture, program compile)
torch.cuda.synchronize()
t0 = time.perf_counter()
try:
    out = sched.sample(model, 32, 2048)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{steps} steps: {dt:.3f} s wall, {steps / dt:.1f} steps/s, output finite: {bool(torch.isfinite(out).all())}, "
          f"|x|max = {float(out.abs().max()):.3g}")
except RuntimeError as e:
    print("sampling stopped:", str(e)[:200])
