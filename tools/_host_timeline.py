"""Host-side timeline of one sampling step in steady state (no device syncs between steps): when the planner is enqueued,
when each level group is collected, when the op ranges are launched.  usage: python tools/_host_timeline.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, spvd_b200
from spvd_b200.sampler import DDPMSparseSchedulerGPU

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = spvd_b200.build_model("SPVD").to(dev).eval()
sched = DDPMSparseSchedulerGPU()
B, N = 32, 2048
strat, xs, ts = bench.make_workload(14, B, N, 1234)
xs = [x.to(dev) for x in xs]
marks = []


def step(j, trace=False):
    t0 = time.perf_counter()
    st = sched.torch2sparse(xs[j], (B, N, 3))
    tb = torch.full((B,), ts[j], device=dev, dtype=torch.long)
    t1 = time.perf_counter()
    if trace:
        for pool in model._plan_ws.values():
            for w in pool[0]:
                w.trace = []
        for p in model._programs.values():
            p.trace = []
    eps = model((st, tb))
    t2 = time.perf_counter()
    out = sched.update_rule(st, eps, ts[j], j, (B, N, 3), dev)
    t3 = time.perf_counter()
    if trace:
        ev = [("step start", t0), ("torch2sparse issued", t1)]
        for pool in model._plan_ws.values():
            for w in pool[0]:
                ev += w.trace or []
                w.trace = None
        for p in model._programs.values():
            ev += p.trace or []
            p.trace = None
        ev += [("forward returned", t2), ("update issued", t3)]
        marks.append(sorted(ev, key=lambda e: e[1]))
    return out


with torch.no_grad():
    for j in range(6):
        step(j)
    for j in range(6, 12):
        step(j, trace=True)
    torch.cuda.synchronize()
for ev in marks[2:5]:
    t0 = ev[0][1]
    print("  ".join(f"{k} +{(v - t0) * 1e3:.3f}" for k, v in ev))
