"""Where does one denoising step go?  CPU launch time vs GPU time, geometry vs feature path."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, spvd_b200
from spvd_b200.sampler import DDPMSparseSchedulerGPU

dev = torch.device("cuda", 0)
preset = os.environ.get("SPVD_PRESET", "SPVD")
torch.manual_seed(0); model = spvd_b200.build_model(preset); model = model.to(dev).eval()
sched = DDPMSparseSchedulerGPU()
K, W = 10, 3
strat, xs, ts = bench.make_workload(K + W, 32, 2048, 1234)
shape = (32, 2048, 3)
xs = [x.to(dev) for x in xs]
E = lambda: torch.cuda.Event(enable_timing=True)
rows = []
with torch.no_grad():
    for j in range(K + W):
        torch.cuda.synchronize()
        e0, e1, e2, e3 = E(), E(), E(), E()
        c0 = time.perf_counter(); e0.record()
        st = sched.torch2sparse(xs[j], shape)
        tb = torch.full((32,), ts[j], device=dev, dtype=torch.long)
        e1.record()
        for pool in model._plan_ws.values():
            for w in pool[0]:
                w.trace = []
        cg = time.perf_counter()
        geo = model.plan_geometry(st.coords_float, 32, st.points_per_cloud, model.plan_split_level)
        gtrace = [(k, round((v - cg) * 1e3, 3)) for pool in model._plan_ws.values() for w in pool[0] for k, v in (w.trace or [])]
        c1 = time.perf_counter(); e2.record()
        for p in model._programs.values():
            p.trace = [("geometry returned", c1)]
        eps = model((st, tb), geometry=geo)
        new = strat.update(st.F, eps, ts[j], j, lambda s: torch.randn(s, device=dev))
        sched.torch2sparse(new, shape)
        e3.record(); c2 = time.perf_counter()
        torch.cuda.synchronize(); c3 = time.perf_counter()
        if j >= W:
            rows.append((e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), (c1 - c0) * 1e3, (c2 - c1) * 1e3, (c3 - c0) * 1e3))
print("last step, ms after plan_geometry was entered:", gtrace)
for p in model._programs.values():
    print("last step, ms after the geometry call returned:", [(k, round((v - p.trace[0][1]) * 1e3, 3)) for k, v in p.trace[1:]])
import statistics
names = ["gpu t2s", "gpu geometry(incl sync)", "gpu feature path", "cpu till geometry done", "cpu launch feature path", "wall total"]
for i, n in enumerate(names):
    print(f"{n:28s} {statistics.mean(r[i] for r in rows):8.3f} ms")
