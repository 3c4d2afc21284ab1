"""Kernel-time breakdown of one sampling step (SPVD, B=32x2048, executor): torch profiler, CUDA activity.
usage: python tools/step_profile.py"""
import os, sys, collections, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, spvd_b200
from spvd_b200.sampler import DDPMSparseSchedulerGPU
from torch.profiler import ProfilerActivity, profile

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = spvd_b200.build_model("SPVD").to(dev).eval()
sched = DDPMSparseSchedulerGPU()
B, N = 32, 2048
strat, xs, ts = bench.make_workload(12, B, N, 1234)
xs = [x.to(dev) for x in xs]
def step(j):
    st = sched.torch2sparse(xs[j], (B, N, 3))
    tb = torch.full((B,), ts[j], device=dev, dtype=torch.long)
    eps = model((st, tb))
    return sched.update_rule(st, eps, ts[j], j, (B, N, 3), dev)
with torch.no_grad():
    for j in range(4):
        step(j)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for j in range(4, 8):
        step(j)
    torch.cuda.synchronize()
    print("wall ms/step", (time.perf_counter() - t0) * 1e3 / 4)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(8)
        torch.cuda.synchronize()
    # steady state: three steps back to back (the host runs ahead of the device), the middle one analysed
    with profile(activities=[ProfilerActivity.CUDA]) as prof3:
        for j in (9, 10, 11):
            step(j)
        torch.cuda.synchronize()
ev = [e for e in prof.events() if "cuda" in str(getattr(e, "device_type", "")).lower()]
tot = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    k = e.name.split("(")[0][:80]
    tot[k][0] += 1
    tot[k][1] += e.device_time
all_us = sum(v[1] for v in tot.values())
t_lo = min(e.time_range.start for e in ev); t_hi = max(e.time_range.end for e in ev)
print(f"GPU kernel time {all_us/1e3:.3f} ms in {len(ev)} launches; first-to-last {(t_hi - t_lo)/1e3:.3f} ms")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{v[1]/1e3:8.3f} ms {v[0]:5d}  {k}")

# ---- timeline: when is NO kernel running, and between which kernels?
iv = sorted(((e.time_range.start, e.time_range.end, e.name.split("(")[0][:60]) for e in ev), key=lambda x: x[0])
busy_end, idle, gaps = iv[0][0], 0.0, []
last_name = ""
for s, e, nme in iv:
    if s > busy_end:
        idle += s - busy_end
        gaps.append((s - busy_end, last_name, nme, s - t_lo))
    if e > busy_end:
        busy_end, last_name = e, nme
print(f"idle (no kernel on any stream) {idle/1e3:.3f} ms of {(t_hi - t_lo)/1e3:.3f} ms in {len(gaps)} gaps")
hist = collections.Counter(min(int(g[0]), 20) for g in gaps)
print("gap histogram (us: count):", dict(sorted(hist.items())))
for g in sorted(gaps, key=lambda g: -g[0])[:25]:
    print(f"  {g[0]:7.1f} us at +{g[3]/1e3:6.3f} ms   {g[1]}  ->  {g[2]}")


# ---- the same for a step in steady state (middle of three back-to-back steps)
ev3 = sorted((e for e in prof3.events() if "cuda" in str(getattr(e, "device_type", "")).lower()), key=lambda e: e.time_range.start)
upd = [i for i, e in enumerate(ev3) if "k_ddpm_update" in e.name]
# a step ends with its k_ddpm_update_min launches; take the kernels between the end of step 9's update and step 10's
ends = [i for i in upd if i + 1 >= len(ev3) or "k_ddpm_update" not in ev3[i + 1].name]
if len(ends) >= 2:
    lo, hi = ends[0] + 1, ends[1] + 1
    win = ev3[lo:hi]
    w0, w1 = win[0].time_range.start, max(e.time_range.end for e in win)
    busy_end, idle, gaps, last = w0, 0.0, [], ""
    for e in win:
        s_, e_ = e.time_range.start, e.time_range.end
        if s_ > busy_end:
            idle += s_ - busy_end
            gaps.append((s_ - busy_end, last, e.name.split("(")[0][:60], s_ - w0))
        if e_ > busy_end:
            busy_end, last = e_, e.name.split("(")[0][:60]
    print(f"steady state: step window {(w1 - w0)/1e3:.3f} ms, {len(win)} launches, idle {idle/1e3:.3f} ms in {len(gaps)} gaps")
    for g in sorted(gaps, key=lambda g: -g[0])[:12]:
        print(f"  {g[0]:7.1f} us at +{g[3]/1e3:6.3f} ms   {g[1]}  ->  {g[2]}")
    first_conv = next((e for e in win if "k_conv" in e.name), None)
    if first_conv is not None:
        print(f"first conv kernel at +{(first_conv.time_range.start - w0)/1e3:.3f} ms")
    for e in win[:int(os.environ.get("SPVD_TIMELINE_ROWS", "40"))]:
        print(f"    +{(e.time_range.start - w0)/1e3:6.3f} {e.device_time:7.1f} us  {e.name.split('(')[0][:70]}")
