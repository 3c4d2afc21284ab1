"""Where does the HOST spend a training step?  cProfile over a few steps (SPVD, B=32x2048).  usage: python tools/_train_host_profile.py"""
import cProfile, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, spvd_b200
from spvd_b200.parallel import GradientAllReduce, train_step
from spvd_b200.optim import FlatAdam
from spvd_b200.sampler import torch2sparse

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = spvd_b200.build_model("SPVD").to(dev).train()
red = GradientAllReduce(model.parameters())
opt = FlatAdam(red, lr=1e-4)
strat, xs, ts = bench.make_workload(12, 32, 2048, 7)
batches = []
for j in range(12):
    st = torch2sparse(xs[j].to(dev))
    batches.append((st, torch.randint(0, 1000, (32,), device=dev), torch.randn(st.F.shape, device=dev)))
for j in range(4):
    train_step(model, opt, batches[j], red)
torch.cuda.synchronize()
t0 = time.perf_counter()
for j in range(4, 8):
    train_step(model, opt, batches[j], red)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue time {1e3 * (t1 - t0) / 4:.2f} ms/step, with final sync {1e3 * (t2 - t0) / 4:.2f} ms/step")
for rep in range(2):
    t0 = time.perf_counter()
    geo = model.prefetch_geometry(batches[0][0], 32)
    for j in range(12):
        geo_next = model.prefetch_geometry(batches[(j + 1) % 12][0], 32) if j + 1 < 12 else None
        train_step(model, opt, batches[j], red, geometry=geo)
        geo = geo_next
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"with geometry prefetch: host issue time {1e3 * (t1 - t0) / 12:.2f} ms/step, with final sync {1e3 * (t2 - t0) / 12:.2f} ms/step")
# host time per phase (no device syncs inside: what the Python thread spends issuing)
from spvd_b200 import functional as SF
ph = [0.0] * 5
for j in range(4, 8):
    x, t, target = batches[j]
    a = time.perf_counter(); red.zero_grad()
    b = time.perf_counter(); pred = model((x, t))
    c = time.perf_counter(); loss = SF.mse_loss(pred, target)
    d = time.perf_counter(); loss.backward()
    e = time.perf_counter(); opt.step(max_norm=10.0, pending_scale=red.finish())
    f = time.perf_counter()
    for i, v in enumerate((b - a, c - b, d - c, e - d, f - e)):
        ph[i] += v / 4
torch.cuda.synchronize()
print("host ms/step: zero_grad %.2f  forward %.2f  loss %.2f  backward %.2f  optimizer %.2f" % tuple(1e3 * v for v in ph))
if hasattr(model, "_geo_trace"):
    pass
pr = cProfile.Profile()
pr.enable()
for j in range(8, 12):
    train_step(model, opt, batches[j], red)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
