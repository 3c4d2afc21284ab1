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
pr = cProfile.Profile()
pr.enable()
for j in range(8, 12):
    train_step(model, opt, batches[j], red)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
