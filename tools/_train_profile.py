import os, sys, cProfile, pstats, io
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, spvd_b200
from spvd_b200.sampler import torch2sparse
dev = torch.device("cuda", 0)
torch.manual_seed(0); model = spvd_b200.build_model("SPVD"); model = model.to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
strat, xs, ts = bench.make_workload(6, 32, 2048, 7)
def step(j):
    st = torch2sparse(xs[j].to(dev))
    t = torch.randint(0, 1000, (32,), device=dev)
    eps = torch.randn(st.F.shape, device=dev)
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(model((st, t)), eps)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
    opt.step()
    torch.cuda.synchronize()
for j in range(3): step(j)
pr = cProfile.Profile(); pr.enable()
for j in range(3, 6): step(j)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
