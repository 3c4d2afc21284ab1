"""One training step (SPVD, B=32x2048: forward + MSE + backward + clip + Adam) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum`: the launch list shows which kernels a training step runs
(row TR: no cuBLAS / cuDNN / CUTLASS / ATen batch-norm / attention / index kernels).  usage: python tools/train_ncu.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, spvd_b200
from spvd_b200.parallel import GradientAllReduce, train_step
from spvd_b200.optim import FlatAdam
from spvd_b200.sampler import torch2sparse

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = spvd_b200.build_model(os.environ.get("SPVD_PRESET", "SPVD")).to(dev).train()
red = GradientAllReduce(model.parameters())
opt = FlatAdam(red, lr=1e-4)          # flat-buffer Adam + clip + average on the library's kernels
strat, xs, ts = bench.make_workload(4, 32, 2048, 7)
for j in range(4):
    st = torch2sparse(xs[j].to(dev))
    t = torch.randint(0, 1000, (32,), device=dev)
    eps = torch.randn(st.F.shape, device=dev)
    if j == 3:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    train_step(model, opt, (st, t, eps), red)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
