"""Time one training step (fwd + bwd + Adam) of SPVD at B=32x2048 on the CUDA path: python tools/train_bench.py [steps]"""
import os, sys, time, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, spvd_b200
from spvd_b200.sampler import torch2sparse, DDPM
from spvd_b200.parallel import train_step

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B = int(os.environ.get("SPVD_BATCH", "32"))
dev = torch.device("cuda", 0)
torch.manual_seed(0); model = spvd_b200.build_model("SPVD"); model = model.to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
strat, xs, ts = bench.make_workload(steps + 2, B, 2048, 7)
E = lambda: torch.cuda.Event(enable_timing=True)
for j in range(steps + 2):
    st = torch2sparse(xs[j].to(dev))
    t = torch.randint(0, 1000, (B,), device=dev)
    eps = torch.randn(st.F.shape, device=dev)
    torch.cuda.synchronize()
    a, b, c = E(), E(), E()
    a.record()
    opt.zero_grad(set_to_none=True)
    pred = model((st, t))
    loss = torch.nn.functional.mse_loss(pred, eps)
    b.record()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
    opt.step()
    c.record()
    torch.cuda.synchronize()
    print(f"step {j}: fwd {a.elapsed_time(b):.1f} ms, bwd+opt {b.elapsed_time(c):.1f} ms, loss {loss.item():.4f}")
