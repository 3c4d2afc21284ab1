"""Data-parallel training step (BASELINE config 3: SPVD fwd+bwd+Adam, B=32 x 2048 per GPU, NCCL gradient all-reduce):
python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench_ddp.py [steps]"""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, spvd_b200
from spvd_b200.sampler import torch2sparse
from spvd_b200.parallel import GradientAllReduce, train_step
from spvd_b200.optim import FlatAdam

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0); model = spvd_b200.build_model("SPVD"); model = model.to(dev).train()
red = GradientAllReduce(model.parameters(), overlap=os.environ.get('SPVD_DDP_OVERLAP', '1') == '1')
opt = FlatAdam(red, lr=1e-4) if os.environ.get('SPVD_FLAT_ADAM', '1') == '1' else torch.optim.Adam(model.parameters(), lr=1e-4)
W = 3
strat, xs, ts = bench.make_workload(steps + W, 32, 2048, 7 + rank)
times = []
for j in range(steps + W):
    st = torch2sparse(xs[j].to(dev))
    t = torch.randint(0, 1000, (32,), device=dev)
    eps = torch.randn(st.F.shape, device=dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    train_step(model, opt, (st, t, eps), red)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if j >= W:
        times.append(dt)
ms = torch.tensor([sum(times) / len(times) * 1e3], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"SPVD training step, B=32x2048 per GPU, {world} GPU(s): {float(ms):.1f} ms/step -> {world * 32 / float(ms) * 1e3:.0f} clouds/s")
if world > 1:
    dist.destroy_process_group()
