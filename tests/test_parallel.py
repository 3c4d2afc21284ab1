"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel path (cloud sharding and the
bucketed gradient all-reduce).  The CUDA kernels are not involved."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spvd_b200.parallel import GradientAllReduce, shard_clouds
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.SiLU(), torch.nn.Linear(16, 3))
    data = torch.randn(10, 8, generator=torch.Generator().manual_seed(1))
    target = torch.randn(10, 3, generator=torch.Generator().manual_seed(2))
    lo, hi = shard_clouds(10, rank, world)
    # sum-of-squares loss scaled so that averaging shard gradients equals the full-batch gradient
    loss = ((net(data[lo:hi]) - target[lo:hi]) ** 2).sum() * world / 10
    loss.backward()
    GradientAllReduce(net.parameters(), bucket_bytes=256)()      # tiny buckets: several messages
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.SiLU(), torch.nn.Linear(16, 3))
    ref.load_state_dict(net.state_dict())
    (((ref(data) - target) ** 2).sum() / 10).backward()
    err = max(float((a.grad - b.grad).abs().max()) for a, b in zip(net.parameters(), ref.parameters()))
    # second scenario: the reducer exists before backward (gradients accumulate into its flat buffer, the per-parameter
    # hooks launch the buckets during backward), the 1 / world average is folded into the clip
    red = GradientAllReduce(net.parameters(), bucket_bytes=256)
    for _ in range(2):                                           # two steps: the buffer and the hooks are reusable
        red.zero_grad()
        (((net(data[lo:hi]) - target[lo:hi]) ** 2).sum() * world / 10 * 50).backward()
        assert all(p.grad.data_ptr() == red.flat.data_ptr() + red._offset_of(p) for p in net.parameters())
        norm = red.clip_(1.0, red.finish())
    ref.zero_grad()
    (((ref(data) - target) ** 2).sum() / 10 * 50).backward()
    want_norm = torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
    err2 = max(float((a.grad - b.grad).abs().max()) for a, b in zip(net.parameters(), ref.parameters()))
    err2 = max(err2, abs(float(norm) - float(want_norm)) / float(want_norm))
    if rank == 0:
        out.put(max(err, err2))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_clouds_partitions_exactly():
    from spvd_b200.parallel import shard_clouds
    for n in (1, 7, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [shard_clouds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_gradient_allreduce_equals_full_batch_gradient():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-6
