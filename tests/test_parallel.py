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


def test_flat_gradient_slots_and_in_place_claims():
    """Host logic of the in-place parameter gradients (no kernels, CPU tensors): slots start on 256-byte boundaries and the
    padding stays zero; a slot can be claimed once per step and only by a parameter used exactly once since zero_grad();
    a direct write counts towards the bucket like an accumulate hook; everything else goes through autograd as before."""
    from spvd_b200.parallel import GradientAllReduce
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.SiLU(), torch.nn.Linear(5, 3))
    red = GradientAllReduce(net.parameters(), bucket_bytes=64)
    ps = list(net.parameters())
    assert red.flat.numel() == sum((p.numel() + 63) // 64 * 64 for p in ps)
    assert all(red._offset_of(p) % 256 == 0 for p in ps)
    assert all(p.grad.data_ptr() == red.flat.data_ptr() + red._offset_of(p) and p.grad.shape == p.shape for p in ps)
    assert all(getattr(p, "_spvd_sink") is red for p in ps)
    w = ps[0]
    red.zero_grad()
    assert red.claim(w) is None                      # not used in a forward since zero_grad()
    red.note_use(w)
    slot = red.claim(w)
    assert slot is not None and slot.data_ptr() == w.grad.data_ptr()
    pending_before = red.buckets[red._bucket_of[id(w)]][2]
    slot.copy_(torch.ones_like(w))                   # "the kernel writes the gradient"
    red.wrote(w)
    assert red.buckets[red._bucket_of[id(w)]][2] == pending_before - 1
    assert red.claim(w) is None                      # once per step
    red.zero_grad()
    red.note_use(w); red.note_use(w)
    assert red.claim(w) is None                      # used twice: autograd accumulates
    red.zero_grad()
    red.note_use(w)
    w.grad = torch.zeros_like(w)                     # somebody replaced .grad: no claim, finish() would re-attach it
    assert red.claim(w) is None
    red.direct = False
    red.zero_grad()
    red.note_use(ps[1])
    assert red.claim(ps[1]) is None
    # the ordinary path still works with the padded layout: gradients land in the slots, padding stays zero
    red2 = GradientAllReduce(net.parameters(), bucket_bytes=64)
    red2.zero_grad()
    net(torch.randn(4, 7)).sum().backward()
    used = torch.zeros_like(red2.flat, dtype=torch.bool)
    for p in ps:
        o = red2._offset_of(p) // 4
        used[o:o + p.numel()] = True
        assert torch.equal(red2.flat[o:o + p.numel()].view_as(p), p.grad)
    assert float(red2.flat[~used].abs().max()) == 0.0
    assert float(red2.clip_(1e9)) > 0


def test_flat_adam_refuses_cpu_parameters():
    import pytest
    from spvd_b200.optim import FlatAdam
    from spvd_b200.parallel import GradientAllReduce
    net = torch.nn.Linear(4, 4)
    with pytest.raises(RuntimeError):
        FlatAdam(GradientAllReduce(net.parameters()))
