"""Batch-sharded data parallelism for the denoiser (BASELINE config 3): one process per GPU, every rank
trains on its own clouds, one NCCL all-reduce (sum, then / world) of the gradients per step, issued before
gradient clipping so the clip sees the global gradient (utils/callbacks.py:62-63 clips at 10.0).

The reference has no distributed code; this is the only strategy the path needs:
nothing in the forward is coupled across clouds except BatchNorm statistics, which stay per-rank exactly like
the single-GPU reference (no SyncBN).  Sampling needs no collective at all.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_clouds(n_clouds: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of the clouds owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n_clouds, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradientAllReduce:
    """Flat-bucket gradient all-reduce.  Gradients are copied into a few large contiguous buckets
    (default 64 MiB: NVSwitch makes the cost latency- not link-bound, so few big messages win), reduced
    asynchronously on the NCCL stream and copied back; ``average`` divides by the world size."""

    def __init__(self, params, bucket_bytes: int = 64 << 20, average: bool = True, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.average, self.group = average, group
        self.buckets, cur, size = [], [], 0
        for p in self.params:
            n = p.numel() * p.element_size()
            if cur and size + n > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += n
        if cur:
            self.buckets.append(cur)
        self._flat = [None] * len(self.buckets)

    def __call__(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        works = []
        for i, bucket in enumerate(self.buckets):
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            flat = torch.cat([g.reshape(-1) for g in grads])
            self._flat[i] = (flat, grads)
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for i, (w, bucket) in enumerate(zip(works, self.buckets)):
            w.wait()
            flat, grads = self._flat[i]
            if self.average:
                flat.div_(world)
            off = 0
            for p, g in zip(bucket, grads):
                n = g.numel()
                if p.grad is None:
                    p.grad = flat[off: off + n].view_as(p).clone()
                else:
                    p.grad.copy_(flat[off: off + n].view_as(p))
                off += n
            self._flat[i] = None


def train_step(model, optimizer, batch, reducer: GradientAllReduce = None, clip: float = 10.0, scheduler=None):
    """One optimisation step of train_generation.py:98-112: MSE(model((x_t, t)), eps), backward, gradient
    all-reduce (data parallel), clip_grad_norm_(10), optimizer (+ LR scheduler) step.  ``batch`` = (SparseTensor, t, eps)."""
    x, t, target = batch
    optimizer.zero_grad(set_to_none=True)
    pred = model((x, t))
    loss = torch.nn.functional.mse_loss(pred, target)
    loss.backward()
    if reducer is not None:
        reducer()
    torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss.detach()
