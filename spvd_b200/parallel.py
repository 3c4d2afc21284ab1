"""Batch-sharded data parallelism for the denoiser (BASELINE config 3): one process per GPU, every rank
trains on its own clouds, one NCCL all-reduce (sum, then / world) of the gradients per step, issued before
gradient clipping so the clip sees the global gradient (utils/callbacks.py:62-63 clips at 10.0).

The reference has no distributed code; this is the only strategy the path needs:
nothing in the forward is coupled across clouds except BatchNorm statistics, which stay per-rank exactly like
the single-GPU reference (no SyncBN).  Sampling needs no collective at all.

Layout: every gradient lives in ONE pre-allocated flat fp32 buffer (``p.grad`` are views into it, so there is no
concatenation and no copy-back), cut into contiguous buckets in reverse parameter order -- the order in which backward
finishes them.  A post-accumulate hook per parameter launches a bucket's all-reduce on the NCCL stream as soon as its
last gradient is final, i.e. while the rest of backward is still running; ``finish()`` waits for the stragglers and
the 1 / world average is folded into the clipping scale (one pass over the flat buffer).

The library's own backward functions (spvd_b200/functional.py) go one step further: for a parameter used once per step
their kernels write the gradient straight INTO its slot (``claim`` / ``wrote`` below) and hand autograd nothing, so there
is no temporary and no AccumulateGrad add per parameter either.
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
    def __init__(self, params, bucket_bytes: int = 32 << 20, average: bool = True, group=None, overlap: bool = True):
        self.params = [p for p in params if p.requires_grad]
        self.average, self.group = average, group
        if not self.params:
            raise ValueError("no parameters to reduce")
        dev, dt = self.params[0].device, self.params[0].dtype
        order = list(reversed(self.params))                      # backward produces the last layers' gradients first
        pad = lambda n: (n + 63) // 64 * 64                      # slots start on 256-byte boundaries: the backward kernels
        self.flat = torch.zeros(sum(pad(p.numel()) for p in order), dtype=dt, device=dev)   # store float4s into them; the
        self.buckets, self._bucket_of = [], {}                   # padding stays zero (neutral for the sum and the norm)
        self._offsets = {}                                       # bucket = [lo, hi, pending, n_params, work]
        off, lo, size = 0, 0, 0
        for p in order:
            n = p.numel()
            view = self.flat[off: off + n].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            self._bucket_of[id(p)] = len(self.buckets)
            self._offsets[id(p)] = off * p.element_size()
            off += pad(n)
            size += pad(n) * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append([lo, off, 0, 0, None])
                lo, size = off, 0
        if off > lo:
            self.buckets.append([lo, off, 0, 0, None])
        for p in order:
            self.buckets[min(self._bucket_of[id(p)], len(self.buckets) - 1)][3] += 1
            self._bucket_of[id(p)] = min(self._bucket_of[id(p)], len(self.buckets) - 1)
        self._hooks = []
        self.overlap = bool(overlap)
        if overlap and hasattr(torch.Tensor, "register_post_accumulate_grad_hook"):
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._ready))
        self._uses, self._written = {}, set()
        self.direct = True                                       # in-place parameter gradients (see claim)
        for p in self.params:
            p._spvd_sink = self
        self.reset()

    # -- in-place parameter gradients (called by spvd_b200.functional) ---------------------------------------------
    def note_use(self, p):
        self._uses[id(p)] = self._uses.get(id(p), 0) + 1         # (forwards since zero_grad(); grad mode is off inside Function.forward)

    def claim(self, p):
        """The zeroed slot of ``p``'s gradient if the calling backward may write it in place: p is used once since
        zero_grad(), nothing has been written yet and .grad still is the slot.  The caller MUST then write the full gradient
        there, return None to autograd and call ``wrote(p)``."""
        i = id(p)
        if (not self.direct or self._uses.get(i, 0) != 1 or i in self._written or i not in self._bucket_of or p.grad is None
                or p.grad.data_ptr() != self.flat.data_ptr() + self._offset_of(p)):
            return None
        self._written.add(i)
        return p.grad

    def wrote(self, p):
        if self.overlap:
            self._ready(p)

    # -- per step -------------------------------------------------------------------------------------------------
    def reset(self):
        for b in self.buckets:
            b[2], b[4] = b[3], None

    def zero_grad(self):
        """instead of optimizer.zero_grad(): one memset; the views stay attached"""
        self.flat.zero_()
        self._uses.clear()
        self._written.clear()
        self.reset()

    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _launch(self, b):
        if b[4] is None and self._active():
            b[4] = dist.all_reduce(self.flat[b[0]: b[1]], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _ready(self, p):
        if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + self._offset_of(p):
            return                                               # someone replaced .grad (set_to_none): fall back to finish()
        b = self.buckets[self._bucket_of[id(p)]]
        b[2] -= 1
        if b[2] == 0:
            self._launch(b)

    def _offset_of(self, p):
        return self._offsets[id(p)]

    def finish(self):
        """All buckets reduced (launching those whose hooks did not fire); returns the factor still owed (1 / world when
        averaging) -- apply it with ``scale_`` or fold it into the clip."""
        if not self._active():
            self.reset()
            return 1.0
        for p in self.params:                                    # gradients that were re-created outside the flat buffer
            if p.grad is not None and p.grad.data_ptr() != self.flat.data_ptr() + self._offset_of(p):
                view = self.flat[self._offset_of(p) // p.element_size():][: p.numel()].view_as(p)
                view.copy_(p.grad)
                p.grad = view
        for b in self.buckets:
            self._launch(b)
        for b in self.buckets:
            b[4].wait()
        self.reset()
        return 1.0 / dist.get_world_size(self.group) if self.average else 1.0

    def __call__(self):
        """synchronous form: reduce now and apply the average"""
        s = self.finish()
        if s != 1.0:
            self.flat.mul_(s)

    def clip_(self, max_norm: float, pending_scale: float = 1.0):
        """clip_grad_norm_ over the flat buffer with the owed average folded in: one norm, one multiply"""
        norm = torch.linalg.vector_norm(self.flat) * pending_scale
        scale = torch.clamp(max_norm / (norm + 1e-6), max=1.0) * pending_scale
        self.flat.mul_(scale)
        return norm


def train_step(model, optimizer, batch, reducer: GradientAllReduce = None, clip: float = 10.0, scheduler=None, geometry=None):
    """One optimisation step of train_generation.py:98-112: MSE(model((x_t, t)), eps), backward, gradient
    all-reduce (data parallel, overlapped with backward), clip_grad_norm_(10), optimizer (+ LR scheduler) step.
    ``batch`` = (SparseTensor, t, eps).  ``geometry``: the batch's geometry planned ahead with ``model.prefetch_geometry``."""
    from . import functional as SF
    x, t, target = batch
    if reducer is not None:
        reducer.zero_grad()
    else:
        optimizer.zero_grad(set_to_none=True)
    pred = model((x, t)) if geometry is None else model((x, t), geometry=geometry)
    loss = SF.mse_loss(pred, target) if pred.is_cuda else torch.nn.functional.mse_loss(pred, target)
    loss.backward()
    if reducer is not None and hasattr(optimizer, "flat_p"):      # optim.FlatAdam: norm + clip + average + Adam in two launches
        optimizer.step(max_norm=clip, pending_scale=reducer.finish())
    else:
        if reducer is not None:
            reducer.clip_(clip, reducer.finish())
        else:
            torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
        optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss.detach()
