"""Optimizer tail of the training step on the library's own kernels (train_generation.py:106-112: clip_grad_norm_(10),
Adam step, LR scheduler step).

``FlatAdam`` keeps the parameters, both Adam moments and -- through the ``GradientAllReduce`` it is built on -- the
gradients in flat fp32 buffers with one common slot layout, so a step is two launches whatever the number of parameters:
the squared gradient norm (``spvd_sumsq``) and clip + data-parallel average + Adam (``spvd_adam_step``), 7 streams over
the buffers instead of torch's ~15 (norm, scale, five foreach passes) in ~50 launches.  Semantics are torch.optim.Adam's
(no amsgrad; weight_decay = L2 added to the gradient) and torch.nn.utils.clip_grad_norm_'s; a torch LR scheduler can drive
it (it is a torch.optim.Optimizer with one parameter group).
"""
from __future__ import annotations

import torch

from . import ops


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, reducer, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        """reducer: the parallel.GradientAllReduce of the model (its slot layout is adopted; parameters must already
        live on their CUDA device).  Afterwards every ``p.data`` is a view into ``self.flat_p``."""
        params = list(reducer.params)
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        if not reducer.flat.is_cuda:
            raise RuntimeError("FlatAdam needs CUDA parameters (there is no CPU fallback)")
        self.reducer = reducer
        self.flat_p = torch.zeros_like(reducer.flat)
        with torch.no_grad():
            for p in params:
                o = reducer._offset_of(p) // p.element_size()
                view = self.flat_p[o: o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        self.exp_avg = torch.zeros_like(reducer.flat)
        self.exp_avg_sq = torch.zeros_like(reducer.flat)
        self.steps = 0
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=reducer.flat.device)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=reducer.flat.device)   # of the last step (device scalar)

    @torch.no_grad()
    def step(self, closure=None, max_norm: float = 0.0, pending_scale: float = 1.0):
        """max_norm > 0: clip the global gradient norm first; pending_scale: factor still owed to the gradients (1 / world
        after GradientAllReduce.finish()).  The gradient buffer itself is left as it is (unscaled)."""
        if closure is not None:
            raise NotImplementedError("FlatAdam.step takes no closure")
        g = self.param_groups[0]
        self.steps += 1
        grads = self.reducer.flat
        acc = ops.sumsq(grads, self._sumsq) if max_norm > 0 else None
        ops.adam_step(self.flat_p, grads, self.exp_avg, self.exp_avg_sq, g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                      g["weight_decay"], self.steps, acc, max_norm, pending_scale, self.grad_norm if acc is not None else None)
        # the kernel wrote through raw pointers: tell the version-keyed caches (packed weight images, compiled programs)
        torch.autograd.graph.increment_version(self.reducer.params)

    def zero_grad(self, set_to_none: bool = False):
        self.reducer.zero_grad()

    def state_dict(self):
        return {"steps": self.steps, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}

    def load_state_dict(self, sd):
        self.steps = int(sd["steps"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups[0].update(sd["param_groups"][0])
