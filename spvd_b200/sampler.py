"""Diffusion sampling loop around the denoiser (the caller of the hot path for BASELINE config 2).

Mirrors the public surface of the reference's schedulers (utils/schedulers.py:64-145,172-278:
``DDPMSparseSchedulerGPU(...).sample(model, bs, n_points)``, ``sample_step``, ``torch2sparse``,
``create_noise``, ``update_rule``) with the per-step host work removed: the DDPM coefficients are
plain floats computed once, the noise is drawn on the device (``rng='cpu'`` reproduces the
reference's CPU draw + H2D copy, utils/schedulers.py:81), and re-sparsification is a handful of
sync-free device ops (``order='keep'``; ``order='reference'`` reproduces the reference's
(batch, ravel-hash) row order and first-occurrence de-duplication, models/sparse_utils.py:217-258,
at the price of one torch.unique sync).
"""
from __future__ import annotations

import math

import torch

from .schedule import DDIM, DDPM  # noqa: F401  (re-exported: the reference keeps them in utils/schedulers.py)
from .sparse import SparseTensor


def torch2sparse(pts: torch.Tensor, pres: float = 1e-5, order: str = "keep") -> SparseTensor:
    """[B,N,F] points (xyz first) -> SparseTensor(coords int32 [M,4] = (b, floor((p - min_cloud p)/pres)), feats).
    Reference: SparseSchedulerGPU.torch2sparse, utils/schedulers.py:249-257.  On CUDA with order='keep' this is the
    fused two-kernel path of libspvd_b200 (spvd_ddpm_step without an update)."""
    B, N, Fd = pts.shape
    if order == "keep" and pts.is_cuda and pts.dtype == torch.float32:
        from . import ops
        feats = pts.reshape(-1, Fd).contiguous()
        _, coords, pc = ops.ddpm_step(feats, B, N, pres)
        st = SparseTensor(feats, coords)
        st.points_per_cloud, st.coords_float = N, pc
        return st
    c = pts[..., :3]
    c = c - c.amin(dim=1, keepdim=True)
    vs = torch.full((3,), pres, device=pts.device, dtype=pts.dtype)
    ci = torch.floor(c / vs).int()
    b = torch.arange(B, device=pts.device, dtype=torch.int32).repeat_interleave(N)
    coords = torch.cat([b.view(-1, 1), ci.view(-1, 3)], dim=1)
    feats = pts.reshape(-1, Fd)
    if order == "keep":
        st = SparseTensor(feats.contiguous(), coords.contiguous())
        st.points_per_cloud = N          # batch-major, exactly N rows per cloud: lets the planner sort per cloud
        return st
    if order != "reference":
        raise ValueError(order)
    x = (ci - ci.amin(dim=1, keepdim=True)).long()                       # batched_ravel_hash_torch
    xmax = x.amax(dim=1, keepdim=True) + 1
    h = (x[..., 0] * xmax[..., 1] + x[..., 1]) * xmax[..., 2] + x[..., 2]
    key = torch.stack([b.long(), h.reshape(-1)], dim=1)
    _, inv = torch.unique(key, dim=0, sorted=True, return_inverse=True)
    perm = torch.arange(inv.shape[0], device=pts.device)
    first = inv.new_full((int(inv.max()) + 1,), inv.shape[0]).scatter_reduce_(0, inv, perm, "amin")
    st = SparseTensor(feats[first].contiguous(), coords[first].contiguous())
    st.points_per_cloud = N              # rows are (batch, hash)-sorted, at most N per cloud
    return st


class SparseScheduler:
    def __init__(self, strategy, pres=1e-5, save_process=False, rng="cuda", order="keep"):
        self.strategy, self.pres, self.save_process, self.rng, self.order = strategy, pres, save_process, rng, order

    def _noise(self, device):
        if self.rng == "cpu":
            return lambda shape: torch.randn(shape).to(device)
        return lambda shape: torch.randn(shape, device=device)

    def torch2sparse(self, pts, shape):
        return torch2sparse(pts.reshape(shape), self.pres, self.order)

    def create_noise(self, shape, device):
        noise = self._noise(device)(shape).clamp_(-3, 3)                  # utils/schedulers.py:197-201
        return self.torch2sparse(noise, shape)

    def get_pc(self, x_t, shape):
        return x_t.F.detach().cpu().reshape(shape)

    def update_rule(self, x_t, noise_pred, t, i, shape, device):
        s = self.strategy
        if (self.order == "keep" and type(s) is DDPM and x_t.F.is_cuda and x_t.F.shape[1] == shape[2]
                and x_t.F.shape[0] == shape[0] * shape[1]):
            # fused: DDPM update + per-cloud min shift + quantisation in two kernels
            from . import ops
            z = self._noise(device)(x_t.F.shape) if t > 0 else None
            new, coords, pc = ops.ddpm_step(x_t.F.contiguous(), shape[0], shape[1], self.pres, noise_pred.contiguous(), z,
                                            s.c1[t], s.c2[t], s.sig[t])
            st = SparseTensor(new, coords)
            st.points_per_cloud, st.coords_float = shape[1], pc
            return st
        new = s.update(x_t.F, noise_pred, t, i, self._noise(device))
        return self.torch2sparse(new, shape)

    def sample_step(self, model, x_t, t, i, emb, shape, device):
        t_batch = torch.full((shape[0],), t, device=device, dtype=torch.long)
        noise_pred = model((x_t, t_batch)) if emb is None else model((x_t, t_batch, emb))   # emb = class ids [B]
        return self.update_rule(x_t, noise_pred, t, i, shape, device)

    @torch.no_grad()
    def sample(self, model, bs, n_points=2048, nf=3, emb=None, save_process=False):
        device = next(model.parameters()).device
        shape = (bs, n_points, nf)
        emb = torch.full((bs,), emb, dtype=torch.long, device=device) if emb is not None else None
        x_t = self.create_noise(shape, device)
        preds = [self.get_pc(x_t, shape)] if save_process else None
        for i, t in enumerate(self.strategy.steps):
            x_t = self.sample_step(model, x_t, t, i, emb, shape, device)
            if save_process:
                preds.append(self.get_pc(x_t, shape))
        return preds if save_process else self.get_pc(x_t, shape)


class DDPMSparseSchedulerGPU(SparseScheduler):
    def __init__(self, beta_min=0.0001, beta_max=0.02, n_steps=1000, mode="linear", sigma="bt", pres=1e-5,
                 save_process=False, rng="cuda", order="keep"):
        super().__init__(DDPM(beta_min, beta_max, n_steps, mode, sigma), pres, save_process, rng, order)


class DDIMSparseSchedulerGPU(SparseScheduler):
    def __init__(self, beta_min=0.0001, beta_max=0.02, n_steps=1000, s_steps=100, s_mode="linear", mode="linear",
                 pres=1e-5, save_process=False, rng="cuda", order="keep"):
        super().__init__(DDIM(beta_min, beta_max, n_steps, mode, s_steps, s_mode), pres, save_process, rng, order)


class DDPMSparseCompletionSchedulerGPU(SparseScheduler):
    """Masked (inpainting-style) sampling for completion / super-resolution: the caller of BASELINE config 5.
    Same surface as the reference (utils/completion_schedulers.py:7-93: ``complete(x0, model, n_points)``,
    ``pad_noise``, ``update_rule``): points carry a 4th 0/1 channel, known points (mask 1) are kept, the rest follow the
    DDPM update of the first three predicted channels."""

    def __init__(self, beta_min=0.0001, beta_max=0.02, n_steps=1000, mode="linear", sigma="bt", pres=1e-5,
                 save_process=False, rng="cuda", order="keep"):
        super().__init__(DDPM(beta_min, beta_max, n_steps, mode, sigma), pres, save_process, rng, order)

    def pad_noise(self, x0, n_points):
        B, N, Fd = x0.shape
        noise = self._noise(x0.device)((B, n_points - N, Fd))
        mask = torch.zeros(B, n_points, 1, device=x0.device)
        mask[:, :N] = 1
        padded = torch.cat([torch.cat([x0, noise], dim=1), mask], dim=-1)
        return self.torch2sparse(padded, padded.shape)

    def update_rule(self, x_t, noise_pred, t, i, shape, device):
        mask = x_t.F[:, -1:].bool()
        new = self.strategy.update(x_t.F[:, :3], noise_pred[:, :3], t, i, self._noise(device))
        x = torch.where(mask, x_t.F[:, :3], new)
        return self.torch2sparse(torch.cat([x, mask.float()], dim=-1), shape)

    def get_pc(self, x_t, shape):
        return x_t.F.detach().cpu().reshape(shape)[:, :, :3]

    @torch.no_grad()
    def complete(self, x0, model, n_points, emb=None, save_process=False):
        device = next(model.parameters()).device
        x0 = x0.to(device)
        shape = (x0.shape[0], n_points, x0.shape[-1] + 1)
        x_t = self.pad_noise(x0, n_points)
        preds = [self.get_pc(x_t, shape)] if save_process else None
        for i, t in enumerate(self.strategy.steps):
            x_t = self.sample_step(model, x_t, t, i, emb, shape, device)
            if save_process:
                preds.append(self.get_pc(x_t, shape))
        return preds if save_process else self.get_pc(x_t, shape)
