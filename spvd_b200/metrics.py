"""Evaluation distances on the GPU with the reference's call surface.

``ChamferFunction`` / ``ChamferDistanceL2`` mirror metrics/chamfer_dist/__init__.py:13-45, ``chamfer_3DDist`` the
four-output form used by metrics/evaluation_metrics.py:13,46, ``earth_mover_distance`` metrics/PyTorchEMD/emd.py:5-47.
CUDA only: there is no CPU fallback."""
from __future__ import annotations

import ctypes

import torch

from ._C import check, lib


def _ptr(t):
    if not t.is_cuda or not t.is_contiguous() or t.dtype not in (torch.float32, torch.int32):
        raise RuntimeError("spvd_b200.metrics needs contiguous CUDA float32 / int32 tensors (there is no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _clouds(xyz1, xyz2):
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[2] != 3 or xyz2.shape[2] != 3 or xyz1.shape[0] != xyz2.shape[0]:
        raise ValueError("expected xyz1 [B,n,3] and xyz2 [B,m,3]")
    return xyz1.contiguous().float(), xyz2.contiguous().float()


class ChamferFunction(torch.autograd.Function):
    """(dist1, dist2[, idx1, idx2]): squared distance of every point to its nearest neighbour in the other cloud"""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1, xyz2 = _clouds(xyz1, xyz2)
        B, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
        dist1 = torch.empty((B, n), dtype=torch.float32, device=xyz1.device)
        dist2 = torch.empty((B, m), dtype=torch.float32, device=xyz1.device)
        idx1 = torch.empty((B, n), dtype=torch.int32, device=xyz1.device)
        idx2 = torch.empty((B, m), dtype=torch.int32, device=xyz1.device)
        check(lib.spvd_chamfer_fwd(_ptr(xyz1), _ptr(xyz2), B, n, m, _ptr(dist1), _ptr(dist2), _ptr(idx1), _ptr(idx2), _stream()))
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, g1, g2, _g3=None, _g4=None):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        B, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
        g1 = torch.zeros((B, n), device=xyz1.device) if g1 is None else g1.contiguous().float()
        g2 = torch.zeros((B, m), device=xyz1.device) if g2 is None else g2.contiguous().float()
        gx1, gx2 = torch.empty_like(xyz1), torch.empty_like(xyz2)
        check(lib.spvd_chamfer_bwd(_ptr(xyz1), _ptr(xyz2), _ptr(idx1), _ptr(idx2), _ptr(g1), _ptr(g2), B, n, m, _ptr(gx1), _ptr(gx2),
                                   _stream()))
        return gx1, gx2


class chamfer_3DDist(torch.nn.Module):
    def forward(self, xyz1, xyz2):
        return ChamferFunction.apply(xyz1, xyz2)


class ChamferDistanceL2(torch.nn.Module):
    """mean(dist1) + mean(dist2) (metrics/chamfer_dist/__init__.py:28-45)"""

    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2):
        if xyz1.size(0) == 1 and self.ignore_zeros:
            xyz1 = xyz1[torch.sum(xyz1, dim=2).ne(0)].unsqueeze(0)
            xyz2 = xyz2[torch.sum(xyz2, dim=2).ne(0)].unsqueeze(0)
        d1, d2, _, _ = ChamferFunction.apply(xyz1, xyz2)
        return torch.mean(d1) + torch.mean(d2)


class EarthMoverDistanceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1, xyz2 = _clouds(xyz1, xyz2)
        B, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
        match = torch.empty((B, m, n), dtype=torch.float32, device=xyz1.device)
        ws = torch.empty(lib.spvd_emd_workspace_floats(B, n, m), dtype=torch.float32, device=xyz1.device)
        check(lib.spvd_emd_approxmatch(_ptr(xyz1), _ptr(xyz2), B, n, m, _ptr(match), _ptr(ws), _stream()))
        cost = torch.empty(B, dtype=torch.float32, device=xyz1.device)
        check(lib.spvd_emd_matchcost_fwd(_ptr(xyz1), _ptr(xyz2), _ptr(match), B, n, m, _ptr(cost), _stream()))
        ctx.save_for_backward(xyz1, xyz2, match)
        return cost

    @staticmethod
    def backward(ctx, grad_cost):
        xyz1, xyz2, match = ctx.saved_tensors
        B, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
        g1, g2 = torch.empty_like(xyz1), torch.empty_like(xyz2)
        check(lib.spvd_emd_matchcost_bwd(_ptr(grad_cost.contiguous().float()), _ptr(xyz1), _ptr(xyz2), _ptr(match), B, n, m,
                                         _ptr(g1), _ptr(g2), _stream()))
        return g1, g2


def earth_mover_distance(xyz1, xyz2, transpose=True):
    """approximate EMD per cloud pair -> [B] (metrics/PyTorchEMD/emd.py:26-47; transpose: inputs are [B,3,n])"""
    if xyz1.dim() == 2:
        xyz1 = xyz1.unsqueeze(0)
    if xyz2.dim() == 2:
        xyz2 = xyz2.unsqueeze(0)
    if transpose:
        xyz1, xyz2 = xyz1.transpose(1, 2), xyz2.transpose(1, 2)
    return EarthMoverDistanceFunction.apply(xyz1, xyz2)


EMD = earth_mover_distance
