"""CPU restatement of the reference's evaluation distances -- TEST INFRASTRUCTURE ONLY.

Chamfer: metrics/chamfer_dist/chamfer.cu:15-171 (nearest neighbour in squared L2, both directions, with argmin) and its
backward :173-229.  EMD: metrics/PyTorchEMD/cuda/emd_kernel.cu:26-157 (approxmatch: auction-style soft matching over the
temperature levels -4^7 .. -4^-1, 0), :170-209 (matchcost, squared distances) and :251-330 (its gradients).  The reference
ships these as CUDA-only extensions without assertions, so there are no golden vectors: parity unpinned,
the arithmetic below is the documented contract.  Plain torch float32 on the CPU.
"""
import torch


def _sqdist(a, b):
    d = a[:, :, None, :] - b[:, None, :, :]
    return (d * d).sum(-1)                                   # [B, N, M]


def chamfer(xyz1, xyz2):
    d = _sqdist(xyz1.float(), xyz2.float())
    dist1, idx1 = d.min(dim=2)
    dist2, idx2 = d.min(dim=1)
    return dist1, dist2, idx1.int(), idx2.int()


def approxmatch(xyz1, xyz2):
    """-> match [B, M, N] (the reference's layout: match[b, l, k] pairs xyz2[l] with xyz1[k])"""
    xyz1, xyz2 = xyz1.float(), xyz2.float()
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    multi_l, multi_r = (1.0, float(n // m)) if n >= m else (float(m // n), 1.0)
    d = _sqdist(xyz1, xyz2)                                  # [B, n, m]
    match = torch.zeros(B, m, n)
    remain_l = torch.full((B, n), multi_l)
    remain_r = torch.full((B, m), multi_r)
    for j in range(7, -3, -1):
        level = 0.0 if j == -2 else -(4.0 ** j)
        w = torch.exp(level * d)                             # [B, n, m]
        suml = 1e-9 + (w * remain_r[:, None, :]).sum(2)
        ratio_l = remain_l / suml
        sumr = (w * ratio_l[:, :, None]).sum(1) * remain_r
        consumption = torch.clamp(remain_r / (sumr + 1e-9), max=1.0)
        ratio_r = consumption * remain_r
        remain_r = torch.clamp(remain_r - sumr, min=0.0)
        wm = w * ratio_l[:, :, None] * ratio_r[:, None, :]   # [B, n, m]
        match += wm.transpose(1, 2)
        remain_l = torch.clamp(remain_l - wm.sum(2), min=0.0)
    return match


def matchcost(xyz1, xyz2, match):
    d = _sqdist(xyz1.float(), xyz2.float())                  # [B, n, m]
    return (d * match.transpose(1, 2)).sum((1, 2))


def earth_mover_distance(xyz1, xyz2):
    return matchcost(xyz1, xyz2, approxmatch(xyz1, xyz2))
