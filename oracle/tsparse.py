"""CPU stand-in for the slice of the torchsparse 2.1.0 / torch-scatter API that the reference's
hot path calls (complete list: SURVEY.md section 8b).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Semantics are [UPSTREAM] restatements (torchsparse is not in /root/reference); the call sites
they serve are cited per symbol.  All arithmetic is in oracle/ops.py.
"""
from __future__ import annotations

import math
from itertools import repeat
from typing import Tuple, Union

import numpy as np
import torch
import torch.nn as tnn

from . import ops


class Config:
    downsample_mode = "spconv"      # 'spconv' (torchsparse 2.1 default, boundary drop) | 'floor'
    scale_mode = "mul"              # fp32 recipe of initial_voxelize, see ops.scale_point_coords
    emulate_tf32 = False            # round conv operands to TF32 like the tensor-core path


conf = Config()


def make_ntuple(x, ndim: int = 3):
    if isinstance(x, (int, float)):
        return tuple(repeat(x, ndim))
    x = tuple(x)
    assert len(x) == ndim
    return x


def make_tensor(x, dtype=None, device=None):
    return torch.tensor(x, dtype=dtype, device=device)


class TensorCache:
    """[UPSTREAM] torchsparse.utils.tensor_cache.TensorCache (models/sparse_utils.py:10)."""

    def __init__(self):
        self.cmaps, self.kmaps, self.hashmaps = {}, {}, {}


class SparseTensor:
    """[UPSTREAM] torchsparse.SparseTensor: feats [N,C] fp32, coords int32 [N,4]=(b,x,y,z), stride
    3-tuple, shared ``_caches`` (constructed at models/sparse_utils.py:70,95; utils/schedulers.py:257)."""

    def __init__(self, feats, coords, stride: Union[int, Tuple[int, ...]] = 1, spatial_range=None):
        self.feats, self.coords = feats, coords
        self.stride = make_ntuple(stride, 3)
        self.spatial_range = spatial_range
        self._caches = TensorCache()

    F = property(lambda s: s.feats, lambda s, v: setattr(s, "feats", v))
    C = property(lambda s: s.coords, lambda s, v: setattr(s, "coords", v))
    s = property(lambda s: s.stride, lambda s, v: setattr(s, "stride", make_ntuple(v, 3)))

    def to(self, device, non_blocking=True):
        self.feats, self.coords = self.feats.to(device), self.coords.to(device)
        return self

    def cpu(self):
        return self.to("cpu")

    def cuda(self):
        return self.to("cuda")

    def detach(self):
        self.feats, self.coords = self.feats.detach(), self.coords.detach()
        return self


def _like(x: SparseTensor, feats) -> SparseTensor:
    out = SparseTensor(feats, x.coords, x.stride, x.spatial_range)
    out._caches = x._caches
    return out


def cat(inputs):
    """[UPSTREAM] torchsparse.cat (models/spvd.py:232): channel concat, metadata of the first."""
    return _like(inputs[0], torch.cat([t.feats for t in inputs], dim=1))


# ------------------------------------------------------------------------------------ functional
def scatter_mean(src, index, dim=0, dim_size=None):
    assert dim == 0
    return ops.scatter_mean(src, index, dim_size)


def spdevoxelize(feats, idx, w):
    return ops.spdevoxelize(feats, idx, w)


calc_ti_weights = ops.calc_ti_weights


def _np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def conv3d(x: SparseTensor, weight, kernel_size, bias=None, stride=1, dilation=1, transposed=False):
    """[UPSTREAM] torchsparse.nn.functional.conv3d, the four shapes the path uses
    (models/spvd.py:35,38,113,229; models/ddpm_unet_attn.py:28).  Kernel maps and coordinate sets
    are cached per (tensor_stride, kernel_size, stride, dilation) on the shared _caches
    (frame at experiments/ConditionalGeneration.ipynb:449)."""
    ks, st, dl = make_ntuple(kernel_size), make_ntuple(stride), make_ntuple(dilation)
    assert dl == (1, 1, 1) and len(set(ks)) == 1 and len(set(st)) == 1
    k, s = ks[0], st[0]
    tf32 = conf.emulate_tf32
    if k == 1 and s == 1:
        w = weight if weight.dim() == 2 else weight[0]
        f = (ops.round_tf32(x.feats) @ ops.round_tf32(w)) if tf32 else x.feats @ w
        return _like(x, f if bias is None else f + bias)
    caches = x._caches
    if not transposed:
        key = (x.stride, ks, st, dl)
        if s == 1:
            assert k == 3, "only k=3 submanifold convs are on the path"
            if key not in caches.kmaps:
                caches.kmaps[key] = {"out_in_map": ops.kmap_subm(_np(x.coords), 3), "coords": x.coords}
            km = caches.kmaps[key]
            f = ops.conv_gather_gemm_scatter(x.feats, weight, km["out_in_map"], x.coords.shape[0], bias, tf32)
            return _like(x, f)
        assert (k, s) == (2, 2), "only k=2,s=2 strided convs are on the path"
        out_stride = tuple(a * s for a in x.stride)
        if key not in caches.kmaps:
            coarse, parent, koff = ops.downsample_coords(_np(x.coords), conf.downsample_mode)
            caches.kmaps[key] = {"out_in_map": ops.kmap_down(coarse.shape[0], parent, koff),
                                 "parent": parent, "koff": koff,
                                 "coords": torch.from_numpy(coarse).to(x.coords.device)}
            caches.cmaps.setdefault(x.stride, (x.coords, x.spatial_range))
            caches.cmaps[out_stride] = (caches.kmaps[key]["coords"], None)
        km = caches.kmaps[key]
        n_out = km["coords"].shape[0]
        if n_out == 0:
            raise RuntimeError(f"empty voxel level at tensor stride {out_stride}")
        f = ops.conv_gather_gemm_scatter(x.feats, weight, km["out_in_map"], n_out, bias, tf32)
        out = SparseTensor(f, km["coords"], out_stride)
        out._caches = caches
        return out
    # transposed k=2,s=2: output coordinates = cached fine coordinates, map = down map transposed
    assert (k, s) == (2, 2)
    fine_stride = tuple(a // s for a in x.stride)
    km = caches.kmaps[(fine_stride, ks, st, dl)]
    fine_coords = caches.cmaps[fine_stride][0]
    parent, koff = km["parent"], km["koff"]
    n_fine = fine_coords.shape[0]
    m = np.full((n_fine, 8), -1, dtype=np.int32)
    rows = np.nonzero(parent >= 0)[0]
    m[rows, koff[rows]] = parent[rows]
    f = ops.conv_gather_gemm_scatter(x.feats, weight, m, n_fine, bias, tf32)
    out = SparseTensor(f, fine_coords, fine_stride)
    out._caches = caches
    return out


# --------------------------------------------------------------------------------------- modules
class Conv3d(tnn.Module):
    """[UPSTREAM] torchsparse.nn.Conv3d: parameters ``kernel`` [K,Cin,Cout] ([Cin,Cout] when K=1)
    and ``bias`` [Cout] (default bias=False -- proved by the reference's parameter counts,
    SURVEY.md section 3.1); init U(+-1/sqrt(K*Cin)) (Cout when transposed)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1,
                 bias=False, transposed=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = make_ntuple(kernel_size), make_ntuple(stride), make_ntuple(dilation)
        self.transposed = transposed
        kv = int(np.prod(self.kernel_size))
        self.kernel = tnn.Parameter(torch.zeros(kv, in_channels, out_channels) if kv > 1
                                    else torch.zeros(in_channels, out_channels))
        self.bias = tnn.Parameter(torch.zeros(out_channels)) if bias else None
        std = 1.0 / math.sqrt((out_channels if transposed else in_channels) * kv)
        self.kernel.data.uniform_(-std, std)
        if self.bias is not None:
            self.bias.data.uniform_(-std, std)

    def forward(self, x):
        return conv3d(x, self.kernel, self.kernel_size, self.bias, self.stride, self.dilation, self.transposed)


class BatchNorm(tnn.BatchNorm1d):
    """[UPSTREAM] torchsparse.nn.BatchNorm (models/spvd.py:36; models/ddpm_unet_attn.py:26)."""

    def forward(self, x):
        return _like(x, super().forward(x.feats))


class SiLU(tnn.SiLU):
    def forward(self, x):
        return _like(x, super().forward(x.feats))


# ------------------------------------------------------------------------------------- hash table
class GPUHashTable:
    """[UPSTREAM] torchsparse.backend.GPUHashTable as used by sphashquery
    (models/sparse_utils.py:37-54): coords arrive as (x,y,z,b); lookups are 1-based, 0 = absent."""

    def __init__(self, keys, vals):
        self.keys, self.vals, self._target = keys, vals, None

    def insert_coords(self, coords_xyzb):
        self._target = _np(coords_xyzb)[:, [3, 0, 1, 2]]

    def lookup_coords(self, coords_xyzb, kernel_size, stride, kernel_volume):
        ks = int(round(float(kernel_volume) ** (1.0 / 3.0)))
        q = _np(coords_xyzb)[:, [3, 0, 1, 2]]
        r = ops.hash_query(q, self._target, ks) + 1
        return torch.from_numpy(r).to(coords_xyzb.device)


# ------------------------------------------------------------------------- numpy-side data format
def ravel_hash(x: np.ndarray) -> np.ndarray:
    x = x - x.min(axis=0)
    x = x.astype(np.uint64)
    xmax = x.max(axis=0).astype(np.uint64) + 1
    h = np.zeros(x.shape[0], dtype=np.uint64)
    for k in range(x.shape[1] - 1):
        h += x[:, k]
        h *= xmax[k + 1]
    return h + x[:, -1]


def sparse_quantize(coords, voxel_size=1, *, return_index=False, return_inverse=False):
    """[UPSTREAM] torchsparse.utils.quantize.sparse_quantize (utils/schedulers.py:231)."""
    vs = np.asarray(make_ntuple(voxel_size, 3))
    coords = np.floor(coords / vs).astype(np.int32)
    _, ind, inv = np.unique(ravel_hash(coords), return_index=True, return_inverse=True)
    out = [coords[ind]]
    if return_index:
        out.append(ind)
    if return_inverse:
        out.append(inv)
    return out[0] if len(out) == 1 else out


def sparse_collate(inputs):
    coords, feats = [], []
    for k, x in enumerate(inputs):
        b = torch.full((x.coords.shape[0], 1), k, dtype=torch.int)
        coords.append(torch.cat((b, x.coords.int()), dim=1))
        feats.append(x.feats)
    return SparseTensor(torch.cat(feats, 0), torch.cat(coords, 0), inputs[0].stride)


def sparse_collate_fn(inputs):
    """[UPSTREAM] torchsparse.utils.collate.sparse_collate_fn (utils/schedulers.py:238)."""
    if isinstance(inputs[0], dict):
        out = {}
        for name in inputs[0].keys():
            v = [d[name] for d in inputs]
            if isinstance(v[0], SparseTensor):
                out[name] = sparse_collate(v)
            elif isinstance(v[0], torch.Tensor):
                out[name] = torch.stack(v, 0)
            else:
                out[name] = v
        return out
    return inputs
