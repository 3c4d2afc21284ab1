"""Plugin surface: the slice of the torchsparse 2.1 / torch-scatter API that the reference's hot path
imports (complete list: SURVEY.md section 8b), backed by libspvd_b200.so.

    import spvd_b200.compat; spvd_b200.compat.install()
    import models.spvd            # the reference's own, unmodified module now runs on the sm_100a kernels

``install()`` registers ``torchsparse`` (+ ``.nn``, ``.nn.functional``, ``.nn.functional.devoxelize``,
``.nn.utils``, ``.backend``, ``.utils``, ``.utils.tensor_cache``, ``.utils.quantize``, ``.utils.collate``) and
``torch_scatter`` in ``sys.modules``, plus native-torch stand-ins for the three tiny symbols the model files pull
from torch_geometric / fastcore / timm (``MessagePassing``, ``to_dense_batch``, ``noop``, ``DropPath``).  It refuses to
shadow a real installation.  This path keeps the reference's op-by-op structure (lazy kernel maps cached on
``_caches``, one host sync per dynamic size, as upstream); the fast path is ``spvd_b200.SPVUnet``.
"""
from __future__ import annotations

import math
import sys
import types

import numpy as np
import torch
import torch.nn as tnn

from .. import functional as SF
from .. import ops
from ..geometry import KernelMap
from ..sparse import SparseTensor, TensorCache, make_ntuple


def make_tensor(x, dtype=None, device=None):
    return torch.tensor(x, dtype=dtype, device=device)


# ------------------------------------------------------------------------------------- hash table
class GPUHashTable:
    """torchsparse.backend.GPUHashTable as used by sphashquery (models/sparse_utils.py:37-54): coordinates come
    as (x,y,z,b); lookups are 1-based, 0 = absent, shape [M, kernel_volume]."""

    def __init__(self, keys, vals):
        self.keys, self.vals = keys, vals
        self.capacity = int(keys.shape[0])

    def insert_coords(self, coords):
        c = coords.contiguous().int()
        ops.check(ops.lib.spvd_hash_build(ops._p(c), c.shape[0], None, ops.ORDER_XYZB, ops._p(self.keys, torch.int64),
                                          ops._p(self.vals, torch.int32), self.capacity, None, ops._s()))

    def lookup_coords(self, coords, kernel_size, stride, kernel_volume):
        ks = int(round(float(kernel_volume) ** (1.0 / 3.0)))
        q = coords.contiguous().int()
        out = torch.empty((q.shape[0], ks ** 3), dtype=torch.int32, device=q.device)
        ops.check(ops.lib.spvd_hash_query(ops._p(q), q.shape[0], ops.ORDER_XYZB, ks, ops._p(self.keys), ops._p(self.vals),
                                          self.capacity, ops._p(out), ops._s()))
        return out


# --------------------------------------------------------------------------------------- functional
def scatter_mean(src, index, dim=0, dim_size=None):
    """torch_scatter.scatter_mean(src, index, dim=0); output rows = index.max()+1 like upstream (one host sync)."""
    assert dim == 0
    idx = index.reshape(-1).int().contiguous()
    nv = int(idx.max()) + 1 if dim_size is None else int(dim_size)
    return SF.scatter_mean(src, idx, nv)


def spdevoxelize(feats, idx, weights):
    return SF.devoxelize(feats, idx.int().contiguous(), weights.float().contiguous())


def calc_ti_weights(coords, idx_query, scale=1):
    """torchsparse.nn.functional.devoxelize.calc_ti_weights (call site models/sparse_utils.py:114)."""
    with torch.no_grad():
        p = coords.float()
        f = torch.floor(p / scale) * scale if scale != 1 else torch.floor(p)
        c = f + scale
        x, y, z = p[:, 0:1], p[:, 1:2], p[:, 2:3]
        xf, yf, zf, xc, yc, zc = f[:, 0:1], f[:, 1:2], f[:, 2:3], c[:, 0:1], c[:, 1:2], c[:, 2:3]
        w = torch.cat([(xc - x) * (yc - y) * (zc - z), (xc - x) * (yc - y) * (z - zf), (xc - x) * (y - yf) * (zc - z),
                       (xc - x) * (y - yf) * (z - zf), (x - xf) * (yc - y) * (zc - z), (x - xf) * (yc - y) * (z - zf),
                       (x - xf) * (y - yf) * (zc - z), (x - xf) * (y - yf) * (z - zf)], dim=1)
        if scale != 1:
            w = w / scale ** 3
        w[idx_query == -1] = 0
        w = w / (w.sum(dim=1, keepdim=True) + 1e-8)
    return w


def _like(x, feats):
    out = SparseTensor(feats, x.coords, x.stride, x.spatial_range)
    out._caches = x._caches
    return out


def cat(inputs):
    return _like(inputs[0], torch.cat([t.feats for t in inputs], dim=1))


class Config:
    downsample_mode = "spconv"


conf = Config()


def _level_state(x):
    """per tensor-stride cache on the shared _caches: hash table, kernel maps, coarse coordinates"""
    st = x._caches.kmaps.setdefault(("spvd_b200", x.stride), {})
    if "table" not in st:
        c = x.coords.contiguous().int()
        st["coords"] = c
        st["table"] = ops.HashTable(2 * max(c.shape[0], 1), c.device).build(c)
    return st


def conv3d(x, weight, kernel_size, bias=None, stride=1, dilation=1, transposed=False):
    ks, s = make_ntuple(kernel_size)[0], make_ntuple(stride)[0]
    cache = getattr(weight, "_spvd_packed", None)
    if cache is None:
        cache = SF.PackedWeight()
        try:
            weight._spvd_packed = cache
        except Exception:
            pass
    if ks == 1 and s == 1:
        return _like(x, SF.sparse_conv(x.feats, weight, bias, None, cache))
    st = _level_state(x)
    n = st["coords"].shape[0]
    if not transposed and s == 1:
        assert ks == 3, "only kernel size 3 submanifold convolutions are on the path"
        if "k3" not in st:
            m, k = ops.kmap_subm3(st["coords"], None, st["table"])
            km = KernelMap(m, k, n, n, 27, flip=True)
            km.grad = km
            st["k3"] = km
        return _like(x, SF.sparse_conv(x.feats, weight, bias, st["k3"], cache))
    assert (ks, s) == (2, 2), "only kernel 2 / stride 2 strided convolutions are on the path"
    if not transposed:
        if "down" not in st:
            coarse, cnt = ops.downsample_coords(st["coords"], None, conf.downsample_mode)
            nc = int(cnt.item())
            if nc == 0:
                raise RuntimeError(f"empty voxel level at tensor stride {tuple(a * 2 for a in x.stride)}")
            coarse = coarse[:nc].contiguous()
            table = ops.HashTable(2 * nc, coarse.device).build(coarse)
            md, kd, mu, ku, _, _ = ops.kmap_down2(st["coords"], None, table, ops.round128(nc))
            down, up = KernelMap(md, kd, nc, n, 8, False), KernelMap(mu, ku, n, nc, 8, False)
            down.grad, up.grad = up, down
            st["down"], st["up"], st["coarse"] = down, up, coarse
            out_stride = tuple(a * 2 for a in x.stride)
            x._caches.cmaps.setdefault(x.stride, (st["coords"], x.spatial_range))
            x._caches.cmaps[out_stride] = (coarse, None)
            cs = x._caches.kmaps.setdefault(("spvd_b200", out_stride), {})
            cs.setdefault("coords", coarse)
            cs.setdefault("table", table)
        out = SparseTensor(SF.sparse_conv(x.feats, weight, bias, st["down"], cache), st["coarse"],
                           tuple(a * 2 for a in x.stride))
        out._caches = x._caches
        return out
    fine_stride = tuple(a // 2 for a in x.stride)
    fs = x._caches.kmaps[("spvd_b200", fine_stride)]
    out = SparseTensor(SF.sparse_conv(x.feats, weight, bias, fs["up"], cache), fs["coords"], fine_stride)
    out._caches = x._caches
    return out


class Conv3d(tnn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, bias=False,
                 transposed=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = make_ntuple(kernel_size), make_ntuple(stride), make_ntuple(dilation)
        self.transposed = transposed
        kv = int(np.prod(self.kernel_size))
        std = 1.0 / math.sqrt((out_channels if transposed else in_channels) * kv)
        shape = (kv, in_channels, out_channels) if kv > 1 else (in_channels, out_channels)
        self.kernel = tnn.Parameter(torch.empty(shape).uniform_(-std, std))
        self.bias = tnn.Parameter(torch.empty(out_channels).uniform_(-std, std)) if bias else None

    def forward(self, x):
        return conv3d(x, self.kernel, self.kernel_size, self.bias, self.stride, self.dilation, self.transposed)


class BatchNorm(tnn.BatchNorm1d):
    def forward(self, x):
        return _like(x, super().forward(x.feats))


class SiLU(tnn.SiLU):
    def forward(self, x):
        return _like(x, super().forward(x.feats))


# ------------------------------------------------------------ numpy-side input format (CPU, datasets / CPU scheduler)
from ..data import collate as _collate, ravel_hash, sparse_collate, sparse_quantize  # noqa: E402,F401


def sparse_collate_fn(inputs):
    return _collate(inputs) if isinstance(inputs[0], dict) else inputs


# ------------------------------------------------------------------------ tiny third-party names
class MessagePassing(tnn.Module):
    """Enough of torch_geometric.nn.MessagePassing for TimeEmbeddingBlock (models/modules/graph.py:8-41):
    aggr='add', messages built from ``x_i`` / ``t_j`` style arguments."""

    def __init__(self, aggr="add", flow="source_to_target"):
        super().__init__()
        self.flow = flow

    def propagate(self, edge_index, **kw):
        import inspect
        i, j = (0, 1) if self.flow == "target_to_source" else (1, 0)
        args, n = {}, None
        for name in inspect.signature(self.message).parameters:
            src = kw[name[:-2]]
            args[name] = src[edge_index[i if name.endswith("_i") else j]]
            if name.endswith("_i"):
                n = src.shape[0]
        msg = self.message(**args)
        return torch.zeros((n,) + msg.shape[1:], dtype=msg.dtype, device=msg.device).index_add_(0, edge_index[i], msg)


def to_dense_batch(x, batch):
    nb = int(batch.max()) + 1
    cnt = torch.bincount(batch, minlength=nb)
    nmax = int(cnt.max())
    pos = torch.arange(x.shape[0], device=x.device) - (torch.cumsum(cnt, 0) - cnt)[batch]
    dense = x.new_zeros((nb, nmax) + tuple(x.shape[1:]))
    mask = torch.zeros((nb, nmax), dtype=torch.bool, device=x.device)
    dense[batch, pos] = x
    mask[batch, pos] = True
    return dense, mask


class DropPath(tnn.Module):
    def __init__(self, p=0.0):
        super().__init__()
        self.p = p

    def forward(self, x):
        if self.p == 0.0 or not self.training:
            return x
        keep = torch.rand((x.shape[0],) + (1,) * (x.dim() - 1), device=x.device) >= self.p
        return x * keep / (1 - self.p)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__spvd_b200_compat__ = True
    sys.modules[name] = m
    return m


def _chamfer_forward(xyz1, xyz2):
    from .. import metrics
    with torch.no_grad():
        return metrics.ChamferFunction.apply(xyz1, xyz2)


def _chamfer_backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2):
    from .. import metrics as m
    B, n, mm = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    xyz1, xyz2 = xyz1.contiguous().float(), xyz2.contiguous().float()
    gx1, gx2 = torch.empty_like(xyz1), torch.empty_like(xyz2)
    m.check(m.lib.spvd_chamfer_bwd(m._ptr(xyz1), m._ptr(xyz2), m._ptr(idx1.int().contiguous()), m._ptr(idx2.int().contiguous()),
                                   m._ptr(grad_dist1.contiguous().float()), m._ptr(grad_dist2.contiguous().float()), B, n, mm,
                                   m._ptr(gx1), m._ptr(gx2), m._stream()))
    return gx1, gx2


def _emd_approxmatch(xyz1, xyz2):
    from .. import metrics as m
    xyz1, xyz2 = xyz1.contiguous().float(), xyz2.contiguous().float()
    B, n, mm = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    match = torch.empty((B, mm, n), dtype=torch.float32, device=xyz1.device)
    ws = torch.empty(m.lib.spvd_emd_workspace_floats(B, n, mm), dtype=torch.float32, device=xyz1.device)
    m.check(m.lib.spvd_emd_approxmatch(m._ptr(xyz1), m._ptr(xyz2), B, n, mm, m._ptr(match), m._ptr(ws), m._stream()))
    return match


def _emd_matchcost_forward(xyz1, xyz2, match):
    from .. import metrics as m
    xyz1, xyz2 = xyz1.contiguous().float(), xyz2.contiguous().float()
    B, n, mm = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    cost = torch.empty(B, dtype=torch.float32, device=xyz1.device)
    m.check(m.lib.spvd_emd_matchcost_fwd(m._ptr(xyz1), m._ptr(xyz2), m._ptr(match.contiguous()), B, n, mm, m._ptr(cost), m._stream()))
    return cost


def _emd_matchcost_backward(grad_cost, xyz1, xyz2, match):
    from .. import metrics as m
    xyz1, xyz2 = xyz1.contiguous().float(), xyz2.contiguous().float()
    B, n, mm = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    g1, g2 = torch.empty_like(xyz1), torch.empty_like(xyz2)
    m.check(m.lib.spvd_emd_matchcost_bwd(m._ptr(grad_cost.contiguous().float()), m._ptr(xyz1), m._ptr(xyz2), m._ptr(match.contiguous()),
                                         B, n, mm, m._ptr(g1), m._ptr(g2), m._stream()))
    return g1, g2


def install(third_party=True):
    for real in ("torchsparse", "torch_scatter"):
        if real in sys.modules and not getattr(sys.modules[real], "__spvd_b200_compat__", False):
            raise RuntimeError(f"a different {real} is already imported; spvd_b200.compat will not shadow it")
    devox = _module("torchsparse.nn.functional.devoxelize", calc_ti_weights=calc_ti_weights, spdevoxelize=spdevoxelize)
    fn = _module("torchsparse.nn.functional", spdevoxelize=spdevoxelize, conv3d=conv3d, calc_ti_weights=calc_ti_weights,
                 devoxelize=devox)
    nnutils = _module("torchsparse.nn.utils", __all__=[])
    nn = _module("torchsparse.nn", Conv3d=Conv3d, BatchNorm=BatchNorm, SiLU=SiLU, functional=fn, utils=nnutils)
    tcache = _module("torchsparse.utils.tensor_cache", TensorCache=TensorCache)
    quant = _module("torchsparse.utils.quantize", sparse_quantize=sparse_quantize, ravel_hash=ravel_hash)
    coll = _module("torchsparse.utils.collate", sparse_collate_fn=sparse_collate_fn, sparse_collate=sparse_collate)
    utils = _module("torchsparse.utils", make_ntuple=make_ntuple, make_tensor=make_tensor,
                    __all__=["make_ntuple", "make_tensor"], tensor_cache=tcache, quantize=quant, collate=coll)
    backend = _module("torchsparse.backend", GPUHashTable=GPUHashTable)
    _module("torchsparse", SparseTensor=SparseTensor, cat=cat, nn=nn, utils=utils, backend=backend, conf=conf,
            __version__="2.1.0+spvd_b200")
    _module("torch_scatter", scatter_mean=scatter_mean)
    # the reference's two compiled evaluation extensions (metrics/chamfer_dist/__init__.py:10 `import chamfer`,
    # metrics/PyTorchEMD/emd.py:2 `import emd_cuda`): its own autograd wrappers then run on libspvd_b200
    if "chamfer" not in sys.modules:
        _module("chamfer", forward=_chamfer_forward, backward=_chamfer_backward)
    if "emd_cuda" not in sys.modules:
        _module("emd_cuda", approxmatch_forward=_emd_approxmatch, matchcost_forward=_emd_matchcost_forward,
                matchcost_backward=_emd_matchcost_backward)
    if third_party:
        for name, attrs in (("torch_geometric.nn", dict(MessagePassing=MessagePassing)),
                            ("torch_geometric.utils", dict(to_dense_batch=to_dense_batch))):
            if name not in sys.modules:
                _module(name, **attrs)
        if "torch_geometric" not in sys.modules:
            _module("torch_geometric", nn=sys.modules["torch_geometric.nn"], utils=sys.modules["torch_geometric.utils"])
        if "fastcore" not in sys.modules:
            fc = _module("fastcore.all", noop=lambda x=None, *a, **k: x)
            _module("fastcore", all=fc)
        if "timm" not in sys.modules:
            layers = _module("timm.models.layers", DropPath=DropPath)
            _module("timm", models=_module("timm.models", layers=layers))
