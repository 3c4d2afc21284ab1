"""Install CPU stand-ins for the reference's un-vendored dependencies so that the reference's OWN
Python (models/*.py, utils/schedulers.py) can be imported and executed in the build container.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Used by oracle/make_golden.py (golden-vector
generation) and by the CPU tests that check the oracle's restatement of models/sparse_utils.py
against the reference file itself.  /root/reference does not exist on the GPU box, so nothing that
runs there calls ``install()``.
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as tnn

from . import tsparse as ts


# ---- torch_geometric stand-ins (models/modules/graph.py:4,8; models/modules/transformer.py:5,14)
class MessagePassing(tnn.Module):
    """Minimal PyG MessagePassing: enough for TimeEmbeddingBlock (flow='target_to_source',
    aggr='add'): message(x_i, t_j) with i = edge_index[0], j = edge_index[1], summed at i."""

    def __init__(self, aggr="add", flow="source_to_target"):
        super().__init__()
        assert aggr == "add"
        self.flow = flow

    def propagate(self, edge_index, **kw):
        i, j = (0, 1) if self.flow == "target_to_source" else (1, 0)
        args, n = {}, None
        import inspect
        for name in inspect.signature(self.message).parameters:
            base, sel = name[:-2], name[-2:]
            src = kw[base]
            args[name] = src[edge_index[i if sel == "_i" else j]]
            if sel == "_i":
                n = src.shape[0]
        msg = self.message(**args)
        return torch.zeros((n,) + msg.shape[1:], dtype=msg.dtype, device=msg.device).index_add_(0, edge_index[i], msg)


def to_dense_batch(x, batch):
    """PyG to_dense_batch for an ascending batch vector: [B, Nmax, C] zero padded + bool mask."""
    nb = int(batch.max()) + 1
    cnt = torch.bincount(batch, minlength=nb)
    nmax = int(cnt.max())
    start = torch.cumsum(cnt, 0) - cnt
    pos = torch.arange(x.shape[0], device=x.device) - start[batch]
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


def _noop(x=None, *a, **k):
    return x


def _store_attr(*a, **k):  # fastcore.store_attr is only used by out-of-scope callbacks
    raise NotImplementedError


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__spvd_oracle_shim__ = True
    sys.modules[name] = m
    return m


def install():
    """Register the stand-in modules (idempotent).  Refuses to shadow a real installation."""
    for real in ("torchsparse", "torch_scatter"):
        if real in sys.modules and not getattr(sys.modules[real], "__spvd_oracle_shim__", False):
            raise RuntimeError(f"a real {real} is already imported")
    devox = _module("torchsparse.nn.functional.devoxelize", calc_ti_weights=ts.calc_ti_weights,
                    spdevoxelize=ts.spdevoxelize)
    fn = _module("torchsparse.nn.functional", spdevoxelize=ts.spdevoxelize, conv3d=ts.conv3d,
                 calc_ti_weights=ts.calc_ti_weights, devoxelize=devox)
    nnutils = _module("torchsparse.nn.utils", __all__=[])
    nn = _module("torchsparse.nn", Conv3d=ts.Conv3d, BatchNorm=ts.BatchNorm, SiLU=ts.SiLU,
                 functional=fn, utils=nnutils)
    tcache = _module("torchsparse.utils.tensor_cache", TensorCache=ts.TensorCache)
    quant = _module("torchsparse.utils.quantize", sparse_quantize=ts.sparse_quantize, ravel_hash=ts.ravel_hash)
    coll = _module("torchsparse.utils.collate", sparse_collate_fn=ts.sparse_collate_fn,
                   sparse_collate=ts.sparse_collate)
    utils = _module("torchsparse.utils", make_ntuple=ts.make_ntuple, make_tensor=ts.make_tensor,
                    __all__=["make_ntuple", "make_tensor"], tensor_cache=tcache, quantize=quant, collate=coll)
    backend = _module("torchsparse.backend", GPUHashTable=ts.GPUHashTable)
    _module("torchsparse", SparseTensor=ts.SparseTensor, cat=ts.cat, nn=nn, utils=utils,
            backend=backend, __version__="2.1.0+spvd_oracle")
    _module("torch_scatter", scatter_mean=ts.scatter_mean)
    pyg_nn = _module("torch_geometric.nn", MessagePassing=MessagePassing)
    pyg_utils = _module("torch_geometric.utils", to_dense_batch=to_dense_batch)
    _module("torch_geometric", nn=pyg_nn, utils=pyg_utils)
    fc_all = _module("fastcore.all", noop=_noop, store_attr=_store_attr)
    _module("fastcore", all=fc_all)
    layers = _module("timm.models.layers", DropPath=DropPath)
    models = _module("timm.models", layers=layers)
    _module("timm", models=models)


def import_reference(path="/root/reference"):
    """install() + put the reference checkout on sys.path; returns its ``models`` package."""
    install()
    if path not in sys.path:
        sys.path.insert(0, path)
    import importlib
    return importlib.import_module("models")
