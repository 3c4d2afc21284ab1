"""Containers with the reference's attribute surface (torchsparse.SparseTensor as constructed at
models/sparse_utils.py:70,95 and utils/schedulers.py:257; PointTensor models/sparse_utils.py:23-33)."""
from __future__ import annotations

from itertools import repeat


def make_ntuple(x, ndim=3):
    if isinstance(x, (int, float)):
        return tuple(repeat(x, ndim))
    x = tuple(x)
    assert len(x) == ndim
    return x


class TensorCache:
    def __init__(self):
        self.cmaps, self.kmaps, self.hashmaps = {}, {}, {}
        self.geometry = None        # spvd_b200.geometry.Geometry shared by every tensor derived from one input


class SparseTensor:
    def __init__(self, feats, coords, stride=1, spatial_range=None):
        self.feats, self.coords = feats, coords
        self.stride = make_ntuple(stride, 3)
        self.spatial_range = spatial_range
        self.coords_float = None    # optional float32 copy of coords (saves the x.C.float() of models/spvd.py:436)
        self.points_per_cloud = 0   # optional hint: rows are batch-major with at most this many per cloud (0 = unknown)
        self._caches = TensorCache()

    F = property(lambda self: self.feats, lambda self, v: setattr(self, "feats", v))
    C = property(lambda self: self.coords, lambda self, v: setattr(self, "coords", v))
    s = property(lambda self: self.stride, lambda self, v: setattr(self, "stride", make_ntuple(v, 3)))

    def to(self, device, non_blocking=True):
        self.feats = self.feats.to(device, non_blocking=non_blocking)
        self.coords = self.coords.to(device, non_blocking=non_blocking)
        return self

    def cuda(self):
        return self.to("cuda")

    def cpu(self):
        return self.to("cpu")

    def detach(self):
        self.feats = self.feats.detach()
        return self


class PointTensor(SparseTensor):
    def __init__(self, feats, coords, stride=1):
        super().__init__(feats, coords, stride)
        self._caches.idx_query = {}
        self._caches.idx_query_devox = {}
        self._caches.weights_devox = {}
