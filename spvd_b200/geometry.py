"""Coordinate-only state of one forward pass: voxel levels, hash tables, kernel maps, point<->voxel
indices and trilinear weights -- everything the reference caches lazily in ``_caches``
(models/sparse_utils.py:71,90,122-125 and torchsparse's kmaps/cmaps) -- built up front on the GPU.

None of it depends on features, so the whole pyramid is enqueued back to back with device-side
sizes (every buffer has the point count as its upper bound) and the host reads the level sizes
ONCE.  The reference synchronises at every torch.unique / scatter_mean / kernel-map build
(SURVEY.md section 3.1).
"""
from __future__ import annotations

import torch

from . import ops


class KernelMap:
    """Offset-major map [kvol, ld] (+ per-128-row-tile offset mask) of one sparse convolution and the
    map of its data gradient."""

    __slots__ = ("map_t", "mask", "n_out", "n_in", "kvol", "grad", "flip", "_pairs", "_ws", "_gen", "n_pairs")

    def __init__(self, map_t, mask, n_out, n_in, kvol, flip):
        self.map_t, self.mask, self.n_out, self.n_in, self.kvol, self.flip = map_t, mask, n_out, n_in, kvol, flip
        self.grad = None   # KernelMap that produces d(in) from d(out)
        self._pairs = None
        self._ws, self._gen = None, 0     # PlanWorkspace whose buffers this map views, and its build generation
        self.n_pairs = 0                  # valid (in, out) pairs, when the planner counted them (pair-major kernels)

    def check_current(self):
        """Maps planned by a PlanWorkspace are VIEWS of its persistent buffers: a later build of the same workspace
        overwrites them.  Raise instead of computing gradients on somebody else's geometry."""
        if self._ws is not None and self._ws.generation != self._gen:
            raise RuntimeError("this kernel map was overwritten by a later forward pass that reused its geometry workspace "
                               "(more forwards than model.geometry_pool between a forward and its backward): raise "
                               "model.geometry_pool")

    def pairs(self):
        """pair-major view (pair_in, pair_out, tile_k, n_tiles) of this map, built on first use (the weight-gradient
        kernel reduces over it); costs one host read of the tile count."""
        if self._pairs is None:
            pin, pout, tk, nt, _ = ops.kmap_pairs(self.map_t, self.n_out, self.kvol)
            self._pairs = (pin, pout, tk, int(nt.item()))
        return self._pairs


class Level:
    def __init__(self, stride, coords, n_dev):
        self.stride, self.coords_cap, self.n_dev = stride, coords, n_dev
        self.n = None
        self.table = None
        self._kmap3_raw = None
        self._down_raw = None
        self.kmap3 = None          # KernelMap, k=3 submanifold
        self.kmap_down = None      # KernelMap, this level -> 2*stride (k=2, s=2)
        self.kmap_up = None        # KernelMap, 2*stride -> this level (transposed)
        self.batch_counts = None   # int32 [B] (device)
        self.batch_counts_host = None

    @property
    def coords(self):
        return self.coords_cap[: self.n]


class Geometry:
    def __init__(self, pc: torch.Tensor, pres: float, voxel_size: float, strides=(1, 2, 4, 8, 16),
                 k3_strides=None, p2v_strides=(1,), devox_strides=(1,), n_batch=None,
                 downsample_mode="spconv", scale_mode="mul"):
        """pc: float32 [Np,4] = (batch, x, y, z) point coordinates in units of ``pres`` (x.C.float(),
        models/spvd.py:436)."""
        if not pc.is_cuda:
            raise RuntimeError("spvd_b200.Geometry needs CUDA tensors (no CPU fallback)")
        dev = pc.device
        np_ = pc.shape[0]
        self.n_points = np_
        self.mode = downsample_mode
        k3_strides = strides if k3_strides is None else k3_strides
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.empty(max(ops.lib.spvd_unique_workspace_bytes(np_), 1), dtype=torch.uint8, device=dev)
        ld = ops.round128(np_)
        # V1: coordinates of the points at stride 1, voxel set, point -> voxel index
        self.fcoords, icoords = ops.point_voxel_coords(pc, pres, voxel_size, scale_mode)
        coords, cnt = ops.unique_coords(icoords, None, self.status, ws)
        self.levels = {}
        lv = Level(1, coords, cnt)
        lv.table = ops.HashTable(2 * np_, dev).build(coords, cnt, status=self.status)
        self.levels[1] = lv
        idx1 = ops.point_voxel_query(self.fcoords, 1, lv.table)      # == sphashquery(new_int_coord, sparse_coord)
        self.p2v = {1: idx1}
        # K2: the coarser levels
        prev = lv
        for s in strides:
            if s == 1:
                continue
            assert s == 2 * prev.stride, "strides must be consecutive powers of two"
            c2, n2 = ops.downsample_coords(prev.coords_cap, prev.n_dev, downsample_mode, self.status, ws)
            cur = Level(s, c2, n2)
            cur.table = ops.HashTable(2 * np_, dev).build(c2, n2, status=self.status)
            prev._down_raw = ops.kmap_down2(prev.coords_cap, prev.n_dev, cur.table, ld)
            self.levels[s] = cur
            prev = cur
        # K1
        for s in k3_strides:
            l = self.levels[s]
            l._kmap3_raw = ops.kmap_subm3(l.coords_cap, l.n_dev, l.table, ld)
        # V2 / V3 queries
        for s in p2v_strides:
            if s not in self.p2v:
                self.p2v[s] = ops.point_voxel_query(self.fcoords, s, self.levels[s].table)
        self.devox = {s: ops.devox_query(self.fcoords, s, self.levels[s].table) for s in devox_strides}
        # misses go to row 0 (idx_query.clamp_(0), models/sparse_utils.py:93)
        for s in self.p2v:
            self.p2v[s].clamp_(min=0)
        if n_batch is not None:
            for l in self.levels.values():
                l.batch_counts = ops.batch_counts(l.coords_cap, l.n_dev, n_batch)
        # ---- the one host synchronisation of the forward: level sizes (+ per-cloud counts, status)
        order = sorted(self.levels)
        parts = [self.levels[s].n_dev for s in order] + [self.status]
        if n_batch is not None:
            parts += [self.levels[s].batch_counts for s in order]
        host = torch.cat(parts).cpu()
        sizes = host[: len(order)].tolist()
        st = int(host[len(order)])
        if st:
            raise RuntimeError("spvd_b200: " + ("coordinate outside the packed 16-bit range; " if st & 1 else "") +
                               ("hash table full" if st & 2 else ""))
        for i, s in enumerate(order):
            l = self.levels[s]
            l.n = int(sizes[i])
            if l.n == 0:
                raise RuntimeError(f"empty voxel level at tensor stride {s} (downsample_mode={downsample_mode!r}); "
                                   "the reference crashes here too (experiments/ConditionalGeneration.ipynb:421)")
            if n_batch is not None:
                o = len(order) + 1 + i * n_batch
                l.batch_counts_host = host[o: o + n_batch].tolist()
        for s in order:
            l = self.levels[s]
            if l._kmap3_raw is not None:
                m, k = l._kmap3_raw
                l.kmap3 = KernelMap(m, k, l.n, l.n, 27, flip=True)
                l.kmap3.grad = l.kmap3
            if l._down_raw is not None:
                md, kd, mu, ku, _, _ = l._down_raw
                c = self.levels[2 * s]
                l.kmap_down = KernelMap(md, kd, c.n, l.n, 8, flip=False)
                l.kmap_up = KernelMap(mu, ku, l.n, c.n, 8, flip=False)
                l.kmap_down.grad, l.kmap_up.grad = l.kmap_up, l.kmap_down

    _pending = None          # PlanWorkspace.build(split_level=...): [collect the next group of coarser levels, ...]
    _keep = None
    _ws, _gen = None, 0      # the PlanWorkspace whose persistent buffers this geometry views, and its build generation

    def check_current(self):
        """Index tensors of a planned geometry (point -> voxel rows, devoxelize corners, kernel maps) are views of the
        workspace's buffers: raise when a later build has overwritten them (see KernelMap.check_current)."""
        if self._ws is not None and self._ws.generation != self._gen:
            raise RuntimeError("this geometry was overwritten by a later forward pass that reused its workspace (more forwards "
                               "than model.geometry_pool between a forward and its backward): raise model.geometry_pool")

    def finish(self, upto=None):
        """Wait for (and collect) level groups that are still being planned on the side stream: all of them, or just
        enough to have ``upto`` levels."""
        while self._pending and (upto is None or len(self.levels) < upto):
            self._pending.pop(0)()
        if not self._pending:
            self._pending = None
            self._keep = None
        return self

    @property
    def levels_ready(self):
        return len(self.levels)

    def level(self, s):
        if s not in self.levels:
            self.finish()
        return self.levels[s]

    # ---- derived index tensors (geometry only, cached per level) --------------------------------
    def batch_index(self, s):
        lv = self.levels[s]
        if getattr(lv, "_batch_index", None) is None:
            lv._batch_index = lv.coords[:, 0].long()
        return lv._batch_index

    def dense_layout(self, s):
        """Row -> slot map of the padded [B, Nmax] layout that the reference builds with
        torch_geometric.utils.to_dense_batch (models/modules/transformer.py:12-15,118-125); rows are
        batch-major so slot = batch * Nmax + (row - first row of the batch).  No host sync: Nmax came
        with the level sizes."""
        lv = self.levels[s]
        if getattr(lv, "_dense", None) is None:
            if lv.batch_counts is None:
                raise RuntimeError("Geometry was built without n_batch; attention needs per-cloud counts")
            counts = lv.batch_counts.long()
            nb, nmax = counts.shape[0], max(1, max(lv.batch_counts_host))
            start = torch.cumsum(counts, 0) - counts
            b = self.batch_index(s)
            pos = torch.arange(lv.n, device=b.device) - start[b]
            mask = torch.arange(nmax, device=b.device)[None, :] < counts[:, None]
            lv._dense = (b * nmax + pos, mask, nb, nmax)
        return lv._dense

    def batch_offsets(self, s):
        """int32 [B+1]: first row of every cloud at tensor stride s (rows are batch-major)."""
        lv = self.levels[s]
        if getattr(lv, "_offsets", None) is None:
            if lv.batch_counts is None:
                raise RuntimeError("Geometry was built without n_batch; attention needs per-cloud counts")
            off = torch.zeros(lv.batch_counts.shape[0] + 1, dtype=torch.int32, device=lv.batch_counts.device)
            off[1:] = torch.cumsum(lv.batch_counts, 0)
            lv._offsets = off
        return lv._offsets


# --------------------------------------------------------------------------------------------------
# One-call planner (spvd_plan_build): persistent buffers + a single C call per forward
# --------------------------------------------------------------------------------------------------
import ctypes as _C
import os
import time

from ._C import check as _check, lib as _lib


class _PlanLevel(_C.Structure):
    _fields_ = [(n, _C.c_void_p) for n in ("coords", "count", "keys", "vals", "kmap3", "mask3", "kmap_down", "mask_down",
                                           "kmap_up", "mask_up", "p2v", "devox_idx", "devox_w", "batch_counts",
                                           "batch_offsets", "pair_in", "pair_out", "pair_tile_k", "pair_counts",
                                           "pair_scratch", "dpair_in", "dpair_out", "dpair_tile_k", "dpair_counts",
                                           "p2v_w", "p2v_cnt", "p2v_start", "p2v_cursor", "p2v_order", "p2v_units",
                                           "kmap3_sorted", "mask3_sorted", "perm3", "kmap_down_sorted", "mask_down_sorted",
                                           "perm_down", "kmap_up_sorted", "mask_up_sorted", "perm_up", "sort_ws")] + \
               [("capacity", _C.c_int32), ("pair_tile_cap", _C.c_int32), ("dpair_tile_cap", _C.c_int32),
                ("sort_ws_bytes", _C.c_int32)]


class _Table:
    """view of a workspace hash table with the ops.HashTable attribute surface"""

    def __init__(self, keys, vals):
        self.keys, self.vals, self.capacity = keys, vals, keys.shape[0]

    query = ops.HashTable.query


class PlanWorkspace:
    """All buffers of one Geometry for a fixed (n_points, n_batch, plan), allocated once; ``build`` fills them
    with one C call and returns a Geometry that VIEWS them (valid until the workspace is built again)."""

    def __init__(self, n_points, n_batch, strides, k3_strides, p2v_strides, devox_strides, device, pair_strides=None,
                 sort_strides=()):
        self.n_points, self.n_batch, self.strides = n_points, n_batch, tuple(strides)
        # mask-sorted views of the 3^3 maps (spvd_kmap_sort) for the persistent sorted-tile convolution of the executor
        self.sort_strides = tuple(s for s in sort_strides if s in k3_strides)
        # pair lists at every level: the executor picks pair-major or dense per level from the pair count, the training path's
        # weight gradient reduces over them
        self.pair_strides = tuple(s for s in (k3_strides if pair_strides is None else pair_strides) if s in k3_strides)
        assert all(s == 1 << i for i, s in enumerate(self.strides)), "strides must be 1,2,4,..."
        i32 = dict(dtype=torch.int32, device=device)
        nl, ld = len(self.strides), ops.round128(n_points)
        self.ld = ld
        self.fcoords = torch.empty((n_points, 4), dtype=torch.float32, device=device)
        self.icoords = torch.empty((n_points, 4), **i32)
        self.n_summary = nl + 1 + nl * n_batch + 4 * nl      # level counts | status | per-cloud counts | (pair tiles, pairs) x2
        self.summary = torch.zeros(self.n_summary, **i32)
        self.summary_host = torch.zeros(self.n_summary, dtype=torch.int32).pin_memory()
        # one host copy of the summary per level group of a split build
        self.summary_hosts = [self.summary_host] + [torch.zeros(self.n_summary, dtype=torch.int32).pin_memory() for _ in range(nl - 1)]
        self._graphs, self.pc_stage = {}, None
        self.use_graphs = os.environ.get("SPVD_PLAN_GRAPH", "1") != "0"
        self.group_done = [torch.cuda.Event() for _ in range(nl)]
        self._events_live = False
        self.trace = None            # list -> build() appends (label, perf_counter) marks (tools/step_breakdown.py)
        self.ws = torch.empty(max(_lib.spvd_unique_workspace_bytes(n_points), 256), dtype=torch.uint8, device=device)
        self.c_levels = (_PlanLevel * nl)()
        self.bufs = []
        for i, s in enumerate(self.strides):
            b = dict(coords=torch.empty((n_points, 4), **i32), keys=torch.empty(2 * n_points, dtype=torch.int64, device=device),
                     vals=torch.empty(2 * n_points, **i32), count=self.summary[i: i + 1],
                     batch_counts=self.summary[nl + 1 + i * n_batch: nl + 1 + (i + 1) * n_batch],
                     batch_offsets=torch.zeros(n_batch + 1, **i32))
            if s in k3_strides:
                b["kmap3"], b["mask3"] = torch.empty((27, ld), **i32), torch.empty(ld // 128, **i32)
            if s in self.sort_strides:
                b["kmap3_sorted"], b["mask3_sorted"] = torch.empty((27, ld), **i32), torch.empty(ld // 128, **i32)
                b["perm3"] = torch.empty(ld, **i32)
                b["sort_ws"] = torch.empty(_lib.spvd_kmap_sort_workspace_bytes(n_points), dtype=torch.uint8, device=device)
            if i + 1 < nl:
                b["kmap_down"], b["mask_down"] = torch.empty((8, ld), **i32), torch.empty(ld // 128, **i32)
                b["kmap_up"], b["mask_up"] = torch.empty((8, ld), **i32), torch.empty(ld // 128, **i32)
            if s in p2v_strides:
                b["p2v"] = torch.empty(n_points, **i32)
                b["p2v_w"] = torch.empty(n_points, dtype=torch.float32, device=device)   # 1 / points per voxel, per point
                b["p2v_cnt"] = torch.empty(n_points + 2, **i32)
                if s > 1:           # many points per voxel: segmented reduction instead of RED (stride 1 has ~1 point per voxel)
                    for name in ("p2v_start", "p2v_cursor", "p2v_order"):
                        b[name] = torch.empty(n_points, **i32)
                    b["p2v_units"] = torch.empty(2 * (n_points + n_points // 32 + 1), **i32)
            if s in devox_strides:
                b["devox_idx"] = torch.empty((n_points, 8), **i32)
                b["devox_w"] = torch.empty((n_points, 8), dtype=torch.float32, device=device)
            tile_cap = 0
            if s in self.pair_strides:
                tile_cap = 27 * ld // 128
                b["pair_in"], b["pair_out"] = torch.empty(tile_cap * 128, **i32), torch.empty(tile_cap * 128, **i32)
                b["pair_tile_k"] = torch.empty(tile_cap, **i32)
                o = nl + 1 + nl * n_batch + 2 * i
                b["pair_counts"] = self.summary[o: o + 2]
            dcap = 0
            if i + 1 < nl:                                   # pair-major view of the stride-2 map (down and transposed convs)
                dcap = 8 * ld // 128
                b["dpair_in"], b["dpair_out"] = torch.empty(dcap * 128, **i32), torch.empty(dcap * 128, **i32)
                b["dpair_tile_k"] = torch.empty(dcap, **i32)
                o = nl + 1 + nl * n_batch + 2 * nl + 2 * i
                b["dpair_counts"] = self.summary[o: o + 2]
            # (every level: the stride-2 pair list of level i - 1 is built in the scratch of level i)
            b["pair_scratch"] = torch.empty(32 + 27 * ((n_points + 4095) // 4096 + 1), **i32)
            self.bufs.append(b)
            L = self.c_levels[i]
            for name, t in b.items():
                setattr(L, name, t.data_ptr())
            L.capacity = 2 * n_points
            L.sort_ws_bytes = b["sort_ws"].numel() if "sort_ws" in b else 0
            L.pair_tile_cap = tile_cap
            L.dpair_tile_cap = dcap
        self.status = self.summary[nl: nl + 1]
        # kernels launched by one build (ops.stats accounting)
        self.n_kernels = (1 + 22 + 23 * (nl - 1) + nl + (nl - 1) + len(k3_strides) + 2 * len(p2v_strides) + len(devox_strides) +
                          2 * nl + 3 * len(self.pair_strides) + 3 * (nl - 1) + 13 * len(self.sort_strides))

    def build(self, pc, pres, voxel_size, scale_mode="mul", downsample_mode="spconv", max_cloud_rows=0, split_level=0,
              side_stream=None, defer=False):
        """max_cloud_rows > 0: the points are batch-major with at most that many per cloud (what torch2sparse
        produces) -> per-cloud shared-memory sorts; 0: no assumption (global radix sort).
        split_level = k > 0 (or increasing cuts (k0, k1, ..)) with a side_stream: only the k finest levels are planned
        (and waited for) here; the coarser groups are planned on side_stream while the caller already works on the fine
        ones, and ``Geometry.finish()`` (or the first access to a coarser level) collects them.
        defer: do not wait for anything here -- the planner is only enqueued (on the current stream) and the caller collects
        with ``Geometry.finish()`` before the first use: planning the NEXT batch on a side stream while this one trains."""
        if not pc.is_cuda or pc.dtype != torch.float32 or not pc.is_contiguous() or pc.shape != (self.n_points, 4):
            raise RuntimeError("PlanWorkspace.build needs a contiguous CUDA float32 [n_points,4] tensor")
        nl = len(self.strides)

        ops.stats.launches += self.n_kernels if not max_cloud_rows else self.n_kernels - 20 * nl
        self.generation = getattr(self, "generation", 0) + 1
        geo = Geometry.__new__(Geometry)
        geo._ws, geo._gen = self, self.generation
        geo.n_points, geo.mode, geo.status = self.n_points, downsample_mode, self.status
        geo.fcoords, geo.levels, geo.p2v, geo.devox, geo.p2v_w, geo.p2v_groups = self.fcoords, {}, {}, {}, {}, {}
        cur = torch.cuda.current_stream()
        cuts = (split_level,) if isinstance(split_level, int) else tuple(split_level)
        cuts = sorted({c for c in cuts if 0 < c < nl}) if side_stream is not None else []
        # level groups [0,c0) [c0,c1) ... : one C call enqueues everything before the host waits -- the finest group on
        # the current stream, the others on the side stream behind an event -- so the GPU never waits for the host
        bounds = [0] + cuts + [nl]
        ng = len(bounds) - 1
        if not self._events_live:
            for e in self.group_done:
                e.record(cur)                                   # materialises the cudaEvent_t handles
            self._events_live = True
        c_events = (_C.c_void_p * ng)(*[e.cuda_event for e in self.group_done[:ng]])
        side = side_stream.cuda_stream if side_stream is not None else cur.cuda_stream
        key = (int(max_cloud_rows), float(pres), float(voxel_size), scale_mode, downsample_mode, tuple(bounds))
        graph = self._graphs.get(key) if self.use_graphs else None
        if graph is None and self.use_graphs and key not in self._graphs:
            graph = self._graphs[key] = self._capture(key)
        if graph is not None:
            _check(_lib.spvd_plan_graph_launch(graph, pc.data_ptr(), c_events, cur.cuda_stream, side))
        else:
            c_bounds = (_C.c_int32 * (ng + 1))(*bounds)
            c_hosts = (_C.c_void_p * ng)(*[h.data_ptr() for h in self.summary_hosts[:ng]])
            _check(_lib.spvd_plan_build_groups(
                pc.data_ptr(), self.n_points, self.n_batch, int(max_cloud_rows), pres, voxel_size, ops._SCALE[scale_mode],
                ops._DOWNSAMPLE[downsample_mode], self.c_levels, nl, c_bounds, ng, self.ld, self.fcoords.data_ptr(),
                self.icoords.data_ptr(), self.status.data_ptr(), self.summary.data_ptr(), c_hosts, self.n_summary, c_events,
                self.ws.data_ptr(), self.ws.numel(), cur.cuda_stream, side))
        if self.trace is not None:
            self.trace.append(("plan enqueued", time.perf_counter()))
        pending = []
        for g in range(1, ng):
            def collect(g=g, lo=bounds[g], hi=bounds[g + 1]):
                self.group_done[g].synchronize()
                self._collect(geo, lo, hi, self.summary_hosts[g])
            pending.append(collect)
        if defer:
            def collect0(hi=bounds[1]):
                self.group_done[0].synchronize()
                self._collect(geo, 0, hi)
            geo._pending = [collect0] + pending
            geo._keep = pc                      # the planner reads it on this stream until then
            return geo
        self.group_done[0].synchronize()
        if self.trace is not None:
            self.trace.append(("finest group done", time.perf_counter()))
        self._collect(geo, 0, bounds[1])
        if self.trace is not None:
            self.trace.append(("finest group collected", time.perf_counter()))
        geo._pending = pending or None
        return geo

    def _capture(self, key):
        """The planner for one (mode, level grouping) as CUDA graphs (spvd_plan_graph_create); None if capture fails."""
        max_cloud_rows, pres, voxel_size, scale_mode, downsample_mode, bounds = key
        ng, nl = len(bounds) - 1, len(self.strides)
        if self.pc_stage is None:
            self.pc_stage = torch.empty((self.n_points, 4), dtype=torch.float32, device=self.fcoords.device)
        c_bounds = (_C.c_int32 * (ng + 1))(*bounds)
        c_hosts = (_C.c_void_p * ng)(*[h.data_ptr() for h in self.summary_hosts[:ng]])
        handle = _C.c_void_p()
        torch.cuda.synchronize()
        rc = _lib.spvd_plan_graph_create(self.n_points, self.n_batch, max_cloud_rows, pres, voxel_size, ops._SCALE[scale_mode],
                                         ops._DOWNSAMPLE[downsample_mode], self.c_levels, nl, c_bounds, ng, self.ld,
                                         self.pc_stage.data_ptr(), self.fcoords.data_ptr(), self.icoords.data_ptr(),
                                         self.status.data_ptr(), self.summary.data_ptr(), c_hosts, self.n_summary,
                                         self.ws.data_ptr(), self.ws.numel(), _C.byref(handle))
        if rc != 0:
            import warnings
            warnings.warn("spvd_b200: planner graph capture failed (" + _lib.spvd_last_error().decode() + "); launching directly")
            return None
        return handle

    def __del__(self):
        for g in getattr(self, "_graphs", {}).values():
            if g is not None:
                try:
                    _lib.spvd_plan_graph_destroy(g)
                except Exception:
                    pass

    def _collect(self, geo, lo, hi, summary=None):
        downsample_mode = geo.mode
        host = (self.summary_host if summary is None else summary).tolist()
        nl, nb = len(self.strides), self.n_batch
        st = host[nl]
        if st:
            raise RuntimeError("spvd_b200: " + ("coordinate outside the packed 16-bit range; " if st & 1 else "") +
                               ("hash table full; " if st & 2 else "") +
                               ("a cloud has more points than the max_cloud_rows hint (or the points are not batch-major)"
                                if st & 4 else ""))
        for i in range(lo, hi):
            s = self.strides[i]
            b = self.bufs[i]
            lv = Level(s, b["coords"], b["count"])
            lv.n = host[i]
            if lv.n == 0:
                raise RuntimeError(f"empty voxel level at tensor stride {s} (downsample_mode={downsample_mode!r}); "
                                   "the reference crashes here too (experiments/ConditionalGeneration.ipynb:421)")
            lv.table = _Table(b["keys"], b["vals"])
            lv.batch_counts = b["batch_counts"]
            lv.batch_counts_host = host[nl + 1 + i * nb: nl + 1 + (i + 1) * nb]
            lv._offsets = b["batch_offsets"]
            if "pair_in" in b:
                o = nl + 1 + nl * nb + 2 * i
                lv.pairs = (b["pair_in"], b["pair_out"], b["pair_tile_k"], host[o], host[o + 1])   # (.., tiles, pairs)
            geo.levels[s] = lv
            if "p2v" in b:
                geo.p2v[s] = b["p2v"]
                geo.p2v_w[s] = b["p2v_w"]
                if "p2v_order" in b:
                    geo.p2v_groups[s] = (b["p2v_cnt"], b["p2v_start"], b["p2v_order"], b["p2v_units"])
            if "devox_idx" in b:
                geo.devox[s] = (b["devox_idx"], b["devox_w"])
            if "kmap3" in b:
                lv.kmap3 = KernelMap(b["kmap3"], b["mask3"], lv.n, lv.n, 27, flip=True)
                lv.kmap3.grad = lv.kmap3
                lv.kmap3._ws, lv.kmap3._gen = self, geo._gen
                if "kmap3_sorted" in b:
                    lv.kmap3_sorted = (b["kmap3_sorted"], b["mask3_sorted"], b["perm3"])
                if "pair_in" in b:       # the weight-gradient kernel reduces over the planner's pair list: no lazy build, no sync
                    lv.kmap3._pairs = lv.pairs[:4]
                    lv.kmap3.n_pairs = lv.pairs[4]
        for i in range(max(lo - 1, 0), hi - 1):           # stride-2 maps between level i and i + 1 (planned with level i + 1)
            b, lv, c = self.bufs[i], geo.levels[self.strides[i]], geo.levels[self.strides[i + 1]]
            if "dpair_in" in b:
                o = nl + 1 + nl * nb + 2 * nl + 2 * i
                lv.dpairs = (b["dpair_in"], b["dpair_out"], b["dpair_tile_k"], host[o], host[o + 1])
            if "kmap_down" in b:
                lv.kmap_down = KernelMap(b["kmap_down"], b["mask_down"], c.n, lv.n, 8, flip=False)
                lv.kmap_up = KernelMap(b["kmap_up"], b["mask_up"], lv.n, c.n, 8, flip=False)
                lv.kmap_down.grad, lv.kmap_up.grad = lv.kmap_up, lv.kmap_down
                for km in (lv.kmap_down, lv.kmap_up):
                    km._ws, km._gen = self, geo._gen
                if "dpair_in" in b:      # one (fine, coarse) pair list serves the strided conv and, roles swapped, its transpose
                    dp = lv.dpairs
                    lv.kmap_down._pairs = (dp[0], dp[1], dp[2], dp[3])
                    lv.kmap_up._pairs = (dp[1], dp[0], dp[2], dp[3])
                    lv.kmap_down.n_pairs = lv.kmap_up.n_pairs = dp[4]
