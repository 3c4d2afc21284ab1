"""Thin torch-tensor wrappers over the C ABI (one function per entry point of include/spvd_b200.h).

torch is plumbing here: it owns device memory and the stream; every byte of arithmetic happens in
libspvd_b200.so.  All wrappers require CUDA tensors -- there is no CPU path in the product.
"""
from __future__ import annotations

import ctypes
import os

import torch

from ._C import check, lib

SCALE_MUL, SCALE_DIV = 0, 1
DOWNSAMPLE_SPCONV, DOWNSAMPLE_FLOOR = 0, 1
ORDER_BXYZ, ORDER_XYZB = 0, 1
CONV_AUTO, CONV_SIMT, CONV_UMMA = 0, 1, 2
STATUS_COORD_RANGE, STATUS_HASH_FULL = 1, 2

_DOWNSAMPLE = {"spconv": DOWNSAMPLE_SPCONV, "floor": DOWNSAMPLE_FLOOR}
_SCALE = {"mul": SCALE_MUL, "div": SCALE_DIV}


def _p(t, dtype=None):
    """device address of a tensor for a void* argument (a plain int: ctypes converts it; ~3000 calls per training step)"""
    if t is None:
        return None
    if not (t.is_cuda and t.is_contiguous()):
        if not t.is_cuda:
            raise RuntimeError("spvd_b200 ops need CUDA tensors (there is no CPU fallback)")
        raise RuntimeError("spvd_b200 ops need contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"expected {dtype}, got {t.dtype}")
    return t.data_ptr()


def _s():
    # torch.cuda.current_stream() costs ~15 us of Python per call (hundreds of calls per training step); the raw getter is
    # a plain C call returning the same cudaStream_t
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


class _Stats:
    """Kernel-launch accounting (bench.py's ``gpu_launches``) and optional CUDA-event timing of the
    convolution launches (bench.py's live roofline measurement)."""

    def __init__(self):
        self.launches = 0
        self.conv_events = None      # list of (start, end, tag) when enabled

    def reset(self, time_convs=False):
        self.launches = 0
        self.conv_events = [] if time_convs else None


stats = _Stats()
# kernels launched by one call of each entry point (memsets are not kernels)
_K = {"point_voxel_coords": 1, "unique_coords": 22, "downsample_coords": 23, "hash_build": 1, "hash_query": 1,
      "kmap_subm3": 1, "kmap_down2": 1, "kmap_pairs": 3, "segmented_unique": 3, "batch_counts": 1, "point_voxel_query": 1, "devox_query": 1,
      "scatter_mean_fwd": 2, "scatter_mean_bwd": 1, "devoxelize_fwd": 1, "devoxelize_bwd": 1, "pack_weights": 1,
      "transpose_weights": 1, "conv_fwd": 1, "conv_wgrad": 1, "ddpm_step": 2, "round_tf32": 1, "affine_act": 1, "film": 1, "kmap_sort": 13}


def _count(name):
    stats.launches += _K[name]


def round128(n: int) -> int:
    return max(128, (n + 127) // 128 * 128)


# ------------------------------------------------------------------------------------ coordinates
def point_voxel_coords(pc, init_res, after_res, scale_mode="mul", want_int=True):
    _count("point_voxel_coords")
    n = pc.shape[0]
    f = torch.empty_like(pc)
    i = torch.empty((n, 4), dtype=torch.int32, device=pc.device) if want_int else None
    check(lib.spvd_point_voxel_coords(_p(pc, torch.float32), n, init_res, after_res, _SCALE[scale_mode], _p(f), _p(i), _s()))
    return f, i


def unique_coords(icoords, n_dev=None, status=None, workspace=None):
    """-> (coords [n_cap,4] sorted, count int32[1]) ; sizes stay on the device"""
    _count("unique_coords")
    n_cap = icoords.shape[0]
    out = torch.empty((n_cap, 4), dtype=torch.int32, device=icoords.device)
    cnt = torch.empty(1, dtype=torch.int32, device=icoords.device)
    nbytes = lib.spvd_unique_workspace_bytes(n_cap)
    ws = workspace if workspace is not None and workspace.numel() >= nbytes else torch.empty(nbytes, dtype=torch.uint8, device=icoords.device)
    check(lib.spvd_unique_coords(_p(icoords, torch.int32), n_cap, _p(n_dev), _p(out), _p(cnt), _p(status), _p(ws),
                                 ws.numel(), _s()))
    return out, cnt


def downsample_coords(coords, n_dev=None, mode="spconv", status=None, workspace=None):
    _count("downsample_coords")
    n_cap = coords.shape[0]
    out = torch.empty((n_cap, 4), dtype=torch.int32, device=coords.device)
    cnt = torch.empty(1, dtype=torch.int32, device=coords.device)
    nbytes = lib.spvd_downsample_workspace_bytes(n_cap)
    ws = workspace if workspace is not None and workspace.numel() >= nbytes else torch.empty(nbytes, dtype=torch.uint8, device=coords.device)
    check(lib.spvd_downsample_coords(_p(coords, torch.int32), n_cap, _p(n_dev), _DOWNSAMPLE[mode], _p(out), _p(cnt),
                                     _p(status), _p(ws), ws.numel(), _s()))
    return out, cnt


def segmented_unique(coords, n_batch, max_cloud_rows, parent=False, mode="spconv", n_dev=None, status=None):
    """sorted-unique (parent=False) or stride-2 parents (parent=True) of batch-major rows, one CTA per cloud.
    -> (coords [n_cap,4], count int32[1], batch_counts int32[n_batch], batch_offsets int32[n_batch+1])"""
    _count("segmented_unique")
    n_cap = coords.shape[0]
    dev = coords.device
    out = torch.empty((n_cap, 4), dtype=torch.int32, device=dev)
    cnt = torch.empty(1, dtype=torch.int32, device=dev)
    bc = torch.empty(n_batch, dtype=torch.int32, device=dev)
    bo = torch.empty(n_batch + 1, dtype=torch.int32, device=dev)
    ws = torch.empty(lib.spvd_unique_workspace_bytes(n_cap), dtype=torch.uint8, device=dev)
    check(lib.spvd_segmented_unique(_p(coords, torch.int32), n_cap, _p(n_dev), n_batch, int(max_cloud_rows), int(parent),
                                    _DOWNSAMPLE[mode], _p(out), _p(cnt), _p(bc), _p(bo), _p(status), _p(ws), ws.numel(), _s()))
    return out, cnt, bc, bo


class HashTable:
    """keys int64[capacity] + vals int32[capacity]; values are row+1, 0 = absent."""

    def __init__(self, capacity, device):
        self.capacity = int(capacity)
        self.keys = torch.empty(self.capacity, dtype=torch.int64, device=device)
        self.vals = torch.empty(self.capacity, dtype=torch.int32, device=device)

    def build(self, coords, n_dev=None, order=ORDER_BXYZ, status=None):
        _count("hash_build")
        check(lib.spvd_hash_build(_p(coords, torch.int32), coords.shape[0], _p(n_dev), order, _p(self.keys), _p(self.vals),
                                  self.capacity, _p(status), _s()))
        return self

    def query(self, q, kernel_size=1, order=ORDER_BXYZ):
        """-> int32 [m, kernel_size^3], 1-based rows, 0 = absent"""
        _count("hash_query")
        m = q.shape[0]
        out = torch.empty((m, kernel_size ** 3), dtype=torch.int32, device=q.device)
        check(lib.spvd_hash_query(_p(q, torch.int32), m, order, kernel_size, _p(self.keys), _p(self.vals), self.capacity,
                                  _p(out), _s()))
        return out


def kmap_subm3(coords, n_dev, table: HashTable, ld=None):
    _count("kmap_subm3")
    n_cap = coords.shape[0]
    ld = ld or round128(n_cap)
    map_t = torch.empty((27, ld), dtype=torch.int32, device=coords.device)
    mask = torch.empty(ld // 128, dtype=torch.int32, device=coords.device)
    check(lib.spvd_kmap_subm3(_p(coords, torch.int32), n_cap, _p(n_dev), _p(table.keys), _p(table.vals), table.capacity,
                              _p(map_t), ld, _p(mask), _s()))
    return map_t, mask


def kmap_down2(fine_coords, n_fine_dev, coarse_table: HashTable, ld_coarse, want_parent=False):
    _count("kmap_down2")
    n_cap = fine_coords.shape[0]
    dev = fine_coords.device
    ld_fine = round128(n_cap)
    md = torch.empty((8, ld_coarse), dtype=torch.int32, device=dev)
    mu = torch.empty((8, ld_fine), dtype=torch.int32, device=dev)
    kd = torch.empty(ld_coarse // 128, dtype=torch.int32, device=dev)
    ku = torch.empty(ld_fine // 128, dtype=torch.int32, device=dev)
    parent = torch.empty(n_cap, dtype=torch.int32, device=dev) if want_parent else None
    koff = torch.empty(n_cap, dtype=torch.int32, device=dev) if want_parent else None
    check(lib.spvd_kmap_down2(_p(fine_coords, torch.int32), n_cap, _p(n_fine_dev), _p(coarse_table.keys),
                              _p(coarse_table.vals), coarse_table.capacity, _p(md), ld_coarse, _p(kd), _p(mu), ld_fine,
                              _p(ku), _p(parent), _p(koff), _s()))
    return md, kd, mu, ku, parent, koff


def batch_counts(coords, n_dev, nb):
    _count("batch_counts")
    out = torch.empty(nb, dtype=torch.int32, device=coords.device)
    check(lib.spvd_batch_counts(_p(coords, torch.int32), coords.shape[0], _p(n_dev), nb, _p(out), _s()))
    return out


def point_voxel_query(fcoords, stride, table: HashTable):
    _count("point_voxel_query")
    n = fcoords.shape[0]
    idx = torch.empty(n, dtype=torch.int32, device=fcoords.device)
    check(lib.spvd_point_voxel_query(_p(fcoords, torch.float32), n, int(stride), _p(table.keys), _p(table.vals),
                                     table.capacity, _p(idx), _s()))
    return idx


def devox_query(fcoords, stride, table: HashTable):
    _count("devox_query")
    n = fcoords.shape[0]
    idx8 = torch.empty((n, 8), dtype=torch.int32, device=fcoords.device)
    w8 = torch.empty((n, 8), dtype=torch.float32, device=fcoords.device)
    check(lib.spvd_devox_query(_p(fcoords, torch.float32), n, int(stride), _p(table.keys), _p(table.vals), table.capacity,
                               _p(idx8), _p(w8), _s()))
    return idx8, w8


# --------------------------------------------------------------------------------- point <-> voxel
def scatter_mean_fwd(src, idx, nv):
    _count("scatter_mean_fwd")
    n, c = src.shape
    out = torch.empty((nv, c), dtype=torch.float32, device=src.device)
    counts = torch.empty(nv, dtype=torch.float32, device=src.device)
    check(lib.spvd_scatter_mean_fwd(_p(src, torch.float32), _p(idx, torch.int32), n, c, _p(out), _p(counts), nv, _s()))
    return out, counts


def scatter_mean_bwd(gout, idx, counts, n):
    _count("scatter_mean_bwd")
    c = gout.shape[1]
    gsrc = torch.empty((n, c), dtype=torch.float32, device=gout.device)
    check(lib.spvd_scatter_mean_bwd(_p(gout, torch.float32), _p(idx, torch.int32), _p(counts, torch.float32), n, c,
                                    _p(gsrc), _s()))
    return gsrc


def devoxelize_fwd(feats, idx8, w8):
    _count("devoxelize_fwd")
    n, c = idx8.shape[0], feats.shape[1]
    out = torch.empty((n, c), dtype=torch.float32, device=feats.device)
    check(lib.spvd_devoxelize_fwd(_p(feats, torch.float32), _p(idx8, torch.int32), _p(w8, torch.float32), n, c, _p(out), _s()))
    return out


def devoxelize_bwd(gout, idx8, w8, nv):
    _count("devoxelize_bwd")
    n, c = gout.shape
    g = torch.empty((nv, c), dtype=torch.float32, device=gout.device)
    check(lib.spvd_devoxelize_bwd(_p(gout, torch.float32), _p(idx8, torch.int32), _p(w8, torch.float32), n, c, _p(g), nv, _s()))
    return g


# --------------------------------------------------------------------------------------- conv
def umma_supported(kvol, cin, cout):
    return bool(lib.spvd_conv_umma_supported(kvol, cin, cout))


def pack_weights(w3):
    """w3 [kvol, cin, cout] -> tensor-core tile image (same element count)"""
    _count("pack_weights")
    kvol, cin, cout = w3.shape
    wp = torch.empty(kvol * cin * cout, dtype=torch.float32, device=w3.device)
    check(lib.spvd_conv_pack_weights(_p(w3, torch.float32), kvol, cin, cout, _p(wp), _s()))
    return wp


def pack_weights_f16(w3):
    """w3 [kvol, cin, cout] -> f16 tensor-core tile image (cin padded to a multiple of 64)"""
    _count("pack_weights")
    kvol, cin, cout = w3.shape
    wp = torch.empty(lib.spvd_conv_packed_weight_count_f16(kvol, cin, cout), dtype=torch.float16, device=w3.device)
    check(lib.spvd_conv_pack_weights_f16(_p(w3, torch.float32), kvol, cin, cout, _p(wp), _s()))
    return wp


def transpose_weights(w3, flip):
    _count("transpose_weights")
    kvol, cin, cout = w3.shape
    wt = torch.empty((kvol, cout, cin), dtype=torch.float32, device=w3.device)
    check(lib.spvd_conv_transpose_weights(_p(w3, torch.float32), kvol, cin, cout, int(flip), _p(wt), _s()))
    return wt


def conv_fwd(x, w3, w_packed, map_t, mask, n_out, bias=None, impl=CONV_AUTO, bn=0, nsplit=0, a_tf32=False, residual=None, mt=0,
             out_rows=None):
    """x [n_in, cin]; w3 [kvol, cin, cout] (may be None when w_packed is given and the UMMA path applies);
    map_t [kvol, ld] or None (identity)."""
    if w3 is not None:
        kvol, cin, cout = w3.shape
    else:
        raise RuntimeError("conv_fwd needs the weight tensor for its shape")
    _count("conv_fwd")
    out = torch.empty((n_out, cout), dtype=torch.float32, device=x.device)
    ld = map_t.shape[1] if map_t is not None else 0
    ev = None
    if stats.conv_events is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    if impl == CONV_UMMA or (impl == CONV_AUTO and w_packed is not None and umma_supported(kvol, cin, cout)):
        # operand mode from the dtypes: f16 weight image -> f16 operands (x f16 or fp32 converted on the fly)
        a_mode = ((3 if x.dtype == torch.float16 else 2) if w_packed.dtype == torch.float16 else int(bool(a_tf32)))
        if x.dtype == torch.float16 and w_packed.dtype != torch.float16:
            raise RuntimeError("an f16 input needs the f16 weight image (pack_weights_f16)")
        check(lib.spvd_conv_fwd_umma_rows(_p(x), _p(w_packed), _p(map_t), ld, _p(mask), _p(out_rows), n_out, kvol,
                                          cin, cout, _p(bias), _p(residual), _p(out), int(bn), int(nsplit), a_mode, int(mt), _s()))
    else:
        check(lib.spvd_conv_fwd_simt(_p(x, torch.float32), _p(w3, torch.float32), _p(map_t), ld, _p(mask), n_out, kvol, cin,
                                     cout, _p(bias), _p(residual), _p(out), _s()))
    if ev is not None:
        ev[1].record()
        umma = impl == CONV_UMMA or (impl == CONV_AUTO and w_packed is not None and umma_supported(kvol, cin, cout))
        stats.conv_events.append((ev[0], ev[1], (kvol, cin, cout, n_out, "umma" if umma else "simt", map_t)))
    return out


def kmap_pairs(map_t, n, kvol=27, tile_cap=None):
    """-> (pair_in, pair_out, tile_k, n_tiles int32[1], n_pairs int32[1])"""
    _count("kmap_pairs")
    ld = map_t.shape[1]
    tile_cap = tile_cap or (kvol * ld // 128)
    dev = map_t.device
    pin = torch.empty(tile_cap * 128, dtype=torch.int32, device=dev)
    pout = torch.empty(tile_cap * 128, dtype=torch.int32, device=dev)
    tk = torch.empty(tile_cap, dtype=torch.int32, device=dev)
    cnt = torch.empty(2, dtype=torch.int32, device=dev)
    scratch = torch.empty(32 + kvol * ((int(n) + 4095) // 4096 + 1), dtype=torch.int32, device=dev)
    check(lib.spvd_kmap_pairs(_p(map_t, torch.int32), ld, int(n), None, kvol, tile_cap, _p(pin), _p(pout), _p(tk),
                              _p(cnt[0:1]), _p(cnt[1:2]), _p(scratch), _s()))
    return pin, pout, tk, cnt[0:1], cnt[1:2]


def conv_fwd_pairs(x, w3, w_packed, pair_in, pair_out, tile_k, n_tiles, n_out, bias=None, residual=None, a_tf32=True,
                   center=None):
    _count("conv_fwd")
    kvol, cin, cout = w3.shape
    out = torch.empty((n_out, cout), dtype=torch.float32, device=x.device)
    center = (kvol // 2 if kvol % 2 else -1) if center is None else center
    a_mode = ((3 if x.dtype == torch.float16 else 2) if w_packed.dtype == torch.float16 else int(bool(a_tf32)))
    check(lib.spvd_conv_fwd_pairs(_p(x), _p(w_packed), _p(pair_in), _p(pair_out), _p(tile_k),
                                  int(n_tiles), center, n_out, kvol, cin, cout, _p(bias), _p(residual), _p(out),
                                  a_mode, _s()))
    return out


def conv_wgrad(x, gout, map_t, n_out, kvol):
    _count("conv_wgrad")
    cin, cout = x.shape[1], gout.shape[1]
    gw = torch.empty((kvol, cin, cout), dtype=torch.float32, device=x.device)
    ld = map_t.shape[1] if map_t is not None else 0
    check(lib.spvd_conv_wgrad(_p(x, torch.float32), _p(gout, torch.float32), _p(map_t), ld, n_out, kvol, cin, cout, _p(gw), _s()))
    return gw


def conv_wgrad_umma(x, gout, pairs, n_rows, kvol):
    """tensor-core weight gradient; x / gout must be TF32-representable.  pairs = (pair_in, pair_out, tile_k, n_tiles)
    from kmap_pairs, or None for kernel size 1 (identity pairs over n_rows)."""
    _count("conv_wgrad")
    cin, cout = x.shape[1], gout.shape[1]
    gw = torch.empty((kvol, cin, cout), dtype=torch.float32, device=x.device)
    if pairs is None:
        check(lib.spvd_conv_wgrad_umma(_p(x, torch.float32), _p(gout, torch.float32), None, None, None, 0, int(n_rows), kvol,
                                       cin, cout, _p(gw), _s()))
    else:
        pin, pout, tk, nt = pairs
        check(lib.spvd_conv_wgrad_umma(_p(x, torch.float32), _p(gout, torch.float32), _p(pin), _p(pout), _p(tk), int(nt), 0,
                                       kvol, cin, cout, _p(gw), _s()))
    return gw


def round_tf32(x):
    _count("round_tf32")
    y = torch.empty_like(x)
    check(lib.spvd_round_tf32(_p(x, torch.float32), _p(y), x.numel(), _s()))
    return y


# ----------------------------------------------------------------------------------- fused glue
def affine_act(x, scale, shift, act=1):
    """act bit 0: SiLU; bit 1: round the result to TF32 (so the next conv may gather it with cp.async)"""
    _count("affine_act")
    rows, c = x.shape
    y = torch.empty_like(x)
    check(lib.spvd_bn_silu_fwd(_p(x, torch.float32), _p(scale, torch.float32), _p(shift, torch.float32), rows, c, act, _p(y), _s()))
    return y


def film(x, tt, coords):
    _count("film")
    rows, c = x.shape
    y = torch.empty_like(x)
    check(lib.spvd_film_fwd(_p(x, torch.float32), _p(tt, torch.float32), _p(coords, torch.int32), rows, c, _p(y), _s()))
    return y


# --------------------------------------------------------------------------------------- sampler
def ddpm_step(x_t, n_clouds, n_per_cloud, pres, eps=None, z=None, c1=1.0, c2=0.0, sigma=0.0):
    """fused DDPM update + re-sparsification -> (x_new [M,F], coords int32 [M,4], pc float32 [M,4])"""
    _count("ddpm_step")
    m, f = x_t.shape
    dev = x_t.device
    x_new = torch.empty_like(x_t) if eps is not None else None
    coords = torch.empty((m, 4), dtype=torch.int32, device=dev)
    pc = torch.empty((m, 4), dtype=torch.float32, device=dev)
    scratch = torch.empty(3 * n_clouds, dtype=torch.int32, device=dev)
    check(lib.spvd_ddpm_step(_p(x_t, torch.float32), _p(eps), _p(z), float(c1), float(c2), float(sigma), n_clouds, n_per_cloud,
                             f, float(pres), _p(x_new), _p(coords), _p(pc), _p(scratch), _s()))
    return (x_new if eps is not None else x_t), coords, pc


# ------------------------------------------------------------------------------- sorted tiles + fused persistent conv
class ConvSpArgs(ctypes.Structure):
    """spvd_conv_sp_args_t (include/spvd_b200.h)"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("inp", "w", "map_t", "tile_mask", "out_rows", "in2", "w2", "bias", "residual",
                                               "film", "coords", "aff_scale", "aff_shift", "out32", "out16")] + \
               [(n, ctypes.c_int32) for n in ("ld", "n_out", "kvol", "cin", "cout", "cin2", "film_ld", "act", "n_in", "pad")]


def kmap_sort(map_t, n, kvol, n_dev=None):
    """-> (map_sorted [kvol, ld], mask_sorted [ld/128], perm [ld]) : rows sorted by their active-offset bit set"""
    _count("kmap_sort")
    ld = map_t.shape[1]
    dev = map_t.device
    ms = torch.empty((kvol, ld), dtype=torch.int32, device=dev)
    tm = torch.empty(ld // 128, dtype=torch.int32, device=dev)
    perm = torch.empty(ld, dtype=torch.int32, device=dev)
    ws = torch.empty(lib.spvd_kmap_sort_workspace_bytes(int(n)), dtype=torch.uint8, device=dev)
    check(lib.spvd_kmap_sort(_p(map_t, torch.int32), ld, int(n), _p(n_dev), kvol, _p(ms), _p(tm), _p(perm), _p(ws), ws.numel(), _s()))
    return ms, tm, perm


def conv_sp(x16, w_packed, kvol, cin, cout, n_out, map_t=None, tile_mask=None, out_rows=None, bias=None, residual=None,
            in2=None, w2_packed=None, film=None, coords=None, aff=None, silu=False, want32=True, want16=False, bn=0, stages=0):
    """spvd_conv_sp_f16: persistent sorted-tile convolution with fused epilogue.  x16 / in2 f16; returns (out32, out16)."""
    _count("conv_fwd")
    dev = x16.device
    if x16.dtype != torch.float16 or w_packed.dtype != torch.float16:
        raise RuntimeError("conv_sp needs f16 rows and the f16 weight image")
    a = ConvSpArgs()
    out32 = torch.empty((n_out, cout), dtype=torch.float32, device=dev) if want32 else None
    out16 = torch.empty((n_out, cout), dtype=torch.float16, device=dev) if want16 else None
    a.inp, a.w, a.map_t, a.tile_mask, a.out_rows = _p(x16), _p(w_packed), _p(map_t), _p(tile_mask), _p(out_rows)
    a.in2, a.w2 = _p(in2), _p(w2_packed)
    a.bias, a.residual, a.coords = _p(bias), _p(residual), _p(coords)
    a.film = ctypes.c_void_p(film.data_ptr()) if film is not None else None     # may be a column slice of a wider table
    a.aff_scale, a.aff_shift = (_p(aff[0]), _p(aff[1])) if aff is not None else (None, None)
    a.out32, a.out16 = _p(out32), _p(out16)
    a.ld = map_t.shape[1] if map_t is not None else 0
    a.n_out, a.kvol, a.cin, a.cout = int(n_out), int(kvol), int(cin), int(cout)
    a.cin2 = in2.shape[1] if in2 is not None else 0
    a.film_ld = film.stride(0) if film is not None else 0
    a.act = 1 if silu else 0
    a.n_in = x16.shape[0]
    check(lib.spvd_conv_sp_f16(ctypes.byref(a), int(bn), int(stages), _s()))
    return out32, out16


# ------------------------------------------------------------------------------------ training-mode glue (csrc/train.cu)
_K.update({"bn_train_fwd": 3, "bn_train_bwd": 3, "colsum": 2, "film_bwd": 1, "layernorm_train": 1, "layernorm_bwd": 4,
           "attention_train": 1, "attention_bwd": 2, "cat_any": 1, "split": 1, "silu": 1, "add": 1, "embedding": 1,
           "timestep_embedding": 1, "mse": 2})
_TRAIN_WS = {}


def train_ws(device):
    """fp64 scratch of the per-channel reductions (zero on entry, left zero by every kernel that uses it)"""
    ws = _TRAIN_WS.get(device)
    if ws is None:
        ws = _TRAIN_WS[device] = torch.zeros(8192, dtype=torch.float64, device=device)
    return ws


def bn_train_fwd(x, gamma, beta, eps, momentum, running_mean, running_var, act=1):
    """-> y, (save_mean, save_invstd); running statistics updated in place"""
    _count("bn_train_fwd")
    rows, c = x.shape
    dev = x.device
    stats = torch.empty((4, c), dtype=torch.float32, device=dev)       # mean | invstd | scale | shift
    y = torch.empty_like(x)
    check(lib.spvd_bn_train_fwd(_p(x, torch.float32), rows, c, _p(gamma), _p(beta), float(eps), float(momentum),
                                _p(running_mean), _p(running_var), _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(stats[3]),
                                int(act), _p(y), _p(train_ws(dev)), _s()))
    return y, stats


def bn_train_bwd(x, gy, gamma, beta, stats, act=1, gg_out=None, gb_out=None):
    """gg_out / gb_out: write the parameter gradients there (a slot of the flat gradient buffer) instead of new tensors"""
    _count("bn_train_bwd")
    rows, c = x.shape
    gx = torch.empty_like(x)
    gg = gg_out if gg_out is not None else torch.empty(c, dtype=torch.float32, device=x.device)
    gb = gb_out if gb_out is not None else torch.empty(c, dtype=torch.float32, device=x.device)
    check(lib.spvd_bn_train_bwd(_p(x, torch.float32), _p(gy, torch.float32), rows, c, _p(gamma), _p(beta), _p(stats[0]),
                                _p(stats[1]), int(act), _p(gx), _p(gg), _p(gb), _p(train_ws(x.device)), _s()))
    return gx, gg, gb


def colsum(g):
    _count("colsum")
    rows, c = g.shape
    out = torch.empty(c, dtype=torch.float32, device=g.device)
    check(lib.spvd_colsum(_p(g, torch.float32), rows, c, _p(out), _p(train_ws(g.device)), _s()))
    return out


def film_bwd(x, gy, tt, offsets, max_rows):
    _count("film_bwd")
    rows, c = x.shape
    gx = torch.empty_like(x)
    gtt = torch.zeros_like(tt)
    check(lib.spvd_film_bwd(_p(x, torch.float32), _p(gy, torch.float32), _p(tt, torch.float32), tt.shape[1], _p(offsets, torch.int32),
                            tt.shape[0], int(max_rows), c, _p(gx), _p(gtt), gtt.shape[1], _s()))
    return gx, gtt


def layernorm_train_fwd(x, gamma, beta, eps):
    _count("layernorm_train")
    rows, c = x.shape
    y = torch.empty_like(x)
    st = torch.empty((2, max(rows, 1)), dtype=torch.float32, device=x.device)
    check(lib.spvd_layernorm_train_fwd(_p(x, torch.float32), _p(gamma), _p(beta), rows, c, float(eps), _p(y), _p(st[0]), _p(st[1]), _s()))
    return y, st


def layernorm_bwd(x, gy, gamma, st, gg_out=None, gb_out=None):
    _count("layernorm_bwd")
    rows, c = x.shape
    gx = torch.empty_like(x)
    gg = gg_out if gg_out is not None else torch.empty(c, dtype=torch.float32, device=x.device)
    gb = gb_out if gb_out is not None else torch.empty(c, dtype=torch.float32, device=x.device)
    check(lib.spvd_layernorm_bwd(_p(x, torch.float32), _p(gy, torch.float32), _p(gamma), _p(st[0]), _p(st[1]), rows, c, _p(gx),
                                 _p(gg), _p(gb), _p(train_ws(x.device)), _s()))
    return gx, gg, gb


def attention_train_fwd(qkv, offsets, n_batch, max_len, heads):
    _count("attention_train")
    rows, c3 = qkv.shape
    c = c3 // 3
    out = torch.empty((rows, c), dtype=torch.float32, device=qkv.device)
    lse = torch.empty((rows, heads), dtype=torch.float32, device=qkv.device)
    check(lib.spvd_attention_train_fwd(_p(qkv, torch.float32), _p(offsets, torch.int32), n_batch, int(max_len), c, heads, _p(out),
                                       _p(lse), _s()))
    return out, lse


def attention_bwd(qkv, out, gout, lse, offsets, n_batch, max_len, heads):
    _count("attention_bwd")
    rows, c3 = qkv.shape
    gqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    check(lib.spvd_attention_bwd(_p(qkv, torch.float32), _p(out, torch.float32), _p(gout, torch.float32), _p(lse), _p(offsets, torch.int32),
                                 n_batch, int(max_len), c3 // 3, heads, _p(gqkv), _p(delta), _s()))
    return gqkv


def cat_any(a, b):
    _count("cat_any")
    y = torch.empty((a.shape[0], a.shape[1] + b.shape[1]), dtype=torch.float32, device=a.device)
    check(lib.spvd_cat_any_fwd(_p(a, torch.float32), _p(b, torch.float32), a.shape[0], a.shape[1], b.shape[1], _p(y), _s()))
    return y


def split(g, ca):
    _count("split")
    rows, c = g.shape
    ga = torch.empty((rows, ca), dtype=torch.float32, device=g.device)
    gb = torch.empty((rows, c - ca), dtype=torch.float32, device=g.device)
    check(lib.spvd_split_fwd(_p(g, torch.float32), rows, ca, c - ca, _p(ga), _p(gb), _s()))
    return ga, gb


def silu_fwd(x):
    _count("silu")
    y = torch.empty_like(x)
    check(lib.spvd_silu_fwd(_p(x, torch.float32), x.numel(), _p(y), _s()))
    return y


def silu_bwd(x, gy):
    _count("silu")
    gx = torch.empty_like(x)
    check(lib.spvd_silu_bwd(_p(x, torch.float32), _p(gy, torch.float32), x.numel(), _p(gx), _s()))
    return gx


def add(a, b):
    _count("add")
    y = torch.empty_like(a)
    check(lib.spvd_add_fwd(_p(a, torch.float32), _p(b, torch.float32), a.numel(), _p(y), _s()))
    return y


def embedding_fwd(table, idx, base=None):
    _count("embedding")
    out = torch.empty((idx.shape[0], table.shape[1]), dtype=torch.float32, device=table.device)
    check(lib.spvd_embedding_fwd(_p(table, torch.float32), _p(idx, torch.int64), _p(base), idx.shape[0], table.shape[1], _p(out), _s()))
    return out


def embedding_bwd(g, idx, n_rows, out=None):
    """out: a ZEROED [n_rows, d] tensor to accumulate into (a slot of the flat gradient buffer)"""
    _count("embedding")
    gt = out if out is not None else torch.zeros((n_rows, g.shape[1]), dtype=torch.float32, device=g.device)
    check(lib.spvd_embedding_bwd(_p(g, torch.float32), _p(idx, torch.int64), idx.shape[0], g.shape[1], _p(gt), _s()))
    return gt


def timestep_embedding(t, d):
    _count("timestep_embedding")
    out = torch.empty((t.shape[0], d), dtype=torch.float32, device=t.device)
    check(lib.spvd_timestep_embedding(_p(t, torch.int64), t.shape[0], d, _p(out), _s()))
    return out


def mse_loss(pred, target, want_grad=True, grad_scale=1.0):
    """-> (loss [1], d loss / d pred or None)"""
    _count("mse")
    loss = torch.empty(1, dtype=torch.float32, device=pred.device)
    gp = torch.empty_like(pred) if want_grad else None
    check(lib.spvd_mse_loss(_p(pred, torch.float32), _p(target, torch.float32), pred.numel(), float(grad_scale), _p(loss), _p(gp),
                            _p(train_ws(pred.device)), _s()))
    return loss, gp


class ConvBwdArgs(ctypes.Structure):
    """spvd_conv_bwd_args_t (include/spvd_b200.h)"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("x", "g", "w", "map_t", "gmap_t", "gmap_mask", "pair_in", "pair_out", "pair_tile_k",
                                               "gx", "gw", "gbias", "workspace", "red_workspace", "w_packed_t")] + \
               [("workspace_bytes", ctypes.c_size_t)] + \
               [(n, ctypes.c_int32) for n in ("map_ld", "gmap_ld", "n_pair_tiles", "n_in", "n_out", "kvol", "cin", "cout", "flip",
                                              "x_is_tf32", "gw_transposed", "allow_tensor_cores", "n_pairs", "w_packed_t_flip")]


_K["conv_bwd"] = 8
_BWD_WS = {}


def pack_weights_train(w, kvol, cin, cout, linear=False, flip_t=0, want_w3=False, want_t=True):
    """One launch for everything a training step derives from a weight tensor (spvd_conv_pack_weights_train):
    -> (w3 [1, cin, cout] or None, forward tile image or None, transposed tile image or None)."""
    _count("pack_weights")
    dev = w.device
    w3 = torch.empty((kvol, cin, cout), dtype=torch.float32, device=dev) if want_w3 else None
    wp = torch.empty(kvol * cin * cout, dtype=torch.float32, device=dev) if umma_supported(kvol, cin, cout) else None
    wtp = torch.empty(kvol * cin * cout, dtype=torch.float32, device=dev) if (want_t and umma_supported(kvol, cout, cin)) else None
    if w3 is None and wp is None and wtp is None:
        return None, None, None
    check(lib.spvd_conv_pack_weights_train(_p(w, torch.float32), kvol, cin, cout, int(bool(linear)), int(flip_t), _p(w3), _p(wp),
                                           _p(wtp), _s()))
    return w3, wp, wtp


def conv_bwd(x, g, w3, kmap, n_in, n_out, want_gx, want_gw, want_gb, x_is_tf32=False, gw_transposed=False, tensor_cores=True,
             gw_out=None, gb_out=None, w_packed_t=None, w_packed_t_flip=0):
    """The whole backward of a sparse convolution / Linear in one C call (spvd_conv_bwd).  w3 [kvol, cin, cout];
    kmap = geometry.KernelMap of the forward conv or None (kernel size 1).  -> (gx, gw, gbias), None where not wanted;
    gw is [kvol, cin, cout], or [kvol, cout, cin] with gw_transposed.  gw_out / gb_out: contiguous fp32 tensors of those
    element counts to write the parameter gradients into (slots of the flat gradient buffer)."""
    _count("conv_bwd")
    kvol, cin, cout = w3.shape
    dev = g.device
    a = ConvBwdArgs()
    need = lib.spvd_conv_bwd_workspace_bytes(int(n_in), int(n_out), kvol, cin, cout)
    ws = _BWD_WS.get(dev)
    if ws is None or ws.numel() < need:             # one scratch per device, grown to the largest layer (stream-ordered reuse)
        ws = _BWD_WS[dev] = torch.empty(int(need * 1.25) + 4096, dtype=torch.uint8, device=dev)
    gx = torch.empty((n_in, cin), dtype=torch.float32, device=dev) if want_gx else None
    gw = gb = None
    if want_gw:
        gw = gw_out if gw_out is not None else torch.empty((kvol, cout, cin) if gw_transposed else (kvol, cin, cout),
                                                           dtype=torch.float32, device=dev)
        if gw.numel() != kvol * cin * cout or not gw.is_contiguous() or gw.dtype != torch.float32:
            raise ValueError("conv_bwd: gw_out must be a contiguous fp32 tensor of kvol*cin*cout elements")
    if want_gb:
        gb = gb_out if gb_out is not None else torch.empty(cout, dtype=torch.float32, device=dev)
    a.x, a.g, a.w = _p(x), _p(g, torch.float32), _p(w3, torch.float32)
    if kmap is not None:
        gm = kmap.grad
        a.map_t, a.map_ld = _p(kmap.map_t), kmap.map_t.shape[1]
        a.gmap_t, a.gmap_mask, a.gmap_ld = _p(gm.map_t), _p(gm.mask), gm.map_t.shape[1]
        a.flip = 1 if kmap.flip else 0
        if tensor_cores and ((want_gw and umma_supported(kvol, cin, cout)) or kmap.n_pairs):
            pin, pout, tk, nt = kmap.pairs()
            a.pair_in, a.pair_out, a.pair_tile_k, a.n_pair_tiles = _p(pin), _p(pout), _p(tk), int(nt)
            a.n_pairs = int(kmap.n_pairs or 0)
    a.gx, a.gw, a.gbias = _p(gx), _p(gw), _p(gb)
    a.workspace, a.workspace_bytes, a.red_workspace = _p(ws), ws.numel(), _p(train_ws(dev))
    a.n_in, a.n_out, a.kvol, a.cin, a.cout = int(n_in), int(n_out), kvol, cin, cout
    a.x_is_tf32, a.gw_transposed, a.allow_tensor_cores = int(bool(x_is_tf32)), int(bool(gw_transposed)), int(bool(tensor_cores))
    a.w_packed_t, a.w_packed_t_flip = _p(w_packed_t), int(w_packed_t_flip)
    check(lib.spvd_conv_bwd(ctypes.byref(a), _s()))
    return gx, gw, gb


# ------------------------------------------------------------------------------------------------ optimizer tail
_K.update({"sumsq": 1, "adam_step": 1})


def sumsq(flat, acc):
    """acc[0] (float64, device) = sum flat^2"""
    _count("sumsq")
    check(lib.spvd_sumsq(_p(flat, torch.float32), flat.numel(), _p(acc, torch.float64), _s()))
    return acc


def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, sumsq_acc=None, max_norm=0.0, pending_scale=1.0, norm_out=None):
    """in place over flat fp32 buffers: clip (by the norm in sumsq_acc) + scale + Adam (spvd_adam_step)"""
    _count("adam_step")
    check(lib.spvd_adam_step(_p(p, torch.float32), _p(g, torch.float32), _p(m, torch.float32), _p(v, torch.float32), p.numel(),
                             float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
                             _p(sumsq_acc, torch.float64) if sumsq_acc is not None else None, float(max_norm), float(pending_scale),
                             _p(norm_out), _s()))
