"""Primitive ops of the SPVD hot path, restated on the CPU (numpy for integer work, torch-CPU
for floating point so that autograd provides the backward oracle).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Each function cites the reference line it
follows; ``[UPSTREAM]`` marks semantics of torchsparse 2.1.0 / torch-scatter 2.1.2 that are not
in /root/reference (restated from their published behaviour; parity unpinned).

Conventions (they are the contract the CUDA kernels are tested against):
  * coordinates are int32 ``[N,4] = (batch, x, y, z)`` (models/sparse_utils.py:44,62,83);
  * every voxel set is sorted lexicographically by (batch, x, y, z)  (torch.unique(dim=0),
    models/sparse_utils.py:66; [UPSTREAM] spdownsample sorts a ravelled key);
  * kernel offsets: odd kernel volume -> x fastest, even -> z fastest ([UPSTREAM]
    torchsparse.nn.utils.get_kernel_offsets);  cross-correlation: neighbour = out*stride + offset;
  * the stride-1 voxel coordinate of an integer point coordinate ``c`` is
    ``floor(fl32(fl32(c * fl32(init_res)) * fl32(1/fl32(after_res))))`` -- what torch computes ON
    CUDA for ``(z.C[:,1:] * init_res) / after_res`` (models/sparse_utils.py:61-65; the reference
    path only runs on CUDA, models/sparse_utils.py:43).  ``scale_mode='div'`` selects the
    torch-CPU flavour ``fl32(fl32(c*init_res)/after_res)``.
"""
from __future__ import annotations

import numpy as np
import torch

COORD_BIAS = 1 << 15          # 16-bit fields, signed range [-32767, 32767]
COORD_MAX = COORD_BIAS - 1


# ----------------------------------------------------------------------------------------------
# integer work
# ----------------------------------------------------------------------------------------------
def pack_keys(coords) -> np.ndarray:
    """(b,x,y,z) int32 -> uint64-orderable int64 key; key order == lexicographic order."""
    c = np.asarray(coords).astype(np.int64)
    if c.size and (np.abs(c[:, 1:]).max() > COORD_MAX or c[:, 0].min() < 0 or c[:, 0].max() > 0xFFFF):
        raise ValueError("coordinate out of the 16-bit packed range")
    return (c[:, 0] << 48) | ((c[:, 1] + COORD_BIAS) << 32) | ((c[:, 2] + COORD_BIAS) << 16) | (c[:, 3] + COORD_BIAS)


def unpack_keys(keys) -> np.ndarray:
    k = np.asarray(keys).astype(np.int64)
    out = np.empty((k.shape[0], 4), dtype=np.int32)
    out[:, 0] = (k >> 48) & 0xFFFF
    out[:, 1] = ((k >> 32) & 0xFFFF) - COORD_BIAS
    out[:, 2] = ((k >> 16) & 0xFFFF) - COORD_BIAS
    out[:, 3] = (k & 0xFFFF) - COORD_BIAS
    return out


def scale_point_coords(c_xyz, init_res: float, after_res: float, scale_mode: str = "mul") -> np.ndarray:
    """fp32 coordinate scaling of models/sparse_utils.py:61-63, with the rounding sequence pinned.

    ``c_xyz`` are the float32 copies of the integer point coordinates (models/spvd.py:436).
    """
    c = np.asarray(c_xyz, dtype=np.float32)
    a = (c * np.float32(init_res)).astype(np.float32)
    if scale_mode == "mul":      # torch-CUDA: a * (1/after_res), reciprocal rounded to fp32
        inv = np.float32(1.0) / np.float32(after_res)
        return (a * inv).astype(np.float32)
    if scale_mode == "div":      # torch-CPU: true division
        return (a / np.float32(after_res)).astype(np.float32)
    raise ValueError(scale_mode)


def unique_voxels(int_coords):
    """torch.unique(int32[N,4], dim=0) + the inverse index (models/sparse_utils.py:66-67).

    Returns (voxel coords int32 [Nv,4] sorted lexicographically, idx int32 [N])."""
    keys = pack_keys(int_coords)
    ukeys, inv = np.unique(keys, return_inverse=True)
    return unpack_keys(ukeys), inv.astype(np.int32).reshape(-1)


def kernel_offsets(kernel_size: int) -> np.ndarray:
    """[UPSTREAM] torchsparse.nn.utils.get_kernel_offsets: odd volume x fastest, even z fastest."""
    r = np.arange(-kernel_size // 2 + 1, kernel_size // 2 + 1)
    if (kernel_size ** 3) % 2 == 1:
        offs = [[x, y, z] for z in r for y in r for x in r]
    else:
        offs = [[x, y, z] for x in r for y in r for z in r]
    return np.asarray(offs, dtype=np.int32)


def hash_query(query, target, kernel_size: int = 1) -> np.ndarray:
    """sphashquery (models/sparse_utils.py:36-55): for each query row and each of kernel_size^3
    offsets the 0-based row of ``target`` holding query+offset (same batch) or -1.

    K=1 -> shape [N] after the caller's reshape (we return [N,1]); K=8 -> offsets {0,1}^3 with z
    fastest, matching the w0..w7 order of calc_ti_weights ([UPSTREAM])."""
    q = np.asarray(query).astype(np.int64)
    tkeys = pack_keys(target)
    order = np.argsort(tkeys, kind="stable")
    skeys = tkeys[order]
    offs = kernel_offsets(kernel_size).astype(np.int64)
    out = np.full((q.shape[0], offs.shape[0]), -1, dtype=np.int32)
    if skeys.size == 0:
        return out
    for k, d in enumerate(offs):
        qq = q.copy()
        qq[:, 1:] += d
        ok = np.abs(qq[:, 1:]).max(axis=1) <= COORD_MAX if qq.size else np.zeros(0, bool)
        qq[~ok, 1:] = 0
        kk = pack_keys(qq)
        pos = np.searchsorted(skeys, kk)
        pos_c = np.minimum(pos, skeys.size - 1)
        hit = ok & (skeys[pos_c] == kk)
        out[hit, k] = order[pos_c[hit]].astype(np.int32)
    return out


def kmap_subm(coords, kernel_size: int = 3) -> np.ndarray:
    """Submanifold kernel map ([UPSTREAM] build_kmap_implicit_GEMM_hashmap_on_the_fly, subm=True):
    out_in_map[o,k] = row of the voxel at C_o + offset_k, or -1.  int32 [Nv, K]."""
    return hash_query(coords, coords, kernel_size)


def downsample_coords(coords, mode: str = "spconv"):
    """Stride-2, kernel-2 output coordinate set ([UPSTREAM] torchsparse spdownsample).

    parent = floor(c/2) per axis (compact coordinates, relied on by models/sparse_utils.py:80-87).
    mode='spconv' (torchsparse 2.1 default): parents with any coordinate above
    ``(max_in - 1)//2`` (batch-wide max per axis; traceback
    experiments/ConditionalGeneration.ipynb:451) are discarded.  mode='floor': keep all.
    Returns (coarse coords sorted, parent row per fine voxel (-1 if dropped), offset index per
    fine voxel = 4*px+2*py+pz)."""
    c = np.asarray(coords).astype(np.int64)
    if c.shape[0] == 0:
        raise RuntimeError("empty voxel level: cannot downsample (reference crashes here too, "
                           "experiments/ConditionalGeneration.ipynb:421)")
    par = c.copy()
    par[:, 1:] = np.floor_divide(c[:, 1:], 2)
    koff = ((c[:, 1] - 2 * par[:, 1]) * 4 + (c[:, 2] - 2 * par[:, 2]) * 2 + (c[:, 3] - 2 * par[:, 3])).astype(np.int32)
    keep = np.ones(c.shape[0], dtype=bool)
    if mode == "spconv":
        cmax = (c[:, 1:].max(axis=0) - 1) // 2
        keep = (par[:, 1:] <= cmax[None, :]).all(axis=1)
    elif mode != "floor":
        raise ValueError(mode)
    keys = pack_keys(par)
    ukeys = np.unique(keys[keep])
    coarse = unpack_keys(ukeys)
    parent = np.full(c.shape[0], -1, dtype=np.int32)
    if ukeys.size:
        pos = np.minimum(np.searchsorted(ukeys, keys), ukeys.size - 1)
        hit = keep & (ukeys[pos] == keys)
        parent[hit] = pos[hit].astype(np.int32)
    return coarse, parent, koff


def kmap_down(n_coarse: int, parent, koff) -> np.ndarray:
    """out_in_map [n_coarse, 8] of the k=2,s=2 conv: column k holds the fine row whose parity
    offset is k, or -1."""
    m = np.full((n_coarse, 8), -1, dtype=np.int32)
    parent = np.asarray(parent)
    rows = np.nonzero(parent >= 0)[0]
    m[parent[rows], np.asarray(koff)[rows]] = rows.astype(np.int32)
    return m


# ----------------------------------------------------------------------------------------------
# floating point work (torch-CPU; differentiable where the reference is)
# ----------------------------------------------------------------------------------------------
def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim_size: int | None = None) -> torch.Tensor:
    """[UPSTREAM] torch_scatter.scatter_mean(src, index, dim=0): sum / max(count,1); the output
    has index.max()+1 rows (models/sparse_utils.py:69,94)."""
    index = index.long().reshape(-1)          # an [N,1] index broadcasts over channels (models/sparse_utils.py:94)
    n = int(index.max()) + 1 if dim_size is None else dim_size
    out = torch.zeros((n, src.shape[1]), dtype=src.dtype).index_add_(0, index, src)
    cnt = torch.zeros(n, dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    return out / cnt.clamp(min=1).unsqueeze(1)


def calc_ti_weights(pf: torch.Tensor, idx: torch.Tensor, scale: float = 1) -> torch.Tensor:
    """[UPSTREAM] torchsparse.nn.functional.devoxelize.calc_ti_weights (call site
    models/sparse_utils.py:114): trilinear weights of the 8 corners (z fastest), zero where the
    corner voxel is absent, renormalised by (sum + 1e-8).  fp32, no autograd."""
    with torch.no_grad():
        p = pf.float()
        f = torch.floor(p / scale) * scale if scale != 1 else torch.floor(p)
        c = f + scale
        x, y, z = p[:, 0:1], p[:, 1:2], p[:, 2:3]
        xf, yf, zf = f[:, 0:1], f[:, 1:2], f[:, 2:3]
        xc, yc, zc = c[:, 0:1], c[:, 1:2], c[:, 2:3]
        w = torch.cat([
            (xc - x) * (yc - y) * (zc - z), (xc - x) * (yc - y) * (z - zf),
            (xc - x) * (y - yf) * (zc - z), (xc - x) * (y - yf) * (z - zf),
            (x - xf) * (yc - y) * (zc - z), (x - xf) * (yc - y) * (z - zf),
            (x - xf) * (y - yf) * (zc - z), (x - xf) * (y - yf) * (z - zf)], dim=1)
        if scale != 1:
            w = w / scale ** 3
        w[idx == -1] = 0
        w = w / (w.sum(dim=1, keepdim=True) + 1e-8)
    return w


def spdevoxelize(feats: torch.Tensor, idx: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """[UPSTREAM] torchsparse.nn.functional.spdevoxelize (call site models/sparse_utils.py:119):
    out[n] = sum_j w[n,j] * feats[idx[n,j]] over present corners."""
    idx = idx.long()
    valid = (idx >= 0)
    g = feats[idx.clamp(min=0)]                                   # [N,8,C]
    return (g * (w * valid).unsqueeze(-1)).sum(dim=1)


def conv_gather_gemm_scatter(feats: torch.Tensor, weight: torch.Tensor, out_in_map, n_out: int,
                             bias: torch.Tensor | None = None, emulate_tf32: bool = False) -> torch.Tensor:
    """Sparse convolution by per-offset gather-GEMM-scatter:
    out[o] = sum_k feats[map[o,k]] @ weight[k] (+ bias)   ([UPSTREAM] spnn.Conv3d forward).
    ``emulate_tf32`` rounds both operands to TF32 (round-to-nearest, ties away) the way the CUDA
    tensor-core path does, for tight comparisons; the default is plain fp32."""
    m = torch.as_tensor(np.asarray(out_in_map)).long()
    if emulate_tf32:
        feats, weight = round_tf32(feats), round_tf32(weight)
    out = torch.zeros((n_out, weight.shape[-1]), dtype=feats.dtype)
    for k in range(m.shape[1]):
        col = m[:, k]
        o = torch.nonzero(col >= 0).reshape(-1)
        if o.numel() == 0:
            continue
        out = out.index_add(0, o, feats[col[o]] @ weight[k])
    if bias is not None:
        out = out + bias
    return out


class _RoundTF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        b = x.detach().contiguous().view(torch.int32)
        r = ((b + 0x1000) & ~0x1FFF)            # cvt.rna.tf32.f32: add half ulp of the 13 dropped bits
        return r.view(torch.float32).view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    return _RoundTF32.apply(x)
