"""Differentiable front of the CUDA ops (torch.autograd.Function wrappers; coordinates, indices and
trilinear weights carry no gradient, as in the reference -- SURVEY.md section 3.3)."""
from __future__ import annotations

import torch

from . import ops


class PackedWeight:
    """Tensor-core tile image of a conv weight, rebuilt when the parameter changes (optimizer step)."""

    def __init__(self):
        self.key, self.packed = None, None

    def get(self, w3: torch.Tensor):
        key = (w3.data_ptr(), w3._version, tuple(w3.shape))
        if key != self.key:
            self.packed = ops.pack_weights(w3.detach()) if ops.umma_supported(*w3.shape) else None
            self.key = key
        return self.packed


class _SparseConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, kmap, cache, impl, a_tf32=False):
        w3 = weight if weight.dim() == 3 else weight.unsqueeze(0)
        w3 = w3.contiguous()
        wp = cache.get(w3) if (cache is not None and impl != ops.CONV_SIMT) else None
        if impl == ops.CONV_UMMA and wp is None:
            wp = ops.pack_weights(w3)
        n_out = kmap.n_out if kmap is not None else x.shape[0]
        x = x.contiguous()
        out = ops.conv_fwd(x, w3, wp, kmap.map_t if kmap is not None else None,
                           kmap.mask if kmap is not None else None, n_out, bias, impl, a_tf32=a_tf32)
        ctx.save_for_backward(x, weight)
        ctx.kmap, ctx.impl, ctx.has_bias = kmap, impl, bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        kmap, impl = ctx.kmap, ctx.impl
        g = g.contiguous()
        w3 = (weight if weight.dim() == 3 else weight.unsqueeze(0)).contiguous()
        kvol, cin, cout = w3.shape
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            flip = kmap.flip if kmap is not None else False
            wt = ops.transpose_weights(w3, flip)                       # [kvol, cout, cin]
            use_umma = impl != ops.CONV_SIMT and ops.umma_supported(kvol, cout, cin)
            wtp = ops.pack_weights(wt) if use_umma else None
            gm = kmap.grad if kmap is not None else None
            n_in = kmap.n_in if kmap is not None else g.shape[0]
            gx = ops.conv_fwd(g, wt, wtp, gm.map_t if gm is not None else None, gm.mask if gm is not None else None,
                              n_in, None, ops.CONV_AUTO if use_umma else ops.CONV_SIMT)
        if ctx.needs_input_grad[1]:
            n_out = kmap.n_out if kmap is not None else x.shape[0]
            if impl != ops.CONV_SIMT and ops.umma_supported(kvol, cin, cout):
                # tensor cores: reduction over the pair lists, TF32 operands (rounded here), fp32 accumulation
                gw = ops.conv_wgrad_umma(ops.round_tf32(x), ops.round_tf32(g), kmap.pairs() if kmap is not None else None,
                                         n_out, kvol).view_as(weight)
            else:
                gw = ops.conv_wgrad(x, g, kmap.map_t if kmap is not None else None, n_out, kvol).view_as(weight)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g.sum(0)
        return gx, gw, gb, None, None, None, None


def sparse_conv(x, weight, bias, kmap, cache=None, impl=ops.CONV_AUTO, a_tf32=False):
    """out[o] = sum_k x[kmap[k][o]] @ weight[k] (+ bias); kmap None = kernel size 1.  ``a_tf32``: the caller
    guarantees x is already TF32-representable (enables the cp.async gather of the tcgen05 path)."""
    return _SparseConv.apply(x, weight, bias, kmap, cache, impl, a_tf32)


class _ScatterMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, idx, nv):
        src = src.contiguous()
        out, counts = ops.scatter_mean_fwd(src, idx, nv)
        ctx.save_for_backward(idx, counts)
        ctx.n = src.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        idx, counts = ctx.saved_tensors
        return ops.scatter_mean_bwd(g.contiguous(), idx, counts, ctx.n), None, None


def scatter_mean(src, idx, nv):
    """torch_scatter.scatter_mean(src, idx, dim=0) with nv output rows (models/sparse_utils.py:69,94)."""
    return _ScatterMean.apply(src, idx, nv)


class _Devoxelize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, idx8, w8):
        feats = feats.contiguous()
        ctx.save_for_backward(idx8, w8)
        ctx.nv = feats.shape[0]
        return ops.devoxelize_fwd(feats, idx8, w8)

    @staticmethod
    def backward(ctx, g):
        idx8, w8 = ctx.saved_tensors
        return ops.devoxelize_bwd(g.contiguous(), idx8, w8, ctx.nv), None, None


def devoxelize(feats, idx8, w8):
    """torchsparse.nn.functional.spdevoxelize (models/sparse_utils.py:119,128)."""
    return _Devoxelize.apply(feats, idx8, w8)
