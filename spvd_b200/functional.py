"""Differentiable front of the CUDA ops (torch.autograd.Function wrappers; coordinates, indices and
trilinear weights carry no gradient, as in the reference -- SURVEY.md section 3.3)."""
from __future__ import annotations

import torch

from . import ops


# ------------------------------------------------------------------------------------------------------------------
# In-place parameter gradients.  With a parallel.GradientAllReduce attached, every parameter's .grad is a slot of one flat
# buffer that is zeroed once per step.  A backward below then lets its kernel write the parameter gradient straight into
# that slot and returns None to autograd: no temporary, no AccumulateGrad add (one elementwise launch per parameter and
# step, 276 for SPVD), and the reducer learns that the slot is final.  Only for a parameter used exactly once since the
# last zero_grad() (a second use, or a second backward without zero_grad, goes through autograd's accumulation as usual).
# ------------------------------------------------------------------------------------------------------------------
def _use(p):
    sink = getattr(p, "_spvd_sink", None)
    if sink is not None:
        sink.note_use(p)
    return p


def _claim(p, wanted=True):
    sink = getattr(p, "_spvd_sink", None) if (wanted and p is not None) else None
    return sink.claim(p) if sink is not None else None


def _wrote(*ps):
    for p in ps:
        p._spvd_sink.wrote(p)


class PackedWeight:
    """Tensor-core tile image of a conv weight, rebuilt when the parameter changes (optimizer step).  With ``flip_t`` given
    (a training forward whose input needs a gradient) the transposed image for the data gradient comes out of the same
    launch (``packed_t``, built with ``t_flip``)."""

    def __init__(self):
        self.key, self.packed, self.packed_t, self.t_flip = None, None, None, 0

    def get(self, w3: torch.Tensor, flip_t=None):
        key = (w3.data_ptr(), w3._version, tuple(w3.shape))
        if key != self.key or (flip_t is not None and (self.packed_t is None or self.t_flip != flip_t)):
            self.packed_t, self.t_flip = None, 0
            if flip_t is None:
                self.packed = ops.pack_weights(w3.detach()) if ops.umma_supported(*w3.shape) else None
            else:
                _, self.packed, self.packed_t = ops.pack_weights_train(w3.detach(), *w3.shape, flip_t=flip_t)
                self.t_flip = flip_t
            self.key = key
        return self.packed


def _dgrad_flip(kmap):
    """the offset order the data gradient of a conv over ``kmap`` will want for its transposed weights (spvd_conv_bwd):
    pair-major on sparse levels (no flip), dense with flipped offsets on submanifold maps"""
    if kmap is None:
        return 0
    if kmap.n_pairs and kmap.n_pairs < 10 * kmap.n_in:
        return 0
    return 1 if kmap.flip else 0


class _SparseConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, kmap, cache, impl, a_tf32=False, residual=None):
        w3 = weight if weight.dim() == 3 else weight.unsqueeze(0)
        w3 = w3.contiguous()
        flip_t = _dgrad_flip(kmap) if (ctx.needs_input_grad[0] and weight.requires_grad) else None
        wp = cache.get(w3, flip_t) if (cache is not None and impl != ops.CONV_SIMT) else None
        if impl == ops.CONV_UMMA and wp is None:
            wp = ops.pack_weights(w3)
        ctx.cache = cache if flip_t is not None else None
        n_out = kmap.n_out if kmap is not None else x.shape[0]
        x = x.contiguous()
        res = residual.contiguous() if residual is not None else None
        if (wp is not None and kmap is not None and kmap.n_pairs and kmap.n_pairs < 10 * n_out and kmap._pairs is not None
                and (kmap.kvol % 2 == 1 or (bias is None and res is None))):
            # sparse level (few active offsets per voxel): pair-major tensor-core kernel, as the inference executor chooses
            pin, pout, tk, nt = kmap._pairs
            out = ops.conv_fwd_pairs(x, w3, wp, pin, pout, tk, nt, n_out, bias, res, a_tf32=a_tf32)
        else:
            out = ops.conv_fwd(x, w3, wp, kmap.map_t if kmap is not None else None,
                               kmap.mask if kmap is not None else None, n_out, bias, impl, a_tf32=a_tf32, residual=res)
        ctx.save_for_backward(x, weight)
        ctx.kmap, ctx.impl, ctx.has_bias, ctx.a_tf32 = kmap, impl, bias is not None, bool(a_tf32)
        ctx.params = (_use(weight), _use(bias) if bias is not None else None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        kmap, impl = ctx.kmap, ctx.impl
        p_w, p_b = ctx.params
        want_gb = ctx.has_bias and ctx.needs_input_grad[2]
        gw_out, gb_out = _claim(p_w, ctx.needs_input_grad[1]), _claim(p_b, want_gb)
        if kmap is not None:
            kmap.check_current()
        w3 = (weight if weight.dim() == 3 else weight.unsqueeze(0)).contiguous()
        n_out = kmap.n_out if kmap is not None else x.shape[0]
        n_in = kmap.n_in if kmap is not None else g.shape[0]
        # one C call: TF32-rounded gradient -> data gradient (cp.async gather) + weight gradient (tcgen05 over the pair
        # lists; x was rounded by its producer when the forward ran with a_tf32) + bias gradient
        wt, wt_flip = None, 0
        c = ctx.cache
        if c is not None and c.packed_t is not None and c.key == (w3.data_ptr(), w3._version, tuple(w3.shape)):
            wt, wt_flip = c.packed_t, c.t_flip                     # (the weights have not changed since the forward)
        gx, gw, gb = ops.conv_bwd(x, g.contiguous(), w3, kmap, n_in, n_out, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                  want_gb, x_is_tf32=ctx.a_tf32, tensor_cores=impl != ops.CONV_SIMT, gw_out=gw_out, gb_out=gb_out,
                                  w_packed_t=wt, w_packed_t_flip=wt_flip)
        if gw_out is not None:
            _wrote(p_w)
            gw = None
        elif gw is not None:
            gw = gw.view_as(weight)
        if gb_out is not None:
            _wrote(p_b)
            gb = None
        return gx, gw, gb, None, None, None, None, (g if ctx.needs_input_grad[7] else None)


def sparse_conv(x, weight, bias, kmap, cache=None, impl=ops.CONV_AUTO, a_tf32=False, residual=None):
    """out[o] = sum_k x[kmap[k][o]] @ weight[k] (+ bias) (+ residual[o]); kmap None = kernel size 1.  ``a_tf32``: the
    caller guarantees x is already TF32-representable (enables the cp.async gather of the tcgen05 path)."""
    return _SparseConv.apply(x, weight, bias, kmap, cache, impl, a_tf32, residual)


class _ScatterMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, idx, nv, guard=None):
        src = src.contiguous()
        out, counts = ops.scatter_mean_fwd(src, idx, nv)
        ctx.save_for_backward(idx, counts)
        ctx.n, ctx.guard = src.shape[0], guard
        return out

    @staticmethod
    def backward(ctx, g):
        if ctx.guard is not None:
            ctx.guard.check_current()            # idx is a view of a geometry workspace: refuse one that was re-planned
        idx, counts = ctx.saved_tensors
        return ops.scatter_mean_bwd(g.contiguous(), idx, counts, ctx.n), None, None, None


def scatter_mean(src, idx, nv, guard=None):
    """torch_scatter.scatter_mean(src, idx, dim=0) with nv output rows (models/sparse_utils.py:69,94).  guard: the
    Geometry that owns idx (its check_current() runs before the backward uses the indices)."""
    return _ScatterMean.apply(src, idx, nv, guard)


class _Devoxelize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, idx8, w8, guard=None):
        feats = feats.contiguous()
        ctx.save_for_backward(idx8, w8)
        ctx.nv, ctx.guard = feats.shape[0], guard
        return ops.devoxelize_fwd(feats, idx8, w8)

    @staticmethod
    def backward(ctx, g):
        if ctx.guard is not None:
            ctx.guard.check_current()
        idx8, w8 = ctx.saved_tensors
        return ops.devoxelize_bwd(g.contiguous(), idx8, w8, ctx.nv), None, None, None


def devoxelize(feats, idx8, w8, guard=None):
    """torchsparse.nn.functional.spdevoxelize (models/sparse_utils.py:119,128); guard as for scatter_mean."""
    return _Devoxelize.apply(feats, idx8, w8, guard)


# ------------------------------------------------------------------------------------------------------------------
# Training-mode glue on the library's own kernels (csrc/train.cu): with these, a training step of the denoiser
# launches no ATen / cuBLAS / cuDNN kernel between the sparse convolutions.
# ------------------------------------------------------------------------------------------------------------------
class LinearCache:
    """[1, in, out] transpose of an nn.Linear weight + its tensor-core tile image, rebuilt when the parameter changes"""

    def __init__(self):
        self.key, self.w3, self.packed, self.packed_t = None, None, None, None

    def get(self, weight, want_packed, want_t=False):
        """want_t: also the transposed tile image for the data gradient (same launch)"""
        key = (weight.data_ptr(), weight._version)
        cout, cin = weight.shape
        if key != self.key or (want_t and self.packed_t is None and ops.umma_supported(1, cout, cin)):
            if want_packed or want_t:            # conv layout + both tile images from one pass over the weight
                self.w3, self.packed, self.packed_t = ops.pack_weights_train(weight.detach().contiguous(), 1, cin, cout, linear=True,
                                                                             want_w3=True, want_t=want_t)
            else:
                self.w3 = ops.transpose_weights(weight.detach().unsqueeze(0).contiguous(), False)  # [1, out, in] -> [1, in, out]
                self.packed, self.packed_t = None, None
            self.key = key
        if want_packed and self.packed is None and ops.umma_supported(*self.w3.shape):
            self.packed = ops.pack_weights(self.w3)
        return self.w3, (self.packed if want_packed else None)


class _Linear(torch.autograd.Function):
    """nn.Linear (weight [out, in]) as a kernel-size-1 convolution: forward / data gradient on the conv kernels (tcgen05 TF32
    where the widths allow, fp32 SIMT otherwise), weight gradient = transposed conv weight gradient, bias gradient = column
    sum; an optional residual rides on the epilogue."""

    @staticmethod
    def forward(ctx, x, weight, bias, cache, impl, residual, a_tf32=False):
        x = x.contiguous()
        umma = impl != ops.CONV_SIMT and ops.umma_supported(1, weight.shape[1], weight.shape[0])
        want_t = impl != ops.CONV_SIMT and ctx.needs_input_grad[0] and weight.requires_grad
        w3, wp = cache.get(weight, umma, want_t)
        out = ops.conv_fwd(x, w3, wp, None, None, x.shape[0], bias, ops.CONV_AUTO if umma else ops.CONV_SIMT,
                           a_tf32=bool(a_tf32) and umma, residual=residual.contiguous() if residual is not None else None)
        ctx.save_for_backward(x, weight, w3)
        ctx.cache = cache if want_t else None
        ctx.impl, ctx.has_bias, ctx.a_tf32 = impl, bias is not None, bool(a_tf32)
        ctx.params = (_use(weight), _use(bias) if bias is not None else None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight, w3 = ctx.saved_tensors                                       # w3 = [1, in, out] (the conv layout)
        p_w, p_b = ctx.params
        want_gb = ctx.has_bias and ctx.needs_input_grad[2]
        gw_out, gb_out = _claim(p_w, ctx.needs_input_grad[1]), _claim(p_b, want_gb)
        c = ctx.cache
        wt = c.packed_t if (c is not None and c.key == (weight.data_ptr(), weight._version)) else None
        gx, gw, gb = ops.conv_bwd(x, g.contiguous(), w3, None, x.shape[0], x.shape[0], ctx.needs_input_grad[0],
                                  ctx.needs_input_grad[1], want_gb, x_is_tf32=ctx.a_tf32,
                                  gw_transposed=True, tensor_cores=ctx.impl != ops.CONV_SIMT, gw_out=gw_out, gb_out=gb_out,
                                  w_packed_t=wt, w_packed_t_flip=0)
        if gw_out is not None:
            _wrote(p_w)
            gw = None
        elif gw is not None:
            gw = gw[0]                                                           # [out, in], nn.Linear's layout
        if gb_out is not None:
            _wrote(p_b)
            gb = None
        return gx, gw, gb, None, None, (g if ctx.needs_input_grad[5] else None), None


def linear(x, weight, bias, cache, impl=ops.CONV_AUTO, residual=None, a_tf32=False):
    """a_tf32: x is already TF32-representable (its producer rounded it)"""
    return _Linear.apply(x, weight, bias, cache, impl, residual, a_tf32)


class _BatchNormAct(torch.autograd.Function):
    """nn.BatchNorm1d with batch statistics (+ SiLU); updates the running statistics like torch."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, act):
        x = x.contiguous()
        y, stats = ops.bn_train_fwd(x, gamma, beta, eps, momentum, running_mean, running_var, act)
        ctx.save_for_backward(x, gamma, beta, stats)
        ctx.act = act
        ctx.params = (_use(gamma), _use(beta))
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gamma, beta, stats = ctx.saved_tensors
        p_g, p_b = ctx.params
        gg_out = gb_out = None
        if ctx.needs_input_grad[1] and ctx.needs_input_grad[2]:
            gg_out = _claim(p_g)
            gb_out = _claim(p_b) if gg_out is not None else None
        gx, gg, gb = ops.bn_train_bwd(x, gy.contiguous(), gamma, beta, stats, ctx.act, gg_out, gb_out)
        if gg_out is not None:
            _wrote(p_g)
            gg = None
        if gb_out is not None:
            _wrote(p_b)
            gb = None
        return gx, gg, gb, None, None, None, None, None


def batch_norm_act(x, bn, act=True, round_tf32=False):
    """round_tf32: also round the result to TF32 (it is the operand of a tensor-core convolution, which would round it in
    its gather anyway: the forward value is the same, the gather becomes a plain cp.async and the weight gradient can use
    the saved tensor as it is)"""
    return _BatchNormAct.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps,
                               bn.momentum if bn.momentum is not None else 0.1, (1 if act else 0) | (2 if round_tf32 else 0))


class _FiLM(torch.autograd.Function):
    """(1 + scale[b]) * x + shift[b], [scale | shift] = tt[b] (models/modules/graph.py:35-41); rows of a cloud contiguous"""

    @staticmethod
    def forward(ctx, x, tt, coords, offsets, max_rows):
        x, tt = x.contiguous(), tt.contiguous()
        ctx.save_for_backward(x, tt, offsets)
        ctx.max_rows = max_rows
        return ops.film(x, tt, coords)

    @staticmethod
    def backward(ctx, gy):
        x, tt, offsets = ctx.saved_tensors
        gx, gtt = ops.film_bwd(x, gy.contiguous(), tt, offsets, ctx.max_rows)
        return gx, gtt, None, None, None


def film(x, tt, coords, offsets, max_rows):
    return _FiLM.apply(x, tt, coords, offsets, max_rows)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = x.contiguous()
        y, st = ops.layernorm_train_fwd(x, gamma, beta, eps)
        ctx.save_for_backward(x, gamma, st)
        ctx.params = (_use(gamma), _use(beta))
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gamma, st = ctx.saved_tensors
        p_g, p_b = ctx.params
        gg_out = gb_out = None
        if ctx.needs_input_grad[1] and ctx.needs_input_grad[2]:
            gg_out = _claim(p_g)
            gb_out = _claim(p_b) if gg_out is not None else None
        gx, gg, gb = ops.layernorm_bwd(x, gy.contiguous(), gamma, st, gg_out, gb_out)
        if gg_out is not None:
            _wrote(p_g)
            gg = None
        if gb_out is not None:
            _wrote(p_b)
            gb = None
        return gx, gg, gb, None


def layer_norm(x, ln):
    return _LayerNorm.apply(x, ln.weight, ln.bias, ln.eps)


class _Attention(torch.autograd.Function):
    """per-cloud multi-head softmax attention over qkv [rows, 3C] (models/modules/transformer.py:54-67)"""

    @staticmethod
    def forward(ctx, qkv, offsets, n_batch, max_len, heads):
        qkv = qkv.contiguous()
        out, lse = ops.attention_train_fwd(qkv, offsets, n_batch, max_len, heads)
        ctx.save_for_backward(qkv, out, lse, offsets)
        ctx.meta = (n_batch, max_len, heads)
        return out

    @staticmethod
    def backward(ctx, gout):
        qkv, out, lse, offsets = ctx.saved_tensors
        return ops.attention_bwd(qkv, out, gout.contiguous(), lse, offsets, *ctx.meta), None, None, None, None


def attention(qkv, offsets, n_batch, max_len, heads):
    return _Attention.apply(qkv, offsets, n_batch, max_len, heads)


class _Cat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.ca = a.shape[1]
        return ops.cat_any(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, g):
        return ops.split(g.contiguous(), ctx.ca)


def cat(a, b):
    return _Cat.apply(a, b)


class _SiLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.silu_fwd(x)

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        return ops.silu_bwd(x, gy.contiguous())


def silu(x):
    return _SiLU.apply(x)


class _EmbeddingAdd(torch.autograd.Function):
    """base + table[idx] (the class embedding added to the time embedding)"""

    @staticmethod
    def forward(ctx, base, table, idx):
        ctx.save_for_backward(idx)
        ctx.n = table.shape[0]
        ctx.params = (_use(table),)
        return ops.embedding_fwd(table, idx, base.contiguous())

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = g.contiguous()
        out = _claim(ctx.params[0], ctx.needs_input_grad[1])
        gt = ops.embedding_bwd(g, idx, ctx.n, out)
        if out is not None:
            _wrote(ctx.params[0])
            gt = None
        return g, gt, None


def embedding_add(base, table, idx):
    return _EmbeddingAdd.apply(base, table, idx)


class _MSELoss(torch.autograd.Function):
    """nn.MSELoss()(pred, target): loss and its gradient from one pass"""

    @staticmethod
    def forward(ctx, pred, target):
        loss, gp = ops.mse_loss(pred.contiguous(), target.contiguous(), want_grad=True)
        ctx.save_for_backward(gp)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (gp,) = ctx.saved_tensors
        return gp * g, None


def mse_loss(pred, target):
    return _MSELoss.apply(pred, target)
