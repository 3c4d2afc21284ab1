"""Host side of the native inference executor (include/spvd_b200.h, section M2).

``compile_model`` walks an ``SPVUnet`` once and emits a flat program of ``spvd_op_t`` records --
the eval-mode forward of the reference (models/spvd.py:432-457) with BatchNorm folded into affines,
FiLM fused with the following BatchNorm+SiLU, residual adds fused into the convolution epilogues
and every nn.Linear expressed as a kernel-size-1 sparse convolution -- over one activation arena
whose slots are assigned by liveness.  ``Program.run`` then executes a whole forward with a single
C call (spvd_program_run); row counts come from the planned Geometry.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import time

import torch
import torch.nn.functional as F

from . import ops
from ._C import check, lib
from .geometry import Geometry

OP_SCATTER_MEAN, OP_DEVOX, OP_CONV, OP_AFFINE, OP_FILM_AFFINE, OP_CAT, OP_LAYERNORM, OP_ATTENTION, OP_TIME_EMBED, OP_CAT_AFFINE, OP_CONV_SP = range(11)
MAP_NONE, MAP_K3, MAP_DOWN, MAP_UP = range(4)
ROWS_POINTS, ROWS_BATCH = -1, -2
FLAG_SILU, FLAG_ROUND, FLAG_A_TF32, FLAG_SIMT, FLAG_F16 = 1, 2, 4, 8, 16
KERNELS_PER_OP = {OP_SCATTER_MEAN: 2, OP_DEVOX: 1, OP_CONV: 1, OP_AFFINE: 1, OP_FILM_AFFINE: 1, OP_CAT: 1,
                  OP_LAYERNORM: 1, OP_ATTENTION: 1, OP_TIME_EMBED: 1, OP_CAT_AFFINE: 1, OP_CONV_SP: 1}


class LevelGeom(C.Structure):
    _fields_ = [("coords", C.c_void_p), ("kmap3", C.c_void_p), ("mask3", C.c_void_p), ("kmap_down", C.c_void_p),
                ("mask_down", C.c_void_p), ("kmap_up", C.c_void_p), ("mask_up", C.c_void_p), ("p2v", C.c_void_p),
                ("devox_idx", C.c_void_p), ("devox_w", C.c_void_p), ("batch_offsets", C.c_void_p),
                ("pair_in", C.c_void_p), ("pair_out", C.c_void_p), ("pair_tile_k", C.c_void_p),
                ("dpair_in", C.c_void_p), ("dpair_out", C.c_void_p), ("dpair_tile_k", C.c_void_p), ("p2v_w", C.c_void_p),
                ("p2v_cnt", C.c_void_p), ("p2v_start", C.c_void_p), ("p2v_order", C.c_void_p), ("p2v_units", C.c_void_p),
                ("kmap3_sorted", C.c_void_p), ("mask3_sorted", C.c_void_p), ("perm3", C.c_void_p),
                ("kmap_down_sorted", C.c_void_p), ("mask_down_sorted", C.c_void_p), ("perm_down", C.c_void_p),
                ("kmap_up_sorted", C.c_void_p), ("mask_up_sorted", C.c_void_p), ("perm_up", C.c_void_p),
                ("n", C.c_int32), ("ld3", C.c_int32), ("ld_down", C.c_int32), ("ld_up", C.c_int32),
                ("max_per_batch", C.c_int32), ("n_pair_tiles", C.c_int32), ("n_dpair_tiles", C.c_int32), ("pad", C.c_int32)]


class Op(C.Structure):
    _fields_ = [("op", C.c_int32), ("flags", C.c_int32), ("level", C.c_int32), ("map_kind", C.c_int32),
                ("rows_in", C.c_int32), ("rows_out", C.c_int32),
                ("c0", C.c_int32), ("c1", C.c_int32), ("kvol", C.c_int32), ("i0", C.c_int32), ("i1", C.c_int32),
                ("pad", C.c_int32),
                ("src0", C.c_int64), ("src1", C.c_int64), ("dst", C.c_int64), ("aux", C.c_int64),
                ("w", C.c_void_p), ("w_packed", C.c_void_p), ("bias", C.c_void_p), ("p0", C.c_void_p), ("p1", C.c_void_p),
                ("src2", C.c_int64), ("src3", C.c_int64), ("w2_packed", C.c_void_p), ("c2", C.c_int32), ("pad2", C.c_int32)]


class _Reg:
    __slots__ = ("rows", "ch", "offset", "last", "name", "half")

    def __init__(self, rows, ch, name=""):
        self.rows, self.ch, self.offset, self.last, self.name, self.half = rows, ch, -1, -1, name, False


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Program:
    def __init__(self, model, n_points, n_batch):
        self.model, self.n_points, self.n_batch = model, n_points, n_batch
        self.dev = next(model.parameters()).device
        self.pair_density = 10.0  # use pair-major convs where a voxel has fewer active offsets than this on average
        self.trace = None         # list -> run() appends (label, perf_counter) marks (tools/step_breakdown.py)
        self.keep = []            # prepared parameter tensors (kept alive; the ops hold raw pointers)
        self.recs = []            # dicts before register allocation
        self.regs = []
        self.strides = list(model._plan["strides"])
        # tensor-core operands of the executor: "f16" (IEEE half, TF32's 11-bit significand, kind::f16) or "tf32"
        self.f16 = getattr(model, "executor_operands", "f16") == "f16"
        # levels whose 3^3 convolutions run on the persistent sorted-tile kernel with fused epilogues (spvd_conv_sp_f16):
        # the fine, many-tile levels; the coarse ones (tens of tiles) keep the split-K kernel
        sp = os.environ.get("SPVD_SP_LEVELS", "1")
        self.sp_levels = {int(v) for v in sp.split(",") if v != ""} if self.f16 else set()
        self._safe_cache, self.tf32_fallbacks = {}, []     # per-layer operand format guard (_f16_safe)
        self._emit()
        self._allocate()
        self._finalize()
        self.key = self.param_key(model)

    # ------------------------------------------------------------------------------ parameters
    @staticmethod
    def param_key(model):
        return tuple((p.data_ptr(), p._version) for p in list(model.parameters()) + list(model.buffers()))

    def _hold(self, t):
        t = t.detach().contiguous().float()
        self.keep.append(t)
        return t

    def _f16_safe(self, w):
        """May this layer's weights be fed to the tensor cores as IEEE half?  Half keeps TF32's 11-bit significand but only 5
        exponent bits: an output channel whose largest weight is below ~2^-12 would be stored in subnormals (or flushed), one
        above ~3e4 would saturate.  Such layers (none with the reference's initialisation; a guard for odd checkpoints)
        fall back to TF32 operands, and their producers keep fp32 activations.  Decided once per compiled program."""
        key = (w.data_ptr(), w._version)
        hit = self._safe_cache.get(key)
        if hit is None:
            w3 = w.detach()
            w3 = w3 if w3.dim() == 3 else (w3.t().unsqueeze(0) if w3.dim() == 2 else w3.reshape(1, 1, -1))
            col = w3.abs().amax(dim=(0, 1))
            hit = bool((col.min() >= 2.0 ** -12) & (col.max() <= 3.0e4))
            self._safe_cache[key] = hit
            if not hit:
                self.tf32_fallbacks.append(tuple(w3.shape))
        return hit

    def _conv_params(self, w3, bias, force_simt=False, safe=None):
        """safe: the f16 verdict of the ORIGINAL parameter (when w3 is a derived view of it); None = judge w3 itself"""
        w3 = self._hold(w3)
        wp = None
        if not force_simt and ops.umma_supported(*w3.shape):
            f16 = self.f16 and (self._f16_safe(w3) if safe is None else safe)
            wp = ops.pack_weights_f16(w3) if f16 else ops.pack_weights(w3)
            self.keep.append(wp)
        return w3, wp, (self._hold(bias) if bias is not None else None)

    def _linear_params(self, lin):
        return self._conv_params(lin.weight.detach().t().unsqueeze(0), lin.bias, safe=self._f16_safe(lin.weight))

    def _bn(self, bn):
        scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
        shift = bn.bias.detach() - bn.running_mean * scale
        return self._hold(scale), self._hold(shift)

    # ------------------------------------------------------------------------------ emission
    def reg(self, rows, ch, name=""):
        r = _Reg(rows, ch, name)
        self.regs.append(r)
        return r

    def op(self, op, dst, src0, src1=None, aux=None, **kw):
        self.recs.append(dict(op=op, dst=dst, src0=src0, src1=src1, aux=aux, **kw))
        return dst

    def _half(self, cin, cout, kvol=1, w=None):
        """should the producer of a [.., cin] tensor write it as f16?  (its only consumer is that tensor-core conv, whose
        weights ``w`` must be representable in half: _f16_safe)"""
        return self.f16 and bool(ops.umma_supported(kvol, cin, cout)) and (w is None or self._f16_safe(w))

    def conv(self, x, params, kvol, map_kind, level, rows_out, cout, residual=None, a_tf32=False):
        w3, wp, bias = params
        if x.half and wp is None:
            raise RuntimeError("an f16 activation reached a convolution without a tensor-core path")
        is16 = wp is not None and wp.dtype == torch.float16
        if x.half and not is16:
            raise RuntimeError("an f16 activation reached a convolution whose weights are not in the f16 image")
        pre = x.half if is16 else a_tf32                # src0 already in the operand format of the tensor-core kernel
        flags = (FLAG_A_TF32 if (pre and wp is not None) else 0) | (FLAG_SIMT if wp is None else 0) | (FLAG_F16 if is16 else 0)
        return self.op(OP_CONV, self.reg(rows_out, cout), x, residual, flags=flags, level=level, map_kind=map_kind,
                       rows_in=x.rows, rows_out=rows_out, c0=x.ch, c1=cout, kvol=kvol, w=w3, w_packed=wp, bias=bias)

    def conv_sp(self, x, params, kvol, map_kind, level, rows_out, cout, residual=None, src2=None, w2=None, film=None,
                aff=None, silu=False, want32=True, want16=False):
        """spvd_conv_sp_f16 record: x (f16) [+ src2 (f16) @ w2] -> fp32 dst (want32) and / or f16 aux = act(v * aff)."""
        w3, wp, bias = params
        assert x.half and wp is not None
        dst = self.reg(rows_out, cout) if want32 else None
        aux = self.reg(rows_out, cout) if want16 else None
        if aux is not None:
            aux.half = True
        kw = dict(flags=(FLAG_SILU if silu else 0) | FLAG_F16, level=level, map_kind=map_kind, rows_in=x.rows, rows_out=rows_out,
                  c0=x.ch, c1=cout, kvol=kvol, w=w3, w_packed=wp, bias=bias, src2=src2, w2_packed=w2,
                  c2=src2.ch if src2 is not None else 0)
        if film is not None:
            kw.update(src3=film[0], i0=film[1], i1=film[2])
        if aff is not None:
            kw.update(p0=aff[0], p1=aff[1])
        self.recs.append(dict(op=OP_CONV_SP, dst=dst, src0=x, src1=residual, aux=aux, **kw))
        return dst, aux

    def linear(self, x, lin, residual=None, a_tf32=False):
        return self.conv(x, self._linear_params(lin), 1, MAP_NONE, -1, x.rows, lin.out_features, residual, a_tf32)

    def _fmt(self, rnd, half):
        return FLAG_F16 if half else FLAG_ROUND if rnd else 0

    def affine(self, x, bn, silu=True, rnd=False, half=False):
        s, b = self._bn(bn)
        y = self.op(OP_AFFINE, self.reg(x.rows, x.ch), x, flags=(FLAG_SILU if silu else 0) | self._fmt(rnd, half),
                    level=-1, rows_in=x.rows, rows_out=x.rows, c0=x.ch, c1=x.ch, p0=s, p1=b)
        y.half = half
        return y

    def _umma_ok(self, lin_or_conv_in, cout):
        return lin_or_conv_in % 32 == 0 and cout % 32 == 0

    def point_block(self, pb, z1, z):
        lin = pb.conv[2]
        rnd = self._umma_ok(lin.in_features, lin.out_features)
        a = self.affine(z, pb.conv[0], True, rnd, rnd and self._half(lin.in_features, lin.out_features, 1, lin.weight))
        return self.linear(a, lin, residual=z1, a_tf32=rnd)

    def res_block(self, blk, x, lvl, pre=None):
        """pre = (a1, x as f16) when the caller fused cat + BN + SiLU (up path)"""
        c1, c2 = blk.conv1[2], blk.conv2[2]
        nf = c1.out_channels
        if pre is not None:
            a1, x = pre
        else:
            a1 = self.affine(x, blk.conv1[0], True, True, self._half(x.ch, nf, 27, c1.kernel))
        half2 = self._half(nf, nf, 27, c2.kernel)
        if lvl in self.sp_levels and a1.half and half2 and not hasattr(blk, "attn") and \
                (not hasattr(blk, "idconv") or x.half):
            # fused block: conv1 -> (FiLM, BN2, SiLU in its epilogue) -> f16 operand of conv2; the 1x1 shortcut runs as extra
            # K slabs of conv2 (its bias joins conv2's), an identity shortcut as the epilogue's residual
            s2, b2 = self._bn(blk.conv2[0])
            _, a2 = self.conv_sp(a1, self._conv_params(c1.kernel, c1.bias), 27, MAP_K3, lvl, lvl, nf,
                                 film=(self.tt, self.tt_width, self.tt_offsets[id(blk)]), aff=(s2, b2), silu=True,
                                 want32=False, want16=True)
            p2 = self._conv_params(c2.kernel, c2.bias)
            if hasattr(blk, "idconv"):
                w2 = ops.pack_weights_f16(self._hold(blk.idconv.weight.detach().t().unsqueeze(0)))      # (x.half: judged safe)
                self.keep.append(w2)
                bias = self._hold(c2.bias.detach() + blk.idconv.bias.detach())
                h2, _ = self.conv_sp(a2, (p2[0], p2[1], bias), 27, MAP_K3, lvl, lvl, nf, src2=x, w2=w2)
            else:
                h2, _ = self.conv_sp(a2, p2, 27, MAP_K3, lvl, lvl, nf, residual=x)
            return h2
        h1 = self.conv(a1, self._conv_params(c1.kernel, c1.bias), 27, MAP_K3, lvl, lvl, nf, a_tf32=True)
        s2, b2 = self._bn(blk.conv2[0])
        off = self.tt_offsets[id(blk)]
        a2 = self.op(OP_FILM_AFFINE, self.reg(lvl, nf), h1, self.tt, flags=FLAG_SILU | self._fmt(True, half2),
                     level=lvl, rows_in=lvl, rows_out=lvl, c0=nf, c1=nf, i0=self.tt_width, i1=off, p0=s2, p1=b2)
        a2.half = half2
        res = self.linear(x, blk.idconv) if hasattr(blk, "idconv") else x
        h2 = self.conv(a2, self._conv_params(c2.kernel, c2.bias), 27, MAP_K3, lvl, lvl, nf, residual=res, a_tf32=True)
        if hasattr(blk, "attn"):
            at = blk.attn
            ln = self.op(OP_LAYERNORM, self.reg(lvl, nf), h2, flags=self._fmt(True, self._half(nf, 3 * nf, 1, at.attn.qkv.weight)), level=-1,
                         rows_in=lvl, rows_out=lvl, c0=nf, c1=nf, p0=self._hold(at.norm1.weight), p1=self._hold(at.norm1.bias),
                         i0=C.c_int32.from_buffer(C.c_float(at.norm1.eps)).value)
            ln.half = self._half(nf, 3 * nf, 1, at.attn.qkv.weight)
            qkv = self.linear(ln, at.attn.qkv, a_tf32=True)
            ao = self.op(OP_ATTENTION, self.reg(lvl, nf), qkv, flags=self._fmt(True, self._half(nf, nf, 1, at.attn.proj.weight)), level=lvl,
                         rows_in=lvl, rows_out=lvl, c0=3 * nf, c1=nf, i0=at.attn.num_heads)
            ao.half = self._half(nf, nf, 1, at.attn.proj.weight)
            h2 = self.linear(ao, at.attn.proj, residual=h2, a_tf32=True)
            self.attn_levels.add(lvl)
        return h2

    def _emit(self):
        m = self.model
        P = ROWS_POINTS
        pc = m.point_channels
        n_emb = m.emb_mlp[1][1].out_features
        # one GEMM for all FiLM projections: tt = Linear_cat(silu(emb))  [B, sum 2C]
        blocks = [mod for mod in m.modules() if mod.__class__.__name__ == "EmbResBlock"]
        ws, bs, self.tt_offsets, col = [], [], {}, 0
        for b in blocks:
            self.tt_offsets[id(b)] = col
            ws.append(b.t_emb.proj_mlp.weight.detach())
            bs.append(b.t_emb.proj_mlp.bias.detach())
            col += b.t_emb.proj_mlp.out_features
        self.tt_width = col
        self.attn_levels = set()
        self.in_feats = self.reg(P, pc, "feats")
        # silu(emb_mlp(sinusoid(t)) [+ cond_emb(cls)]) in one launch; t / cls pointers are patched in by run()
        bn, l1, l2 = m.emb_mlp[0][0], m.emb_mlp[0][2], m.emb_mlp[1][1]
        freqs = (-math.log(10000) * torch.linspace(0, 1, m.n_temb // 2, device=self.dev)).exp()
        te = self._hold(torch.cat([freqs, *self._bn(bn), l1.weight.detach().reshape(-1), l1.bias.detach(),
                                   l2.weight.detach().reshape(-1), l2.bias.detach()]))
        self.in_emb = self.op(OP_TIME_EMBED, self.reg(ROWS_BATCH, n_emb, "emb_act"), None, flags=FLAG_SILU, level=-1,
                              rows_in=ROWS_BATCH, rows_out=ROWS_BATCH, c0=m.n_temb, c1=n_emb, w=te,
                              p0=self._hold(m.cond_emb.weight) if hasattr(m, "cond_emb") else None)
        wcat = torch.cat(ws, 0).t().unsqueeze(0)          # [1, n_emb, sum 2C]
        self.tt = self.conv(self.in_emb, self._conv_params(wcat, torch.cat(bs, 0), force_simt=True), 1, MAP_NONE, -1,
                            ROWS_BATCH, col)

        def scatter(z, lvl):
            return self.op(OP_SCATTER_MEAN, self.reg(lvl, z.ch), z, aux=self.reg(lvl, 1, "counts"), flags=0, level=lvl,
                           rows_in=P, rows_out=lvl, c0=z.ch, c1=z.ch)

        def devox(x, lvl):
            return self.op(OP_DEVOX, self.reg(P, x.ch), x, flags=0, level=lvl, rows_in=lvl, rows_out=P, c0=x.ch, c1=x.ch)

        z = self.in_feats
        lvl = 0
        x = scatter(z, 0)
        st = m.stem_conv
        c0, c3 = st.voxel_conv[0], st.voxel_conv[3]
        x = self.conv(x, self._conv_params(c0.kernel, c0.bias), 27, MAP_K3, 0, 0, c0.out_channels)
        x = self.affine(x, st.voxel_conv[1], True, True, self._half(c3.in_channels, c3.out_channels, 27, c3.kernel))
        x = self.conv(x, self._conv_params(c3.kernel, c3.bias), 27, MAP_K3, 0, 0, c3.out_channels, a_tf32=True)
        z = self.point_block(st.point_conv, devox(x, 0), z)
        x = scatter(z, 0)
        saved = [x]
        for stage in m.down_stages:
            for blk in stage.downs:
                for r in blk.resnets:
                    x = self.res_block(r, x, lvl)
                    saved.append(x)
                d = blk.down
                if blk.add_down:
                    x = self.conv(x, self._conv_params(d.kernel, d.bias), 8, MAP_DOWN, lvl, lvl + 1, d.out_channels)
                    lvl += 1
                    saved.append(x)
                else:
                    x = self.conv(x, self._conv_params(d.kernel.unsqueeze(0), d.bias), 1, MAP_NONE, -1, lvl, d.out_channels)
            z = self.point_block(stage.point_block, devox(x, lvl), z)
            x = scatter(z, lvl)
        for si, stage in enumerate(m.up_stages):
            for blk in stage.ups:
                for r in blk.resnets:
                    skip = saved.pop()
                    cc, nf = x.ch + skip.ch, r.conv1[2].out_channels
                    if hasattr(r, "idconv") and self._half(cc, nf, 27, r.conv1[2].kernel) and self._half(cc, nf, 1, r.idconv.weight):
                        # cat + BN + SiLU in one pass, both results as f16: the block reads the concatenation only through
                        # its two tensor-core convolutions (first 3^3 conv, 1x1 shortcut)
                        s1, b1 = self._bn(r.conv1[0])
                        xc16 = self.reg(lvl, cc, "cat_f16")
                        a1 = self.op(OP_CAT_AFFINE, self.reg(lvl, cc), x, skip, aux=xc16, flags=FLAG_SILU | FLAG_F16, level=-1,
                                     rows_in=lvl, rows_out=lvl, c0=x.ch, c1=skip.ch, p0=s1, p1=b1)
                        a1.half = xc16.half = True
                        x = self.res_block(r, None, lvl, pre=(a1, xc16))
                    else:
                        xc = self.op(OP_CAT, self.reg(lvl, cc), x, skip, flags=0, level=-1, rows_in=lvl,
                                     rows_out=lvl, c0=x.ch, c1=skip.ch)
                        x = self.res_block(r, xc, lvl)
                u = blk.up
                if blk.add_up:
                    lvl -= 1
                    x = self.conv(x, self._conv_params(u.kernel, u.bias), 8, MAP_UP, lvl, lvl, u.out_channels)
                    self.recs[-1]["rows_in"] = lvl + 1
                else:
                    x = self.conv(x, self._conv_params(u.kernel.unsqueeze(0), u.bias), 1, MAP_NONE, -1, lvl, u.out_channels)
            z = self.point_block(stage.point_block, devox(x, lvl), z)
            if si + 1 < len(m.up_stages):
                x = scatter(z, lvl)
        assert not saved
        a = self.affine(z, m.out_conv[0], True, False)
        self.out = self.linear(a, m.out_conv[2])

    # ------------------------------------------------------------------------------ arena
    def _bytes(self, r):
        rows = self.n_batch if r.rows == ROWS_BATCH else self.n_points
        return (rows * r.ch * 4 + 255) // 256 * 256

    def _allocate(self):
        for i, rec in enumerate(self.recs):
            for k in ("src0", "src1", "src2", "src3", "aux", "dst"):
                if rec.get(k) is not None:
                    rec[k].last = i
        self.in_feats.last = max(self.in_feats.last, 0)
        self.out.last = len(self.recs) + 1
        free, top = [], 0                       # free: list of (offset, size)

        def alloc(r):
            nonlocal top
            need = self._bytes(r)
            for j, (off, size) in enumerate(free):
                if size >= need:
                    if size == need:
                        free.pop(j)
                    else:
                        free[j] = (off + need, size - need)
                    r.offset = off
                    return
            r.offset = top
            top += need

        def release(r):
            free.append((r.offset, self._bytes(r)))
            free.sort()
            merged = []
            for off, size in free:
                if merged and merged[-1][0] + merged[-1][1] == off:
                    merged[-1] = (merged[-1][0], merged[-1][1] + size)
                else:
                    merged.append((off, size))
            free[:] = merged

        alloc(self.in_feats)
        alloc(self.in_emb)
        for i, rec in enumerate(self.recs):
            for k in ("dst", "aux"):
                r = rec[k]
                if r is not None and r.offset < 0:
                    alloc(r)
            for r in {id(rec[k]): rec[k] for k in ("src0", "src1", "src2", "src3", "aux", "dst") if rec.get(k) is not None}.values():
                if r.last == i and r is not self.out:
                    release(r)
        self.arena_bytes = top + 256

    def _finalize(self):
        n = len(self.recs)
        self.c_ops = (Op * n)()
        self.n_kernels = 0
        self.conv_meta = []
        for i, rec in enumerate(self.recs):
            o = self.c_ops[i]
            o.op, o.flags, o.level, o.map_kind = rec["op"], rec.get("flags", 0), rec.get("level", -1), rec.get("map_kind", 0)
            o.rows_in, o.rows_out = rec["rows_in"], rec["rows_out"]
            o.c0, o.c1, o.kvol, o.i0, o.i1 = rec["c0"], rec["c1"], rec.get("kvol", 0), rec.get("i0", 0), rec.get("i1", 0)
            o.src0 = rec["src0"].offset if rec["src0"] is not None else -1
            o.src1 = rec["src1"].offset if rec["src1"] is not None else -1
            o.dst = rec["dst"].offset if rec["dst"] is not None else -1
            o.aux = rec["aux"].offset if rec["aux"] is not None else -1
            o.src2 = rec["src2"].offset if rec.get("src2") is not None else -1
            o.src3 = rec["src3"].offset if rec.get("src3") is not None else -1
            o.c2 = rec.get("c2", 0)
            for k in ("w", "w_packed", "bias", "p0", "p1", "w2_packed"):
                t = rec.get(k)
                setattr(o, k, t.data_ptr() if t is not None else None)
            self.n_kernels += KERNELS_PER_OP[rec["op"]]
            if rec["op"] in (OP_CONV, OP_CONV_SP):
                # (kvol, cin [+ cin2 of the folded shortcut], cout, rows, level, map kind, kernel family)
                self.conv_meta.append((rec["kvol"], rec["c0"], rec["c1"], rec["rows_out"], rec["level"], rec["map_kind"],
                                       "sp" if rec["op"] == OP_CONV_SP else "umma" if (rec.get("w_packed") is not None) else "simt",
                                       rec.get("c2", 0)))
        # finest level an op can run with: it needs every level up to max(level, rows_in, rows_out) planned; the stride-2
        # maps / pair lists of level l are planned together with level l + 1
        self.op_max_level, self.convs_before, nconv = [], [], 0
        for rec in self.recs:
            top = max(rec.get("level", -1), rec["rows_in"], rec["rows_out"])
            self.op_max_level.append(top)
            self.convs_before.append(nconv)
            nconv += rec["op"] in (OP_CONV, OP_CONV_SP)
        self.arena = torch.empty(self.arena_bytes, dtype=torch.uint8, device=self.dev)
        self.arena_f = self.arena.view(torch.float32)
        # leading ops that read nothing but the timestep / class ids: time embedding and the FiLM projection GEMM
        self.n_prologue = 2 if (n >= 2 and self.recs[0]["op"] == OP_TIME_EMBED and self.recs[1]["op"] == OP_CONV and
                                self.recs[1]["src0"] is self.in_emb and self.recs[1]["dst"] is self.tt) else 0
        self.prologue_done = torch.cuda.Event()
        self._prologue_launched, self._time_keep = False, None
        self.c_levels = (LevelGeom * len(self.strides))()

    # ------------------------------------------------------------------------------ execution
    def _view(self, r, rows):
        o = r.offset // 4
        return self.arena_f[o: o + rows * r.ch].view(rows, r.ch)

    def fill_levels(self, geo: Geometry, upto=None, start=0):
        """(Re)write the level records [start, upto); level start - 1 gets only what is planned together with its coarser
        neighbour (stride-2 maps and pair lists), which is all that can have changed for it."""
        n_levels = len(self.strides) if upto is None else upto
        for i in range(max(start - 1, 0), n_levels):
            s = self.strides[i]
            lv, g = geo.level(s), self.c_levels[i]
            if lv.kmap_down is not None:
                g.kmap_down, g.mask_down, g.ld_down = (lv.kmap_down.map_t.data_ptr(), lv.kmap_down.mask.data_ptr(),
                                                       lv.kmap_down.map_t.shape[1])
                g.kmap_up, g.mask_up, g.ld_up = lv.kmap_up.map_t.data_ptr(), lv.kmap_up.mask.data_ptr(), lv.kmap_up.map_t.shape[1]
            dp = getattr(lv, "dpairs", None)          # planned together with the next coarser level
            if dp is not None:
                g.dpair_in, g.dpair_out, g.dpair_tile_k, g.n_dpair_tiles = dp[0].data_ptr(), dp[1].data_ptr(), dp[2].data_ptr(), dp[3]
            else:
                g.dpair_in, g.n_dpair_tiles = None, 0
            if i < start:
                continue
            g.coords, g.n = lv.coords_cap.data_ptr(), lv.n
            if lv.kmap3 is not None:
                g.kmap3, g.mask3, g.ld3 = lv.kmap3.map_t.data_ptr(), lv.kmap3.mask.data_ptr(), lv.kmap3.map_t.shape[1]
            srt = getattr(lv, "kmap3_sorted", None)
            g.kmap3_sorted, g.mask3_sorted, g.perm3 = [t.data_ptr() for t in srt] if srt is not None else [None] * 3
            g.p2v = geo.p2v[s].data_ptr() if s in geo.p2v else None
            pw = getattr(geo, "p2v_w", None)
            g.p2v_w = pw[s].data_ptr() if pw and s in pw else None
            grp = getattr(geo, "p2v_groups", {}).get(s)
            g.p2v_cnt, g.p2v_start, g.p2v_order, g.p2v_units = [t.data_ptr() for t in grp] if grp is not None else [None] * 4
            if s in geo.devox:
                g.devox_idx, g.devox_w = geo.devox[s][0].data_ptr(), geo.devox[s][1].data_ptr()
            pr = getattr(lv, "pairs", None)
            if pr is not None and pr[4] < self.pair_density * lv.n:      # sparse level: run the 3^3 convs pair-major
                g.pair_in, g.pair_out, g.pair_tile_k, g.n_pair_tiles = pr[0].data_ptr(), pr[1].data_ptr(), pr[2].data_ptr(), pr[3]
            else:
                g.pair_in, g.n_pair_tiles = None, 0
            if i in self.attn_levels:
                g.batch_offsets = geo.batch_offsets(s).data_ptr()
                g.max_per_batch = max(lv.batch_counts_host)

    def _bind_time(self, t, cls):
        if t.shape[0] != self.n_batch:
            raise RuntimeError("Program was compiled for a different number of clouds")
        t = t.long().contiguous()
        cls = cls.long().contiguous() if cls is not None else None
        te = self.c_ops[0]                      # OP_TIME_EMBED reads this call's timesteps / class ids
        assert te.op == OP_TIME_EMBED
        te.p1 = t.data_ptr()
        te.bias = cls.data_ptr() if cls is not None else None
        self._time_keep = (t, cls)              # the op holds raw pointers

    def prologue(self, t, cls, side_stream, after=None):
        """The ops that depend on the timestep only (embedding MLP, all FiLM projections) on a side stream, to run beside
        the geometry planning instead of in front of the stem; ``run`` then starts behind them.  ``after``: an event on the
        current stream to wait for instead of everything enqueued on it so far (the planner may already be in the queue)."""
        if self.n_prologue == 0:
            return
        self._bind_time(t, cls)
        cur = torch.cuda.current_stream()
        if after is not None:
            side_stream.wait_event(after)
        else:
            side_stream.wait_stream(cur)        # the previous forward still reads the FiLM table
        check(lib.spvd_program_run(self.c_ops, self.n_prologue, self.c_levels, len(self.strides), self.n_points, self.n_batch,
                                   C.c_void_p(self.arena.data_ptr()), self.arena_bytes, None, C.c_void_p(side_stream.cuda_stream)))
        self.prologue_done.record(side_stream)
        self._prologue_launched = True

    def stage_input(self, feats: torch.Tensor):
        """Copy the point features into the arena now (stream-ordered behind the previous forward): called before the
        planner's host synchronisation so that nothing but the op launches stands between the host waking up and the
        first feature kernel."""
        if feats.shape[0] != self.n_points:
            raise RuntimeError("Program was compiled for a different number of points / clouds")
        self._view(self.in_feats, self.n_points).copy_(feats)
        self._input_staged = feats

    def run(self, geo: Geometry, feats: torch.Tensor, t: torch.Tensor, conv_events=None, cls=None):
        m = self.model
        if feats.shape[0] != self.n_points or t.shape[0] != self.n_batch:
            raise RuntimeError("Program was compiled for a different number of points / clouds")
        first = 0
        if self._prologue_launched and conv_events is None:
            torch.cuda.current_stream().wait_event(self.prologue_done)
            first = self.n_prologue
        else:
            self._bind_time(t, cls)
        self._prologue_launched = False
        if getattr(self, "_input_staged", None) is not feats:
            self._view(self.in_feats, self.n_points).copy_(feats)
        self._input_staged = None
        ev = None
        if conv_events is not None:
            ev = (C.c_void_p * len(conv_events))(*conv_events)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def launch(lo, hi):
            if hi <= lo:
                return
            evp = C.byref(ev, 2 * self.convs_before[lo] * C.sizeof(C.c_void_p)) if ev is not None else None
            check(lib.spvd_program_run(C.byref(self.c_ops, lo * C.sizeof(Op)), hi - lo, self.c_levels, len(self.strides),
                                       self.n_points, self.n_batch, C.c_void_p(self.arena.data_ptr()), self.arena_bytes, evp,
                                       stream))

        n, k, filled = len(self.recs), first, 0
        while geo._pending is not None:
            # the finest levels are planned: run the ops that touch only those while the coarser ones are still being
            # planned on the side stream, then collect the next group of levels and go on
            ready = geo.levels_ready
            k2 = next((i for i, l in enumerate(self.op_max_level) if l >= ready), n)
            self.fill_levels(geo, upto=ready, start=filled)
            filled = ready
            launch(k, k2)
            k = k2
            if self.trace is not None:
                self.trace.append((f"ops on {ready} levels launched", time.perf_counter()))
            geo.finish(upto=ready + 1)
            if self.trace is not None:
                self.trace.append((f"{geo.levels_ready} levels collected", time.perf_counter()))
        self.fill_levels(geo, start=filled)
        launch(k, n)
        if self.trace is not None:
            self.trace.append(("part2 launched", time.perf_counter()))
        ops.stats.launches += self.n_kernels
        return self._view(self.out, self.n_points).clone()
