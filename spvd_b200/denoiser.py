"""The SPVD point-voxel denoiser on the sm_100a backend.

Same module tree, constructor arguments and parameter names as the reference's ``SPVUnet``
(models/spvd.py:367-457; building blocks models/spvd.py:16-271, models/ddpm_unet_attn.py:24-65,
models/modules/graph.py:8-41, models/modules/transformer.py:40-125), so a reference ``state_dict``
loads with ``strict=True`` -- but the forward is organised differently: the coordinate work of the
whole U-Net is planned once (spvd_b200.geometry.Geometry, one host sync instead of ~20) and the
network then runs on plain feature matrices through the CUDA ops of libspvd_b200.so.  The modules
below only own parameters; ``SPVUnet.forward`` drives them.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as SF
from . import ops
from .geometry import Geometry, PlanWorkspace
from .sparse import SparseTensor


# ------------------------------------------------------------------------------ parameter holders
class Conv3d(nn.Module):
    """Parameters of torchsparse.nn.Conv3d: ``kernel`` [K,Cin,Cout] ([Cin,Cout] for K=1), optional
    ``bias``; uniform(+-1/sqrt(K*fan)) init with fan = Cout when transposed else Cin."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, bias=False,
                 transposed=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.transposed = int(kernel_size), int(stride), bool(transposed)
        kvol = self.kernel_size ** 3
        shape = (kvol, in_channels, out_channels) if kvol > 1 else (in_channels, out_channels)
        bound = 1.0 / math.sqrt((out_channels if transposed else in_channels) * kvol)
        self.kernel = nn.Parameter(torch.empty(shape).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.empty(out_channels).uniform_(-bound, bound)) if bias else None
        self._packed = SF.PackedWeight()
        self.impl = ops.CONV_AUTO

    def forward(self, feats, kmap, a_tf32=False, residual=None):
        return SF.sparse_conv(feats, self.kernel, self.bias, kmap, self._packed, self.impl, a_tf32, residual)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"transposed={self.transposed}, bias={self.bias is not None}")


def _bn_silu_linear(ni, nf, bias=True):
    return nn.Sequential(nn.BatchNorm1d(ni), nn.SiLU(), nn.Linear(ni, nf, bias=bias))


class PointBlock(nn.Module):                      # models/spvd.py:16-28
    def __init__(self, ni, nf):
        super().__init__()
        self.conv = _bn_silu_linear(ni, nf)


class StemBlock(nn.Module):                       # models/spvd.py:30-62
    def __init__(self, ni, nf, ks=3):
        super().__init__()
        self.voxel_conv = nn.Sequential(Conv3d(ni, nf, ks), nn.BatchNorm1d(nf), nn.SiLU(), Conv3d(nf, nf, ks))
        self.point_conv = PointBlock(ni, nf)


class TimeEmbeddingBlock(nn.Module):              # models/modules/graph.py:8-41
    def __init__(self, n_emb, embed_dim):
        super().__init__()
        self.proj_mlp = nn.Linear(n_emb, embed_dim * 2)


class _AttentionCore(nn.Module):                  # models/modules/transformer.py:40-67
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.proj = nn.Linear(dim, dim)


class SparseAttention(nn.Module):                 # models/modules/transformer.py:102-125
    def __init__(self, dim, num_heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _AttentionCore(dim, num_heads)


def _unet_conv(ni, nf, ks):                       # models/ddpm_unet_attn.py:24-30
    return nn.Sequential(nn.BatchNorm1d(ni), nn.SiLU(), Conv3d(ni, nf, ks, bias=True))


class EmbResBlock(nn.Module):                     # models/ddpm_unet_attn.py:32-65
    def __init__(self, n_emb, ni, nf=None, ks=3, attn_chans=None):
        super().__init__()
        nf = nf or ni
        self.conv1 = _unet_conv(ni, nf, ks)
        self.conv2 = _unet_conv(nf, nf, ks)
        self.t_emb = TimeEmbeddingBlock(n_emb, nf)
        if ni != nf:
            self.idconv = nn.Linear(ni, nf)
        if attn_chans:
            self.attn = SparseAttention(nf, attn_chans)


class DownBlock(nn.Module):                       # models/spvd.py:102-119
    def __init__(self, n_emb, ni, nf, add_down=True, num_layers=1, attn_chans=None):
        super().__init__()
        self.resnets = nn.ModuleList([EmbResBlock(n_emb, ni, ni, attn_chans=attn_chans) for _ in range(num_layers)])
        self.add_down = add_down
        self.down = Conv3d(ni, nf, 2, stride=2) if add_down else Conv3d(ni, nf, 1)


class SPVDownStage(nn.Module):                    # models/spvd.py:121-159
    def __init__(self, n_emb, nfs, add_down, num_layers, attn_chans):
        super().__init__()
        self.downs = nn.ModuleList([DownBlock(n_emb, nfs[i], nfs[i + 1], add_down[i], num_layers, attn_chans[i])
                                    for i in range(len(nfs) - 1)])
        self.point_block = PointBlock(nfs[0], nfs[-1])


class UpBlock(nn.Module):                         # models/spvd.py:212-233
    def __init__(self, n_emb, prev_nf, ni, add_up=True, num_layers=2, attn_chans=None, ks=3):
        super().__init__()
        self.resnets = nn.ModuleList([EmbResBlock(n_emb, prev_nf + ni, prev_nf, ks=ks, attn_chans=attn_chans)
                                      for _ in range(num_layers)])
        self.add_up = add_up
        self.up = Conv3d(prev_nf, ni, 2, stride=2, transposed=True) if add_up else Conv3d(prev_nf, ni, 1)


class SPVUpStage(nn.Module):                      # models/spvd.py:235-271
    def __init__(self, n_emb, nfs, add_up, num_layers, attn_chans, ks):
        super().__init__()
        self.ups = nn.ModuleList()
        prev = nfs[0]
        for i in range(len(nfs) - 1):
            self.ups.append(UpBlock(n_emb, prev, nfs[i + 1], add_up[i], num_layers, attn_chans[i], ks[i]))
            prev = nfs[i + 1]
        self.point_block = PointBlock(nfs[0], nfs[-1])


# ------------------------------------------------------------------------------------- the network
class SPVUnet(nn.Module):
    """``model((SparseTensor x, LongTensor t)) -> FloatTensor [N_points, point_channels]`` exactly like the
    reference (models/spvd.py:432-457); ``x.C`` int32 [N,4] = (batch, ix, iy, iz) in units of ``pres``,
    ``x.F`` the point features.

    Extra knobs that the reference hard-wires inside torchsparse: ``downsample_mode`` ('spconv' =
    torchsparse 2.1 default boundary rule, 'floor'), ``scale_mode`` ('mul' = what torch-CUDA computes
    for ``c*pres/voxel_size``, 'div' = torch-CPU), ``conv_impl`` (ops.CONV_AUTO / CONV_SIMT / CONV_UMMA).
    """

    def __init__(self, point_channels=3, voxel_size=0.1,
                 down_blocks=[[(64, 128, 192, 256, 384, 384), (True, True, True, True, False), (None, None, None, 8, 8)]],
                 up_blocks=[[(384, 384, 256), (True, True), (8, 8), (3, 3)],
                            [(256, 192, 128, 64), (True, True, False), (None, None, None), (3, 3, 3)]],
                 num_layers=1, pres=1e-5, downsample_mode="spconv", scale_mode="mul", n_classes=None):
        super().__init__()
        self.pres, self.voxel_size = pres, voxel_size
        self.downsample_mode, self.scale_mode = downsample_mode, scale_mode
        self.point_channels = point_channels
        self.n_temb = nf0 = down_blocks[0][0][0]
        n_emb = nf0 * 4
        self.emb_mlp = nn.Sequential(_bn_silu_linear(self.n_temb, n_emb), nn.Sequential(nn.SiLU(), nn.Linear(n_emb, n_emb)))
        if n_classes:        # CondSPVUnet (experiments/ConditionalGeneration.ipynb cell 4): model((x, t, c))
            self.cond_emb = nn.Embedding(n_classes, n_emb)
        self.stem_conv = StemBlock(point_channels, nf0)
        self.down_stages = nn.ModuleList([SPVDownStage(n_emb, nfs, add_down, num_layers, attn)
                                          for nfs, add_down, attn in down_blocks])
        self.mid_stages = nn.Identity()
        self.up_stages = nn.ModuleList([SPVUpStage(n_emb, nfs, add_up, num_layers + 1, attn, ks)
                                        for nfs, add_up, attn, ks in up_blocks])
        c_out = up_blocks[-1][0][-1]
        self.out_conv = _bn_silu_linear(c_out, point_channels, bias=False)
        self._plan = self._make_plan(down_blocks, up_blocks)
        self._folded = {}
        self._lin_cache = {}         # id(nn.Linear) -> SF.LinearCache (transposed weight + tile image of the op-by-op path)
        self._impl = ops.CONV_AUTO
        self._programs = {}          # (n_points, n_batch) -> runtime.Program (native inference executor)
        self.use_executor = True
        self._plan_ws = {}           # (n_points, n_batch, device) -> [PlanWorkspace pool, cursor]
        self.geometry_pool = 2
        # level groups of the executor's planning: it waits for the first group only, the others are planned under the
        # ops of the finer levels ("1,2" = level 0 | level 1 | the rest; "0" = plan everything up front)
        self.plan_split_level = tuple(int(v) for v in os.environ.get("SPVD_PLAN_SPLIT", "1,3").split(","))
        self._side_streams = {}
        self._prefetch_streams = {}
        self._pre_plan_events = {}
        self._aux_streams = {}
        self.side_prologue = os.environ.get("SPVD_SIDE_PROLOGUE", "0") == "1"
        # tensor-core operand format of the executor: "f16" = IEEE half operands (the 11-bit significand of TF32, fp32
        # accumulation, twice the issue rate and half the operand bytes) or "tf32"; training always uses TF32
        self.executor_operands = os.environ.get("SPVD_EXEC_OPERANDS", "f16")
        # strides whose 3^3 maps also get a mask-sorted view (spvd_kmap_sort) for the executor's sorted-tile convolutions
        self.sort_strides = tuple(1 << int(v) for v in os.environ.get("SPVD_SORT_LEVELS", "").split(",") if v != "")

    # -- static description of which strides need which maps (so Geometry can build all of them up front)
    @staticmethod
    def _make_plan(down_blocks, up_blocks):
        s, strides, p2v, devox = 1, [1], [1], [1]
        for nfs, add_down, _ in down_blocks:
            for d in add_down:
                if d:
                    s *= 2
                    strides.append(s)
            p2v.append(s)
            devox.append(s)
        top = s
        for nfs, add_up, _, _ in up_blocks:
            for u in add_up:
                if u:
                    s //= 2
            p2v.append(s)
            devox.append(s)
        assert s >= 1
        return dict(strides=tuple(strides), p2v=tuple(sorted(set(p2v))), devox=tuple(sorted(set(devox))), top=top)

    # -- compiled inference programs hold raw pointers to prepared weights: drop them whenever the
    #    parameters can have changed (mode switch, device move, state-dict load)
    def invalidate_programs(self):
        self._programs.clear()
        self._folded.clear()
        self.__dict__.pop("_stamp_tensors", None)

    def train(self, mode=True):
        self.invalidate_programs()
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_programs()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_programs()
        return super().load_state_dict(*args, **kwargs)

    def set_conv_impl(self, impl):
        self.invalidate_programs()
        self.use_executor = impl == ops.CONV_AUTO
        self._impl = impl
        for m in self.modules():
            if isinstance(m, Conv3d):
                m.impl = impl
        return self

    def _param_stamp(self):
        ps = self.__dict__.get("_stamp_tensors")
        if ps is None:
            ps = self.__dict__["_stamp_tensors"] = list(self.parameters()) + list(self.buffers())
        return sum(p._version for p in ps), sum(p.data_ptr() for p in ps[:8])

    def num_params(self):
        return sum(p.numel() for p in self.parameters())

    # -- helpers -------------------------------------------------------------------------------
    def _fast(self):
        return (not self.training) and (not torch.is_grad_enabled())

    def _bn_act(self, bn: nn.BatchNorm1d, x, act=True, tf32=False):
        """BatchNorm1d (+ SiLU).  Inference: one fused kernel on the folded affine; ``tf32`` also rounds the
        result to TF32 there, so that the convolution consuming it can gather with cp.async.  Training: batch statistics,
        forward and backward on the library's own kernels (csrc/train.cu)."""
        if self._fast() or not bn.training:
            key = id(bn)
            ver = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version)
            hit = self._folded.get(key)
            if hit is None or hit[0] != ver:
                scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
                shift = bn.bias.detach() - bn.running_mean * scale
                hit = (ver, scale.contiguous(), shift.contiguous())
                self._folded[key] = hit
            if self._fast():
                return ops.affine_act(x.contiguous(), hit[1], hit[2], (1 if act else 0) | (2 if tf32 else 0))
            # eval mode with autograd (frozen statistics, trainable affine): plain torch, gradients reach weight / bias
            scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
            y = x * scale + (bn.bias - bn.running_mean * scale)
            return F.silu(y) if act else y
        return SF.batch_norm_act(x, bn, act, round_tf32=tf32)

    def _linear(self, lin: nn.Linear, x, residual=None, a_tf32=False):
        """nn.Linear as a kernel-size-1 convolution of the library (tcgen05 TF32 / fp32 SIMT), forward and backward"""
        cache = self._lin_cache.get(id(lin))
        if cache is None:
            cache = self._lin_cache[id(lin)] = SF.LinearCache()
        return SF.linear(x, lin.weight, lin.bias, cache, self._impl, residual, a_tf32)

    def _tc(self):
        """do the convolutions of this call run on the tensor cores (then their BatchNorm producers round to TF32)"""
        return self._impl != ops.CONV_SIMT

    def _point_block(self, pb: PointBlock, z1, z):
        lin = pb.conv[2]
        tc = self._tc() and self.training and bool(ops.umma_supported(1, lin.in_features, lin.out_features))
        return self._linear(lin, self._bn_act(pb.conv[0], z, tf32=tc), residual=z1, a_tf32=tc)

    def _attention(self, blk: SparseAttention, x, geo: Geometry, s):
        """models/modules/transformer.py:118-125: x + proj(attention(qkv(LayerNorm(x)))) per cloud -- rows of a level are
        batch-major, so the padded [B, Nmax] layout of the reference is replaced by per-cloud row ranges"""
        lv = geo.level(s)
        y = SF.layer_norm(x, blk.norm1)
        qkv = self._linear(blk.attn.qkv, y)
        o = SF.attention(qkv, geo.batch_offsets(s), len(lv.batch_counts_host), max(lv.batch_counts_host), blk.attn.num_heads)
        return self._linear(blk.attn.proj, o, residual=x)

    def _res_block(self, blk: EmbResBlock, x_in, emb_act, geo: Geometry, s):
        lv = geo.level(s)
        fast = self._fast()
        # BatchNorm rounds its output to TF32 for the tensor-core conv that consumes it (inference kernels and the
        # training-mode kernels do; the plain-torch eval-with-autograd branch of _bn_act does not)
        tc = fast or (self._tc() and self.training)
        x = blk.conv1[2](self._bn_act(blk.conv1[0], x_in, tf32=tc), lv.kmap3, tc)
        tt = self._linear(blk.t_emb.proj_mlp, emb_act)                      # [B, 2C]
        if fast:
            x = ops.film(x, tt.contiguous(), lv.coords)
        else:
            x = SF.film(x, tt, lv.coords, geo.batch_offsets(s), max(lv.batch_counts_host))
        res = self._linear(blk.idconv, x_in) if hasattr(blk, "idconv") else x_in
        x = blk.conv2[2](self._bn_act(blk.conv2[0], x, tf32=tc), lv.kmap3, tc, residual=res)
        if hasattr(blk, "attn"):
            x = self._attention(blk.attn, x, geo, s)
        return x

    # -- forward -------------------------------------------------------------------------------
    def plan_geometry(self, coords: torch.Tensor, n_batch: int, max_cloud_rows: int = 0, split_level=0, defer=False) -> Geometry:
        """Plan every coordinate structure of the forward with one C call into persistent buffers.  Two
        buffer sets alternate, so the geometry of the previous call stays valid while the next one is
        planned (``geometry_pool`` > 2 keeps more alive, e.g. when several forwards precede one backward)."""
        p = self._plan
        pc = (coords.float() if coords.dtype != torch.float32 else coords).contiguous()
        key = (pc.shape[0], n_batch, pc.device, self.training)
        pool = self._plan_ws.get(key)
        if pool is None:
            pool = self._plan_ws[key] = [[], 0]
        if len(pool[0]) < self.geometry_pool:
            # training: pair lists at every level (the weight gradient reduces over them); inference: only where the
            # executor can choose pair-major convolutions
            pool[0].append(PlanWorkspace(pc.shape[0], n_batch, p["strides"], p["strides"], p["p2v"], p["devox"], pc.device,
                                         pair_strides=None if self.training else (1, 2),
                                         sort_strides=() if self.training else self.sort_strides))
            ws = pool[0][-1]
            pool[1] = len(pool[0]) - 1           # the cursor follows the set just used: the next call takes the OTHER one
        else:
            pool[1] = (pool[1] + 1) % len(pool[0])
            ws = pool[0][pool[1]]
        side = None
        if split_level and split_level != (0,):
            side = self._side_streams.get(pc.device)
            if side is None:
                side = self._side_streams[pc.device] = torch.cuda.Stream(device=pc.device, priority=-1)
        return ws.build(pc, self.pres, self.voxel_size, self.scale_mode, self.downsample_mode, max_cloud_rows, split_level, side,
                        defer=defer)

    def prefetch_geometry(self, x_in, n_batch: int) -> Geometry:
        """Plan the geometry of a FUTURE batch now, on a side stream, without waiting for it: pass the result as
        ``forward(inp, geometry=...)`` (or ``train_step(..., geometry=...)``) when that batch's turn comes.  The planner's
        one host synchronisation (level sizes) otherwise sits at the start of every forward and drains the GPU: the host
        cannot issue the next step while the device still runs the previous backward.  Call it for batch j + 1 BEFORE issuing
        step j (what a data loader's prefetch would do): the side stream first waits for everything issued so far -- the
        last user of the buffer set it is about to overwrite (two sets alternate) -- and then runs beside step j."""
        coords = x_in.C if hasattr(x_in, "C") else x_in
        cf = getattr(x_in, "coords_float", None)
        if cf is not None and cf.shape[0] == coords.shape[0]:
            coords = cf
        dev = coords.device
        ps = self._prefetch_streams.get(dev)
        if ps is None:
            ps = self._prefetch_streams[dev] = torch.cuda.Stream(device=dev)
        pc = (coords.float() if coords.dtype != torch.float32 else coords).contiguous()
        ps.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(ps):
            geo = self.plan_geometry(pc, n_batch, int(getattr(x_in, "points_per_cloud", 0) or 0), 0, defer=True)
        pc.record_stream(ps)
        return geo

    def forward(self, inp, geometry: Geometry = None):
        x_in, t = inp[0], inp[1]
        coords, feats = (x_in.C, x_in.F) if isinstance(x_in, SparseTensor) or hasattr(x_in, "C") else x_in
        if not feats.is_cuda:
            raise RuntimeError("spvd_b200.SPVUnet runs on CUDA only (no CPU fallback)")
        cf = getattr(x_in, "coords_float", None)          # float copy of the coordinates, when the sampler made one
        if cf is not None and cf.shape[0] == coords.shape[0]:
            coords = cf
        cls = inp[2] if len(inp) > 2 else None
        if cls is not None and not hasattr(self, "cond_emb"):
            raise RuntimeError("this model was built without n_classes; it takes (x, t)")
        executor = self.use_executor and self._fast()
        prog, geo = None, geometry
        if executor:
            from .runtime import Program
            split = self.plan_split_level
            early = geometry is None and getattr(self, "profile_sink", None) is None
            before_plan = None
            if early:
                # the planner goes into the queue FIRST (it needs the coordinates only) and is collected after the host-side
                # work below: program lookup / staleness stamp, the timestep-only ops on a side stream, staging the features.
                # A caller that synchronises every step (host buffers in, host buffers out) no longer pays that work serially.
                before_plan = self._pre_plan_events.get(feats.device)
                if before_plan is None:
                    before_plan = self._pre_plan_events[feats.device] = torch.cuda.Event()
                before_plan.record()
                geo = self.plan_geometry(coords, t.shape[0], int(getattr(x_in, "points_per_cloud", 0) or 0), split, defer=True)
            key = (feats.shape[0], t.shape[0], self.executor_operands)
            prog = self._programs.get(key)
            # a compiled program holds derived copies of the parameters (packed tile images, folded BatchNorm): rebuild it
            # when any parameter or buffer was modified in place since (optimizer / EMA step, manual edits in eval mode)
            stamp = self._param_stamp()
            if prog is None or prog.stamp != stamp:
                prog = self._programs[key] = Program(self, *key[:2])
                prog.stamp = stamp
            if early and split not in (0, (0,)) and self.side_prologue:
                # timestep-only ops run on a side stream beside the geometry planning
                aux = self._aux_streams.get(feats.device)
                if aux is None:
                    aux = self._aux_streams[feats.device] = torch.cuda.Stream(device=feats.device, priority=-1)
                prog.prologue(t, cls, aux, after=before_plan)
            if geometry is None:
                feats = feats.contiguous().float()
                prog.stage_input(feats)          # before the planner's host wait, not after it
            if early:
                geo.finish(upto=1)               # the one host wait: the finest level group's sizes
            elif geometry is None:
                geo = self.plan_geometry(coords, t.shape[0], int(getattr(x_in, "points_per_cloud", 0) or 0), split)
        elif geometry is None:
            geo = self.plan_geometry(coords, t.shape[0], int(getattr(x_in, "points_per_cloud", 0) or 0), 0)
        if geometry is not None and geo._pending:
            geo.finish()                        # prefetched (prefetch_geometry): its planner finished long ago
        if executor:
            sink = getattr(self, "profile_sink", None)
            if sink is None:
                return prog.run(geo, feats.contiguous().float(), t, cls=cls)
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * len(prog.conv_meta))]
            for e in evs:
                e.record()                      # materialises the cudaEvent_t handles
            out = prog.run(geo, feats.contiguous().float(), t, conv_events=[e.cuda_event for e in evs], cls=cls)
            pairs = {}
            for (kvol, cin, cout, rows_out, level, map_kind, kind, cin2) in prog.conv_meta:
                if map_kind and (level, map_kind) not in pairs:
                    lv = geo.level(prog.strides[level])
                    km = {1: lv.kmap3, 2: lv.kmap_down, 3: lv.kmap_up}[map_kind]
                    pairs[(level, map_kind)] = (km.map_t[:, :km.n_out] >= 0).sum()
            sizes = {i: geo.level(s).n for i, s in enumerate(prog.strides)}
            sink.append((evs, prog.conv_meta, {k: int(v.item()) for k, v in pairs.items()}, sizes))
            return out
        # timestep embedding (models/utils.py:15-19; linspace(0,1,d/2) includes 1.0) + MLP
        e = ops.timestep_embedding(t.long().contiguous(), self.n_temb)
        emb = self._linear(self.emb_mlp[0][2], self._bn_act(self.emb_mlp[0][0], e))
        emb = self._linear(self.emb_mlp[1][1], SF.silu(emb))
        if cls is not None:
            emb = SF.embedding_add(emb, self.cond_emb.weight, cls.long().contiguous())
        emb_act = SF.silu(emb)                     # every TimeEmbeddingBlock starts with SiLU(emb)

        def v2p(x, s):
            idx8, w8 = geo.devox[s]
            return SF.devoxelize(x, idx8, w8, geo)

        def p2v(zf, s):
            return SF.scatter_mean(zf, geo.p2v[s], geo.level(s).n, geo)

        z = feats.contiguous().float()
        s = 1
        x = p2v(z, 1)                                                        # initial_voxelize
        st = self.stem_conv
        x = st.voxel_conv[0](x, geo.level(1).kmap3)
        tc = self._fast() or (self._tc() and self.training)
        x = self._bn_act(st.voxel_conv[1], x, tf32=tc)
        x = st.voxel_conv[3](x, geo.level(1).kmap3, tc)
        z = self._point_block(st.point_conv, v2p(x, 1), z)
        x = p2v(z, 1)
        saved = [x]
        for stage in self.down_stages:
            for blk in stage.downs:
                for r in blk.resnets:
                    x = self._res_block(r, x, emb_act, geo, s)
                    saved.append(x)
                if blk.add_down:
                    x = blk.down(x, geo.level(s).kmap_down)
                    s *= 2
                    saved.append(x)
                else:
                    x = blk.down(x, None)
            z = self._point_block(stage.point_block, v2p(x, s), z)
            x = p2v(z, s)
        for si, stage in enumerate(self.up_stages):
            for blk in stage.ups:
                for r in blk.resnets:
                    x = self._res_block(r, SF.cat(x, saved.pop()), emb_act, geo, s)
                if blk.add_up:
                    s //= 2
                    x = blk.up(x, geo.level(s).kmap_up)
                else:
                    x = blk.up(x, None)
            z = self._point_block(stage.point_block, v2p(x, s), z)
            if si + 1 < len(self.up_stages):
                x = p2v(z, s)            # the reference's last point_to_voxel is dead work (models/spvd.py:269)
        return self._linear(self.out_conv[2], self._bn_act(self.out_conv[0], z))


# models/__init__.py:9-38
PRESETS = {
    "SPVD": dict(point_channels=3, voxel_size=0.1, num_layers=1, pres=1e-5,
                 down_blocks=[[(32, 64, 128, 192, 192, 256), (True, True, True, True, False), (None, None, None, 8, 8)]],
                 up_blocks=[[(256, 192, 192), (True, True), (8, 8), (3, 3)],
                            [(192, 128, 64, 32), (True, True, False), (None, None, None), (3, 3, 3)]]),
    "SPVD_L": dict(point_channels=3, voxel_size=0.1, num_layers=1, pres=1e-5,
                   down_blocks=[[(64, 128, 192, 256, 384, 384), (True, True, True, True, False), (None, None, None, 8, 8)]],
                   up_blocks=[[(384, 384, 256), (True, True), (8, 8), (3, 3)],
                              [(256, 192, 128, 64), (True, True, False), (None, None, None), (3, 3, 3)]]),
    "SPVD_TINY": dict(point_channels=3, voxel_size=0.1, num_layers=1, pres=1e-5,
                      down_blocks=[[(32, 32, 64, 64), (True, True, False), (None, 4, 4)]],
                      up_blocks=[[(64, 64), (True,), (4,), (3,)],
                                 [(64, 32, 32), (True, False), (None, None), (3, 3)]]),
    # experiments/ConditionalGeneration.ipynb cell 4 (CondSPVUnet) and experiments/SuperResolution.ipynb cell 12
    "COND_SPVD_L": dict(point_channels=3, voxel_size=0.1, num_layers=1, pres=1e-5, n_classes=55,
                        down_blocks=[[(64, 128, 192, 256, 384, 384), (True, True, True, True, False), (None, None, None, 8, 8)]],
                        up_blocks=[[(384, 384, 256), (True, True), (8, 8), (3, 3)],
                                   [(256, 192, 128, 64), (True, True, False), (None, None, None), (3, 3, 3)]]),
    "SUPERRES": dict(point_channels=4, voxel_size=0.1, num_layers=1, pres=1e-5,
                     down_blocks=[[(64, 128, 192, 256, 256), (True, True, True, False), (None, None, None, None)]],
                     up_blocks=[[(256, 256, 192), (True, True), (None, None), (3, 3)],
                                [(192, 128, 64), (True, False), (None, None), (3, 3)]]),
    "COND_TINY": dict(point_channels=3, voxel_size=0.1, num_layers=1, pres=1e-5, n_classes=7,
                      down_blocks=[[(32, 32, 64, 64), (True, True, False), (None, 4, 4)]],
                      up_blocks=[[(64, 64), (True,), (4,), (3,)],
                                 [(64, 32, 32), (True, False), (None, None), (3, 3)]]),
}


class CondSPVUnet(SPVUnet):
    """``model((x, t, c))`` with a class embedding added to the time embedding (ConditionalGeneration.ipynb cell 4)."""

    def __init__(self, point_channels=3, voxel_size=0.1, n_classes=55, **kw):
        super().__init__(point_channels=point_channels, voxel_size=voxel_size, n_classes=n_classes, **kw)


def build_model(preset="SPVD", **overrides) -> SPVUnet:
    cfg = dict(PRESETS[preset])
    cfg.update(overrides)
    return SPVUnet(**cfg)
