"""Functional CPU restatement of the SPVD denoiser (reference: models/spvd.py:367-457 ``SPVUnet``).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  This file exists because /root/reference is not on
the GPU box: the parity tests, ``smoke()`` and ``bench.py --impl reference`` need a checker that
travels.  It is written as a walk over a *state dict* with the reference's parameter names (so it
also pins checkpoint compatibility); there is no nn.Module here.  It was validated in the build
container against the reference's own Python running on oracle/shim.py (oracle/make_golden.py,
tests/test_oracle_model.py); the outputs of that run are the fixtures in tests/golden/.

All sparse arithmetic goes through oracle/ops.py; dense glue is plain torch-CPU, so autograd gives
the backward oracle.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import ops

# ------------------------------------------------------------------------------------ presets
# models/__init__.py:9-38
PRESETS = {
    "SPVD": dict(point_channels=3, voxel_size=0.1, pres=1e-5, num_layers=1,
                 down_blocks=[[(32, 64, 128, 192, 192, 256), (True, True, True, True, False), (None, None, None, 8, 8)]],
                 up_blocks=[[(256, 192, 192), (True, True), (8, 8), (3, 3)],
                            [(192, 128, 64, 32), (True, True, False), (None, None, None), (3, 3, 3)]]),
    "SPVD_L": dict(point_channels=3, voxel_size=0.1, pres=1e-5, num_layers=1,
                   down_blocks=[[(64, 128, 192, 256, 384, 384), (True, True, True, True, False), (None, None, None, 8, 8)]],
                   up_blocks=[[(384, 384, 256), (True, True), (8, 8), (3, 3)],
                              [(256, 192, 128, 64), (True, True, False), (None, None, None), (3, 3, 3)]]),
    # experiments/ConditionalGeneration.ipynb cell 4: SPVD-L topology + nn.Embedding(55, n_emb) added to the time embedding
    "COND_SPVD_L": dict(point_channels=3, voxel_size=0.1, pres=1e-5, num_layers=1, n_classes=55,
                        down_blocks=[[(64, 128, 192, 256, 384, 384), (True, True, True, True, False), (None, None, None, 8, 8)]],
                        up_blocks=[[(384, 384, 256), (True, True), (8, 8), (3, 3)],
                                   [(256, 192, 128, 64), (True, True, False), (None, None, None), (3, 3, 3)]]),
    # experiments/SuperResolution.ipynb cell 12: 4-channel points (xyz + mask), three strided levels, no attention
    "SUPERRES": dict(point_channels=4, voxel_size=0.1, pres=1e-5, num_layers=1,
                     down_blocks=[[(64, 128, 192, 256, 256), (True, True, True, False), (None, None, None, None)]],
                     up_blocks=[[(256, 256, 192), (True, True), (None, None), (3, 3)],
                                [(192, 128, 64), (True, False), (None, None), (3, 3)]]),
    "COND_TINY": dict(point_channels=3, voxel_size=0.1, pres=1e-5, num_layers=1, n_classes=7,
                      down_blocks=[[(32, 32, 64, 64), (True, True, False), (None, 4, 4)]],
                      up_blocks=[[(64, 64), (True,), (4,), (3,)],
                                 [(64, 32, 32), (True, False), (None, None), (3, 3)]]),
    # a narrow topology with the same structure (strided + k1 convs, attention, three point branches)
    # used by fast CPU tests and the committed golden vectors
    "SPVD_TINY": dict(point_channels=3, voxel_size=0.1, pres=1e-5, num_layers=1,
                      down_blocks=[[(32, 32, 64, 64), (True, True, False), (None, 4, 4)]],
                      up_blocks=[[(64, 64), (True,), (4,), (3,)],
                                 [(64, 32, 32), (True, False), (None, None), (3, 3)]]),
}


def get_config(name_or_cfg):
    return dict(PRESETS[name_or_cfg]) if isinstance(name_or_cfg, str) else dict(name_or_cfg)


# ------------------------------------------------------------------------------- parameter spec
def _bn(spec, p, c):
    spec[p + "weight"] = (c,)
    spec[p + "bias"] = (c,)
    spec[p + "running_mean"] = (c,)
    spec[p + "running_var"] = (c,)
    spec[p + "num_batches_tracked"] = ()


def _linear(spec, p, ni, nf, bias=True):
    spec[p + "weight"] = (nf, ni)
    if bias:
        spec[p + "bias"] = (nf,)


def _conv(spec, p, ni, nf, ks, bias=False):
    spec[p + "kernel"] = (ks ** 3, ni, nf) if ks > 1 else (ni, nf)
    if bias:
        spec[p + "bias"] = (nf,)


def _point_block(spec, p, ni, nf):                      # models/spvd.py:16-28
    _bn(spec, p + "conv.0.", ni)
    _linear(spec, p + "conv.2.", ni, nf)


def _res_block(spec, p, n_emb, ni, nf, ks, attn):       # models/ddpm_unet_attn.py:32-46
    _bn(spec, p + "conv1.0.", ni)
    _conv(spec, p + "conv1.2.", ni, nf, ks, bias=True)
    _bn(spec, p + "conv2.0.", nf)
    _conv(spec, p + "conv2.2.", nf, nf, ks, bias=True)
    _linear(spec, p + "t_emb.proj_mlp.", n_emb, 2 * nf)
    if ni != nf:
        _linear(spec, p + "idconv.", ni, nf)
    if attn:
        spec[p + "attn.norm1.weight"] = (nf,)
        spec[p + "attn.norm1.bias"] = (nf,)
        _linear(spec, p + "attn.attn.qkv.", nf, 3 * nf, bias=False)
        _linear(spec, p + "attn.attn.proj.", nf, nf)


def param_spec(cfg) -> "OrderedDict[str, tuple]":
    """Names and shapes of SPVUnet.state_dict() (models/spvd.py:378-420), in module order."""
    cfg = get_config(cfg)
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    pc = cfg["point_channels"]
    nf0 = cfg["down_blocks"][0][0][0]
    n_emb = nf0 * 4
    nl = cfg["num_layers"]
    _bn(spec, "emb_mlp.0.0.", nf0)
    _linear(spec, "emb_mlp.0.2.", nf0, n_emb)
    _linear(spec, "emb_mlp.1.1.", n_emb, n_emb)
    if cfg.get("n_classes"):
        spec["cond_emb.weight"] = (cfg["n_classes"], n_emb)
    _conv(spec, "stem_conv.voxel_conv.0.", pc, nf0, 3)
    _bn(spec, "stem_conv.voxel_conv.1.", nf0)
    _conv(spec, "stem_conv.voxel_conv.3.", nf0, nf0, 3)
    _point_block(spec, "stem_conv.point_conv.", pc, nf0)
    for si, (nfs, add_down, attn) in enumerate(cfg["down_blocks"]):
        for j in range(len(nfs) - 1):
            ni, nf = nfs[j], nfs[j + 1]
            p = f"down_stages.{si}.downs.{j}."
            for l in range(nl):
                _res_block(spec, p + f"resnets.{l}.", n_emb, ni, ni, 3, attn[j])
            _conv(spec, p + "down.", ni, nf, 2 if add_down[j] else 1)
        _point_block(spec, f"down_stages.{si}.point_block.", nfs[0], nfs[-1])
    for si, (nfs, add_up, attn, ks) in enumerate(cfg["up_blocks"]):
        prev = nfs[0]
        for j in range(len(nfs) - 1):
            ni = nfs[j + 1]
            p = f"up_stages.{si}.ups.{j}."
            for l in range(nl + 1):
                _res_block(spec, p + f"resnets.{l}.", n_emb, prev + ni, prev, ks[j], attn[j])
            _conv(spec, p + "up.", prev, ni, 2 if add_up[j] else 1)
            prev = ni
        _point_block(spec, f"up_stages.{si}.point_block.", nfs[0], nfs[-1])
    c_out = cfg["up_blocks"][-1][0][-1]
    _bn(spec, "out_conv.0.", c_out)
    _linear(spec, "out_conv.2.", c_out, pc, bias=False)
    return spec


def num_params(cfg) -> int:
    return sum(int(np.prod(s)) for k, s in param_spec(cfg).items()
               if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))


def make_state_dict(cfg, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic, platform-independent random-init weights (numpy PCG64 keyed by the parameter
    name), scaled like the reference's initialisers so activations stay O(1).  BN
    running statistics are made non-trivial so that eval-mode parity exercises them."""
    sd = OrderedDict()
    for name, shape in param_spec(cfg).items():
        rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
        leaf = name.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            v = np.zeros((), dtype=np.int64)
        elif leaf == "running_mean":
            v = rng.uniform(-0.1, 0.1, shape)
        elif leaf == "running_var":
            v = rng.uniform(0.8, 1.25, shape)
        elif leaf == "kernel":
            fan = shape[-2] * (shape[0] if len(shape) == 3 else 1)
            v = rng.uniform(-1, 1, shape) / math.sqrt(fan)
        elif leaf == "weight" and len(shape) == 1:        # norm scale
            v = rng.uniform(0.8, 1.2, shape)
        elif leaf == "weight":
            v = rng.uniform(-1, 1, shape) / math.sqrt(shape[1])
        else:                                              # biases
            v = rng.uniform(-0.1, 0.1, shape)
        sd[name] = torch.from_numpy(np.asarray(v)).to(torch.int64 if leaf == "num_batches_tracked" else torch.float32)
    return sd


# --------------------------------------------------------------------------------- input format
def torch2sparse(pts: torch.Tensor, pres: float = 1e-5):
    """SparseSchedulerGPU.torch2sparse (utils/schedulers.py:249-257) + batch_sparse_quantize_torch
    (models/sparse_utils.py:217-258): per-cloud min shift, floor(c/pres) -> int32, de-duplicate per
    cloud keeping the FIRST point of each cell, rows ordered by (batch, ravelled cell hash).
    pts [B,N,3] fp32.  Returns (coords int32 [M,4] = (b,x,y,z), feats fp32 [M,3], indices int64 [M])."""
    B, N, _ = pts.shape
    c = pts[..., :3]
    c = c - c.min(dim=1, keepdim=True).values
    vs = torch.tensor((pres,) * 3)
    ci = torch.floor(c / vs).int()                                       # models/sparse_utils.py:235
    x = ci - ci.min(dim=1, keepdim=True).values                          # batched_ravel_hash_torch :203-214
    x = x.long()
    xmax = x.max(dim=1, keepdim=True).values + 1
    h = torch.zeros((B, N), dtype=torch.long)
    for k in range(2):
        h += x[:, :, k]
        h *= xmax[:, :, k + 1]
    h += x[:, :, 2]
    b = torch.arange(B).repeat_interleave(N)
    key = torch.stack([b, h.reshape(-1)], dim=1)
    # unique rows sorted by (b,h); first occurrence index of each (models/sparse_utils.py:138-160)
    order = np.lexsort((np.arange(B * N), key[:, 1].numpy(), key[:, 0].numpy()))
    ks = key[order]
    first = np.ones(len(order), dtype=bool)
    first[1:] = (ks[1:] != ks[:-1]).any(dim=1).numpy()
    idx = torch.from_numpy(order[first])
    coords = torch.cat([b.view(-1, 1).int(), ci.view(-1, 3)], dim=1)[idx]
    feats = pts.reshape(-1, pts.shape[-1])[idx]
    return coords, feats, idx


# ------------------------------------------------------------------------ point <-> voxel plumbing
class Level:
    """Voxel set of one tensor stride plus the maps built on it."""

    def __init__(self, coords: np.ndarray, stride: int):
        self.coords, self.stride = coords, stride
        self.kmap3 = None            # [Nv,27]
        self.down = None             # (coarse Level, parent, koff, map_down [Nc,8], map_up [Nv,8])

    @property
    def n(self):
        return self.coords.shape[0]


class Geometry:
    """Everything the forward needs that depends on coordinates only (no features): voxel levels,
    kernel maps, point->voxel indices and trilinear weights.  Mirrors what the reference caches in
    ``_caches`` (models/sparse_utils.py:71,90,122-125; torchsparse kmaps/cmaps)."""

    def __init__(self, pc_int: torch.Tensor, pres, voxel_size, downsample_mode="spconv", scale_mode="mul"):
        pc = pc_int.float().numpy()                                      # models/spvd.py:436
        fx = ops.scale_point_coords(pc[:, 1:], pres, voxel_size, scale_mode)   # models/sparse_utils.py:61-63
        self.batch = pc[:, 0].astype(np.int32)
        self.fcoords = fx                                                # z.C[:,1:] after :72
        icoords = np.concatenate([self.batch[:, None], np.floor(fx).astype(np.int32)], 1)  # :65
        vox, idx = ops.unique_voxels(icoords)                            # :66-67
        self.levels = {1: Level(vox, 1)}
        self.mode = downsample_mode
        self.p2v = {1: idx}                                              # idx_query cache (:71)
        self.devox = {}

    def level(self, s):
        return self.levels[s]

    def kmap3(self, s):
        lv = self.levels[s]
        if lv.kmap3 is None:
            lv.kmap3 = ops.kmap_subm(lv.coords, 3)
        return lv.kmap3

    def downsample(self, s):
        lv = self.levels[s]
        if lv.down is None:
            coarse, parent, koff = ops.downsample_coords(lv.coords, self.mode)
            if coarse.shape[0] == 0:
                raise RuntimeError(f"empty voxel level at tensor stride {2 * s}")
            md = ops.kmap_down(coarse.shape[0], parent, koff)
            mu = np.full((lv.n, 8), -1, dtype=np.int32)
            rows = np.nonzero(parent >= 0)[0]
            mu[rows, koff[rows]] = parent[rows]
            self.levels.setdefault(2 * s, Level(coarse, 2 * s))
            lv.down = (self.levels[2 * s], parent, koff, md, mu)
        return lv.down

    def point_to_voxel_idx(self, s):
        """models/sparse_utils.py:79-93 (including the clamp of misses to row 0)."""
        if s not in self.p2v:
            q = np.concatenate([self.batch[:, None], np.floor(self.fcoords / np.float32(s)).astype(np.int32)], 1)
            idx = ops.hash_query(q, self.levels[s].coords, 1).reshape(-1)
            self.p2v[s] = idx
        self.p2v[s] = np.maximum(self.p2v[s], 0)
        return self.p2v[s]

    def devox_idx_weights(self, s):
        """models/sparse_utils.py:108-114."""
        if s not in self.devox:
            pf = (self.fcoords / np.float32(s)).astype(np.float32)
            q = np.concatenate([self.batch[:, None], np.floor(pf).astype(np.int32)], 1)
            idx = ops.hash_query(q, self.levels[s].coords, 2)
            w = ops.calc_ti_weights(torch.from_numpy(pf), torch.from_numpy(idx), scale=1)
            self.devox[s] = (torch.from_numpy(idx), w)
        return self.devox[s]


# ---------------------------------------------------------------------------------- the network
class _Net:
    def __init__(self, sd, cfg, geo: Geometry, training: bool, tf32: bool, trace=None):
        self.sd, self.cfg, self.geo, self.training, self.tf32, self.trace = sd, cfg, geo, training, tf32, trace

    def rec(self, name, t):
        if self.trace is not None:
            self.trace[name] = t.detach().clone()

    # -- dense glue
    def bn(self, p, x):
        sd = self.sd
        if self.training:
            return F.batch_norm(x, None, None, sd[p + "weight"], sd[p + "bias"], True, 0.0, 1e-5)
        return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                            False, 0.0, 1e-5)

    def linear(self, p, x):
        return F.linear(x, self.sd[p + "weight"], self.sd.get(p + "bias"))

    # -- sparse ops
    def conv(self, p, x, s, kind):
        """kind: 'k3' submanifold, 'down' k2s2, 'up' k2s2 transposed, 'k1'.  Returns (feats, stride)."""
        w, b = self.sd[p + "kernel"], self.sd.get(p + "bias")
        if kind == "k1":
            y = (ops.round_tf32(x) @ ops.round_tf32(w)) if self.tf32 else x @ w
            return (y if b is None else y + b), s
        if kind == "k3":
            return ops.conv_gather_gemm_scatter(x, w, self.geo.kmap3(s), x.shape[0], b, self.tf32), s
        if kind == "down":
            coarse, _, _, md, _ = self.geo.downsample(s)
            return ops.conv_gather_gemm_scatter(x, w, md, coarse.n, b, self.tf32), 2 * s
        fine = self.geo.level(s // 2)
        mu = fine.down[4]
        return ops.conv_gather_gemm_scatter(x, w, mu, fine.n, b, self.tf32), s // 2

    def point_block(self, p, z1, z):                       # models/spvd.py:26-28
        return z1 + self.linear(p + "conv.2.", F.silu(self.bn(p + "conv.0.", z)))

    def voxel_to_point(self, x, s):                        # models/sparse_utils.py:103-134
        idx, w = self.geo.devox_idx_weights(s)
        return ops.spdevoxelize(x, idx, w)

    def point_to_voxel(self, zf, s):                       # models/sparse_utils.py:78-98
        idx = torch.from_numpy(self.geo.point_to_voxel_idx(s))
        return ops.scatter_mean(zf, idx, None)

    def attention(self, p, x, batch, heads):               # models/modules/transformer.py:40-67,118-125
        nb = int(batch.max()) + 1
        cnt = torch.bincount(batch, minlength=nb)
        nmax = int(cnt.max())
        start = torch.cumsum(cnt, 0) - cnt
        pos = torch.arange(x.shape[0]) - start[batch]
        dense = x.new_zeros((nb, nmax, x.shape[1])).index_put((batch, pos), x)
        mask = torch.zeros((nb, nmax), dtype=torch.bool).index_put((batch, pos), torch.tensor(True))
        h = F.layer_norm(dense, (x.shape[1],), self.sd[p + "norm1.weight"], self.sd[p + "norm1.bias"], 1e-5)
        B, N, C = h.shape
        qkv = F.linear(h, self.sd[p + "attn.qkv.weight"]).reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        a = (q @ k.transpose(-2, -1)) * (C // heads) ** -0.5
        e = a.exp() * mask[:, None, None, :]               # models/utils.py:22-33 (no max subtraction)
        a = e / e.sum(-1, keepdim=True)
        o = (a @ v).transpose(1, 2).reshape(B, N, C)
        dense = dense + self.linear(p + "attn.proj.", o)
        return dense[mask]

    def res_block(self, p, x_in, s, emb, batch, heads):    # models/ddpm_unet_attn.py:48-65
        x, _ = self.conv(p + "conv1.2.", F.silu(self.bn(p + "conv1.0.", x_in)), s, "k3")
        tt = self.linear(p + "t_emb.proj_mlp.", F.silu(emb))[batch]          # models/modules/graph.py:19-41
        scale, shift = torch.chunk(tt, 2, dim=-1)
        x = (1 + scale) * x + shift
        x, _ = self.conv(p + "conv2.2.", F.silu(self.bn(p + "conv2.0.", x)), s, "k3")
        x = x + (self.linear(p + "idconv.", x_in) if (p + "idconv.weight") in self.sd else x_in)
        if heads:
            x = self.attention(p + "attn.", x, batch, heads)
        self.rec(p[:-1], x)
        return x

    def batch_of(self, s):
        return torch.from_numpy(self.geo.level(s).coords[:, 0].astype(np.int64))

    def forward(self, feats, t, c=None):
        cfg, sd = self.cfg, self.sd
        nf0 = cfg["down_blocks"][0][0][0]
        nl = cfg["num_layers"]
        # timestep embedding, models/utils.py:15-19 (linspace(0,1,d/2) INCLUDES 1.0)
        exponent = -math.log(10000) * torch.linspace(0, 1, nf0 // 2)
        e = t[:, None].float() * exponent.exp()[None, :]
        e = torch.cat([e.sin(), e.cos()], dim=-1)
        emb = self.linear("emb_mlp.0.2.", F.silu(self.bn("emb_mlp.0.0.", e)))
        emb = self.linear("emb_mlp.1.1.", F.silu(emb))
        if c is not None:                                  # CondSPVUnet: emb = emb_mlp(t) + cond_emb(c)
            emb = emb + sd["cond_emb.weight"][c]
        self.rec("emb", emb)
        # initial voxelization, models/sparse_utils.py:60-73
        z = feats
        x = ops.scatter_mean(z, torch.from_numpy(self.geo.p2v[1]), None)
        self.rec("x0", x)
        s = 1
        # stem, models/spvd.py:51-62
        x, _ = self.conv("stem_conv.voxel_conv.0.", x, s, "k3")
        x = F.silu(self.bn("stem_conv.voxel_conv.1.", x))
        x, _ = self.conv("stem_conv.voxel_conv.3.", x, s, "k3")
        z1 = self.point_block("stem_conv.point_conv.", self.voxel_to_point(x, s), z)
        x = self.point_to_voxel(z1, s)
        z = z1
        self.rec("stem.x", x)
        self.rec("stem.z", z)
        saved = [(x, s)]
        for si, (nfs, add_down, attn) in enumerate(cfg["down_blocks"]):   # models/spvd.py:143-159
            for j in range(len(nfs) - 1):
                p = f"down_stages.{si}.downs.{j}."
                for l in range(nl):
                    x = self.res_block(p + f"resnets.{l}.", x, s, emb, self.batch_of(s), attn[j])
                    saved.append((x, s))
                if add_down[j]:
                    x, s = self.conv(p + "down.", x, s, "down")
                    saved.append((x, s))
                else:
                    x, s = self.conv(p + "down.", x, s, "k1")
            z1 = self.point_block(f"down_stages.{si}.point_block.", self.voxel_to_point(x, s), z)
            x = self.point_to_voxel(z1, s)
            z = z1
            self.rec(f"down_stages.{si}.x", x)
            self.rec(f"down_stages.{si}.z", z)
        for si, (nfs, add_up, attn, ks) in enumerate(cfg["up_blocks"]):   # models/spvd.py:258-271
            for j in range(len(nfs) - 1):
                p = f"up_stages.{si}.ups.{j}."
                for l in range(nl + 1):
                    skip, ss = saved.pop()
                    assert ss == s, (ss, s)
                    x = self.res_block(p + f"resnets.{l}.", torch.cat([x, skip], dim=1), s, emb,
                                       self.batch_of(s), attn[j])
                x, s = self.conv(p + "up.", x, s, "up" if add_up[j] else "k1")
            z1 = self.point_block(f"up_stages.{si}.point_block.", self.voxel_to_point(x, s), z)
            if si + 1 < len(cfg["up_blocks"]):
                x = self.point_to_voxel(z1, s)               # the last one is dead work (models/spvd.py:269)
            z = z1
            self.rec(f"up_stages.{si}.z", z)
        assert not saved
        return self.linear("out_conv.2.", F.silu(self.bn("out_conv.0.", z)))    # models/spvd.py:457


def forward(sd, cfg, coords: torch.Tensor, feats: torch.Tensor, t: torch.Tensor, c: torch.Tensor = None, *,
            training=False, downsample_mode="spconv", scale_mode="mul", emulate_tf32=False, trace=None, geometry=None):
    """SPVUnet.forward((SparseTensor(feats, coords), t)) on the CPU.  coords int32 [N,4]=(b,x,y,z) in
    units of ``pres``; feats [N,point_channels]; t int64 [B].  Returns the predicted noise [N,3]."""
    cfg = get_config(cfg)
    geo = geometry or Geometry(coords, cfg["pres"], cfg["voxel_size"], downsample_mode, scale_mode)
    return _Net(sd, cfg, geo, training, emulate_tf32, trace).forward(feats, t, c)
