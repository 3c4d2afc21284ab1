"""Generate tests/golden/*.npz by running the REFERENCE's own Python (models/spvd.py,
models/sparse_utils.py, utils/schedulers.py -- unmodified, imported from /root/reference) on top of
the oracle's stand-ins for its un-vendored dependencies (oracle/shim.py).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Run in the build container only (the GPU box has no
/root/reference):   python -m oracle.make_golden

What this pins: the reference's composition code (module tree, parameter names, op order, caching,
clamp_(0), skip-list order, attention/FiLM glue, torch2sparse ordering).  What it cannot pin: the
arithmetic inside torchsparse / torch_scatter themselves (absent; see oracle/__init__.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import model as omodel
from . import shim
from . import tsparse as ts

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def synth_cloud(B, N, seed, kind="noise"):
    g = torch.Generator().manual_seed(seed)
    if kind == "noise":                       # SparseScheduler.create_noise, utils/schedulers.py:197-201
        return torch.randn(B, N, 3, generator=g).clamp(-3, 3)
    x = torch.randn(B, N, 3, generator=g)     # sphere of radius sqrt(3): SURVEY.md section 8d config 2(i)
    return x / x.norm(dim=-1, keepdim=True) * 3 ** 0.5


def reference_model(cfg, sd):
    import importlib
    spvd = importlib.import_module("models.spvd")
    m = spvd.SPVUnet(point_channels=cfg["point_channels"], voxel_size=cfg["voxel_size"],
                     down_blocks=cfg["down_blocks"], up_blocks=cfg["up_blocks"],
                     num_layers=cfg["num_layers"], pres=cfg["pres"])
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m


def run_model_case(name, preset, B, N, seed, kind, t_val, mode, training=False):
    from utils.schedulers import DDPMSparseSchedulerGPU
    cfg = omodel.get_config(preset)
    sd = omodel.make_state_dict(cfg, seed=seed)
    ts.conf.downsample_mode = mode
    ts.conf.scale_mode = "mul"
    ts.conf.emulate_tf32 = False
    m = reference_model(cfg, sd)
    m.train(training)
    pts = synth_cloud(B, N, seed, kind)
    xt = DDPMSparseSchedulerGPU().torch2sparse(pts, (B, N, 3))
    t = torch.full((B,), t_val, dtype=torch.long)
    if training:      # distinct timesteps per cloud (datasets/utils.py:29-42); identical rows would make the
        t = (t + 379 * torch.arange(B)) % 1000          # training-mode BatchNorm1d of emb_mlp degenerate (0/0)
    out = {"pts": pts.numpy(), "coords": xt.C.numpy().copy(), "feats": xt.F.numpy().copy(), "t": t.numpy()}
    if training:
        y = m((xt, t))
        target = synth_cloud(B, N, seed + 1, "noise").reshape(-1, 3)[: y.shape[0]]
        loss = torch.nn.functional.mse_loss(y, target)
        loss.backward()
        out["target"] = target.numpy()
        out["loss"] = np.float64(loss.item())
        names, norms = [], []
        for k, p in m.named_parameters():
            names.append(k)
            norms.append(float(p.grad.double().norm()) if p.grad is not None else -1.0)
        out["grad_names"] = np.array(names)
        out["grad_norms"] = np.array(norms)
        # two full gradients, small ones, for direction checks
        out["grad_out_w"] = dict(m.named_parameters())["out_conv.2.weight"].grad.numpy()
        out["grad_stem_k"] = dict(m.named_parameters())["stem_conv.voxel_conv.0.kernel"].grad.numpy()
    else:
        with torch.no_grad():
            y = m((xt, t))
    out["out"] = y.detach().numpy()
    out["meta"] = np.array([preset, mode, kind, str(seed), str(int(training))])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N", xt.C.shape[0], "out", y.shape, float(y.abs().mean()))
    return out


def run_ops_case(name, B, N, seed):
    """The reference's point<->voxel operators (models/sparse_utils.py:60-134) and its input
    quantiser, executed from the reference file itself."""
    import importlib
    su = importlib.import_module("models.sparse_utils")
    from utils.schedulers import DDPMSparseSchedulerGPU
    import torchsparse
    ts.conf.downsample_mode = "spconv"
    pts = synth_cloud(B, N, seed, "noise")
    xt = DDPMSparseSchedulerGPU().torch2sparse(pts, (B, N, 3))
    z = su.PointTensor(xt.F, xt.C.float())
    x0 = su.initial_voxelize(z, 1e-5, 0.1)
    out = {"pts": pts.numpy(), "coords": xt.C.numpy().copy(), "feats": xt.F.numpy().copy(),
           "vox_coords": x0.C.numpy().copy(), "vox_feats": x0.F.numpy().copy(),
           "idx_query_s1": z._caches.idx_query[(1, 1, 1)].numpy().copy(), "z_coords": z.C.numpy().copy()}
    # two stride-2 convs to get levels 2 and 4, then devoxelize / point_to_voxel at stride 4
    g = torch.Generator().manual_seed(seed)
    w1 = torch.randn(8, 3, 8, generator=g) * 0.3
    w2 = torch.randn(8, 8, 8, generator=g) * 0.3
    x2 = torchsparse.nn.functional.conv3d(x0, w1, 2, stride=2)
    x4 = torchsparse.nn.functional.conv3d(x2, w2, 2, stride=2)
    z4 = su.voxel_to_point(x4, z)
    out["s2_coords"], out["s4_coords"] = x2.C.numpy().copy(), x4.C.numpy().copy()
    out["s4_feats"] = x4.F.numpy().copy()
    out["devox_idx_s4"] = z._caches.idx_query_devox[(4, 4, 4)].numpy().copy()
    out["devox_w_s4"] = z._caches.weights_devox[(4, 4, 4)].numpy().copy()
    out["devox_feats_s4"] = z4.F.numpy().copy()
    v4 = su.point_to_voxel(x4, z4)
    out["p2v_idx_s4"] = z._caches.idx_query[(4, 4, 4)].numpy().copy()       # after the in-place clamp
    out["p2v_feats_s4"] = v4.F.numpy().copy()
    out["w1"], out["w2"] = w1.numpy(), w2.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "Nv", x0.C.shape[0], x2.C.shape[0], x4.C.shape[0])


def main():
    os.makedirs(OUT, exist_ok=True)
    shim.import_reference()
    torch.set_num_threads(4)
    run_ops_case("ops_b2_n256", 2, 256, 11)
    run_model_case("tiny_eval_spconv", "SPVD_TINY", 2, 256, 3, "noise", 999, "spconv")
    run_model_case("tiny_eval_floor", "SPVD_TINY", 2, 256, 3, "sphere", 10, "floor")
    run_model_case("tiny_train_floor", "SPVD_TINY", 4, 128, 5, "noise", 500, "floor", training=True)
    run_model_case("spvd_eval_spconv", "SPVD", 2, 256, 7, "noise", 999, "spconv")
    run_model_case("spvd_train_floor", "SPVD", 4, 128, 9, "noise", 321, "floor", training=True)


if __name__ == "__main__":
    sys.exit(main())
