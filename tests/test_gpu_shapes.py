"""Parity at the shapes the numbers are quoted on (BASELINE.json configs[1], [3], [4]): the native executor against the
CPU oracle (oracle/model.py) on the bench's own synthetic input, with both tensor-core operand formats and both
downsample modes; plus a dynamic-range guard for the f16 operand path ("Next round" item 1).

Tolerance: BASELINE.json north_star -- point features within 1e-3 relative error (fp32 accumulate).
"""
import math

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def sp():
    import spvd_b200
    return spvd_b200


def surrogate(B, N, t, seed=1234):
    """bench.py's open-loop input: x_t = sqrt(abar_t) x0 + sqrt(1 - abar_t) eps, x0 on the sphere of radius sqrt(3)."""
    import bench
    ah = float(bench._schedule().DDPM().alpha_hat[t])
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(B, N, 3, generator=g)
    x0 = x0 / x0.norm(dim=-1, keepdim=True) * math.sqrt(3.0)
    return (math.sqrt(ah) * x0 + math.sqrt(1 - ah) * torch.randn(B, N, 3, generator=g)).contiguous()


def build(sp, preset, mode, seed=0):
    from oracle import model as om
    sd = om.make_state_dict(preset, seed=seed)
    m = sp.build_model(preset, downsample_mode=mode)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


@pytest.mark.parametrize("mode", ["spconv", "floor"])
def test_config2_b32x2048_executor_vs_oracle(sp, mode):
    """SPVD, B=32 x 2048, the bench input at t = 999 / 500 / 0, f16 and TF32 operands, through the sampler's
    torch2sparse (the call path bench.py times)."""
    from oracle import model as om
    from spvd_b200.sampler import DDPMSparseSchedulerGPU
    m, sd = build(sp, "SPVD", mode)
    sched = DDPMSparseSchedulerGPU()
    B, N = 32, 2048
    for t in (999, 500, 0):
        pts = surrogate(B, N, t)
        coords, feats, _ = om.torch2sparse(pts)
        tb = torch.full((B,), t, dtype=torch.long)
        with torch.no_grad():
            want = om.forward(sd, "SPVD", coords, feats, tb, downsample_mode=mode, scale_mode="mul").numpy()
            st = sched.torch2sparse(pts.cuda(), (B, N, 3))             # the fused re-sparsification of the sampling loop:
            a = st.C.cpu().numpy()                                     # same coordinate set as the oracle's (row order differs:
            b = coords.numpy()                                         # the fused path keeps the point order)
            assert np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])
            x = sp.SparseTensor(feats.cuda(), coords.cuda())
            x.points_per_cloud = N
            for operands in ("f16", "tf32"):
                m.executor_operands = operands
                got = m((x, tb.cuda())).cpu().numpy()
                assert got.shape == want.shape
                assert rel_err(got, want) < TOL, (t, operands, rel_err(got, want))


def test_config4_cond_b64x2048_vs_oracle(sp):
    """Conditional SPVD-L (55 classes), B=64 x 2048 (experiments/ConditionalGeneration.ipynb cell 4 topology)."""
    from oracle import model as om
    m, sd = build(sp, "COND_SPVD_L", "spconv", seed=3)
    B, N, t = 64, 2048, 500
    pts = surrogate(B, N, t, seed=7)
    coords, feats, _ = om.torch2sparse(pts)
    g = torch.Generator().manual_seed(1)
    tb = torch.randint(0, 1000, (B,), generator=g)
    cls = torch.randint(0, 55, (B,), generator=g)
    with torch.no_grad():
        want = om.forward(sd, "COND_SPVD_L", coords, feats, tb, cls).numpy()
        got = m((sp.SparseTensor(feats.cuda(), coords.cuda()), tb.cuda(), cls.cuda())).cpu().numpy()
    assert rel_err(got, want) < TOL


def test_config5_b8x8192_vs_oracle(sp):
    """High-resolution clouds: SPVD, B=8 x 8192 points."""
    from oracle import model as om
    m, sd = build(sp, "SPVD", "spconv", seed=4)
    B, N, t = 8, 8192, 300
    pts = surrogate(B, N, t, seed=11)
    coords, feats, _ = om.torch2sparse(pts)
    tb = torch.full((B,), t, dtype=torch.long)
    with torch.no_grad():
        want = om.forward(sd, "SPVD", coords, feats, tb).numpy()
        for operands in ("f16", "tf32"):
            m.executor_operands = operands
            got = m((sp.SparseTensor(feats.cuda(), coords.cuda()), tb.cuda())).cpu().numpy()
            assert rel_err(got, want) < TOL, operands


@pytest.mark.parametrize("act_scale", [1e-4, 1.0, 1e3])
def test_f16_operands_dynamic_range_guard(sp, act_scale):
    """The f16 operand path keeps TF32's 11-bit significand but only a 5-bit exponent.  Scale the activations entering
    the convolutions by 1e-4 / 1e3 (through the first BatchNorm affine of every block: running_var) and spread the conv
    weights over 1e-6..1 (per-output-channel gains): the f16 executor must stay within 1e-3 of the TF32 executor and of
    the fp32 oracle -- per-slab subnormals would show up here."""
    from oracle import model as om
    torch.manual_seed(0)
    sd = om.make_state_dict("SPVD", seed=5)
    g = torch.Generator().manual_seed(9)
    for k, v in sd.items():
        if k.endswith("conv1.0.running_var") or k.endswith("conv2.0.running_var"):
            v.mul_(1.0 / act_scale ** 2)                       # BN output (the conv operand) scales by act_scale
        if k.endswith(".kernel") and v.dim() == 3 and "conv1.2" in k:
            gains = torch.logspace(-6, 0, v.shape[2])[torch.randperm(v.shape[2], generator=g)]
            v.mul_(gains[None, None, :] / act_scale)           # wide weight spread; overall magnitude compensated
    m = sp.build_model("SPVD")
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    B, N, t = 4, 2048, 500
    pts = surrogate(B, N, t, seed=3)
    coords, feats, _ = om.torch2sparse(pts)
    tb = torch.full((B,), t, dtype=torch.long)
    with torch.no_grad():
        want = om.forward(sd, "SPVD", coords, feats, tb).numpy()
        outs = {}
        for operands in ("f16", "tf32"):
            m.executor_operands = operands
            outs[operands] = m((sp.SparseTensor(feats.cuda(), coords.cuda()), tb.cuda())).cpu().numpy()
    assert np.isfinite(outs["f16"]).all()
    assert rel_err(outs["f16"], outs["tf32"]) < TOL
    assert rel_err(outs["f16"], want) < TOL
