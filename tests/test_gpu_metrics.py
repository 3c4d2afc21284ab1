"""Chamfer / approximate EMD kernels against the CPU restatement in oracle/metrics.py."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,n,m", [(3, 500, 700), (2, 2048, 2048), (1, 37, 1500)])
def test_chamfer_forward_backward(B, n, m):
    from oracle import metrics as om
    from spvd_b200 import metrics as gm
    g = torch.Generator().manual_seed(n + m)
    a, b = torch.randn(B, n, 3, generator=g), torch.randn(B, m, 3, generator=g)
    d1, d2, i1, i2 = om.chamfer(a, b)
    ad, bd = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    g1, g2, j1, j2 = gm.chamfer_3DDist()(ad, bd)
    assert np.allclose(g1.detach().cpu().numpy(), d1.numpy(), rtol=1e-5, atol=1e-7)
    assert np.allclose(g2.detach().cpu().numpy(), d2.numpy(), rtol=1e-5, atol=1e-7)
    # argmin: identical except on exact ties; the distance at the returned index is the minimum
    assert (j1.cpu() == i1).float().mean() > 0.999 and (j2.cpu() == i2).float().mean() > 0.999
    picked = ((a - torch.gather(b, 1, j1.cpu().long()[..., None].expand(-1, -1, 3))) ** 2).sum(-1)
    assert np.allclose(picked.numpy(), d1.numpy(), rtol=1e-5, atol=1e-7)
    # gradients of the reference's loss (mean + mean) against autograd through the oracle's gather formulation
    ac, bc = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    dd = ((ac[:, :, None] - bc[:, None]) ** 2).sum(-1)
    (dd.min(2)[0].mean() + dd.min(1)[0].mean()).backward()
    gm.ChamferDistanceL2()(ad, bd).backward()
    assert rel_err(ad.grad.cpu().numpy(), ac.grad.numpy()) < 1e-5 and rel_err(bd.grad.cpu().numpy(), bc.grad.numpy()) < 1e-5


@pytest.mark.parametrize("B,n,m", [(2, 256, 256), (3, 512, 256), (1, 300, 900), (2, 1024, 1024)])
def test_emd_match_cost_and_gradients(B, n, m):
    from oracle import metrics as om
    from spvd_b200 import metrics as gm
    g = torch.Generator().manual_seed(7 * n + m)
    a, b = torch.rand(B, n, 3, generator=g) - 0.5, torch.rand(B, m, 3, generator=g) - 0.5
    match = om.approxmatch(a, b)
    want = om.matchcost(a, b, match)
    ad, bd = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    cost = gm.earth_mover_distance(ad, bd, transpose=False)
    assert rel_err(cost.detach().cpu().numpy(), want.numpy()) < 1e-3          # __expf vs exp over ten levels
    # the matching is (nearly) a transport plan: every point of the smaller-multiplicity side is matched once
    ws = torch.empty(gm.lib.spvd_emd_workspace_floats(B, n, m), device="cuda")
    mt = torch.empty(B, m, n, device="cuda")
    gm.check(gm.lib.spvd_emd_approxmatch(gm._ptr(ad.detach()), gm._ptr(bd.detach()), B, n, m, gm._ptr(mt), gm._ptr(ws), gm._stream()))
    assert rel_err(mt.cpu().numpy(), match.numpy()) < 2e-3
    multi_l = 1.0 if n >= m else float(m // n)
    assert np.allclose(mt.sum(1).cpu().numpy(), multi_l, atol=2e-2)
    # gradients: matchcost is bilinear in the squared distances with the matching held fixed
    cost.sum().backward()
    ac, bc = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    om.matchcost(ac, bc, match).sum().backward()
    assert rel_err(ad.grad.cpu().numpy(), ac.grad.numpy()) < 2e-3 and rel_err(bd.grad.cpu().numpy(), bc.grad.numpy()) < 2e-3
    # transposed input convention of the reference ([B,3,n])
    c2 = gm.EMD(a.cuda().transpose(1, 2).contiguous(), b.cuda().transpose(1, 2).contiguous())
    assert torch.allclose(c2, cost.detach(), rtol=1e-6)


def test_metrics_refuse_cpu_tensors():
    from spvd_b200 import metrics as gm
    with pytest.raises(RuntimeError):
        gm.chamfer_3DDist()(torch.zeros(1, 8, 3), torch.zeros(1, 8, 3))


def test_reference_extension_names_resolve_to_this_library():
    """`import chamfer` / `import emd_cuda` (the reference's compiled extensions) after spvd_b200.compat.install()"""
    import spvd_b200.compat as compat
    from spvd_b200 import metrics as gm
    compat.install()
    import chamfer, emd_cuda
    g = torch.Generator().manual_seed(3)
    a, b = torch.rand(2, 300, 3, generator=g).cuda(), torch.rand(2, 400, 3, generator=g).cuda()
    d1, d2, i1, i2 = chamfer.forward(a, b)
    w1, w2, j1, j2 = gm.chamfer_3DDist()(a, b)
    assert torch.equal(d1, w1) and torch.equal(d2, w2) and torch.equal(i1, j1) and torch.equal(i2, j2)
    g1, g2 = chamfer.backward(a, b, i1, i2, torch.ones_like(d1), torch.ones_like(d2))
    assert g1.shape == a.shape and g2.shape == b.shape and torch.isfinite(g1).all()
    match = emd_cuda.approxmatch_forward(a, b)
    cost = emd_cuda.matchcost_forward(a, b, match)
    assert torch.allclose(cost, gm.EMD(a, b, transpose=False), rtol=1e-6)
    ga, gb = emd_cuda.matchcost_backward(torch.ones_like(cost), a, b, match)
    assert ga.shape == a.shape and gb.shape == b.shape


@pytest.mark.parametrize("B,n,m", [(2, 512, 512), (3, 700, 300), (1, 2048, 2048)])
def test_against_the_reference_cuda_extensions(B, n, m):
    """Row N4 pinned against the REFERENCE ITSELF: metrics/chamfer_dist/chamfer.cu and metrics/PyTorchEMD/cuda/emd_kernel.cu,
    compiled unmodified from /root/reference by oracle/build_ref_metrics.py into the git-ignored oracle/_ref/ (built by
    __graft_entry__.build() where the reference exists; absent -> skip), run on the same inputs as csrc/metrics.cu."""
    from oracle.build_ref_metrics import load
    from spvd_b200 import metrics as gm
    ref_ch, ref_emd = load("ref_chamfer"), load("ref_emd_cuda")
    if ref_ch is None or ref_emd is None:
        pytest.skip("oracle/_ref is not built (needs /root/reference at build time)")
    g = torch.Generator().manual_seed(11 * n + m)
    a, b = (torch.rand(B, n, 3, generator=g) - 0.5).cuda(), (torch.rand(B, m, 3, generator=g) - 0.5).cuda()
    # Chamfer: distances bit-for-bit comparable (same fp32 formula), argmin identical except on exact ties
    d1, d2, i1, i2 = ref_ch.forward(a, b)
    w1, w2, j1, j2 = gm.chamfer_3DDist()(a, b)
    assert np.allclose(w1.cpu().numpy(), d1.cpu().numpy(), rtol=1e-5, atol=1e-7)
    assert np.allclose(w2.cpu().numpy(), d2.cpu().numpy(), rtol=1e-5, atol=1e-7)
    assert (j1 == i1).float().mean() > 0.999 and (j2 == i2).float().mean() > 0.999
    gd1, gd2 = torch.rand_like(d1), torch.rand_like(d2)
    r1, r2 = ref_ch.backward(a, b, i1, i2, gd1, gd2)
    ad, bd = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    o1, o2, _, _ = gm.chamfer_3DDist()(ad, bd)
    (o1 * gd1).sum().add((o2 * gd2).sum()).backward()
    assert rel_err(ad.grad.cpu().numpy(), r1.cpu().numpy()) < 1e-4 and rel_err(bd.grad.cpu().numpy(), r2.cpu().numpy()) < 1e-4
    # approximate EMD: matching, cost and gradients
    match = ref_emd.approxmatch_forward(a, b)
    cost = ref_emd.matchcost_forward(a, b, match)
    ad, bd = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ours = gm.earth_mover_distance(ad, bd, transpose=False)
    assert rel_err(ours.detach().cpu().numpy(), cost.cpu().numpy()) < 1e-3
    ws = torch.empty(gm.lib.spvd_emd_workspace_floats(B, n, m), device="cuda")
    mt = torch.empty(B, m, n, device="cuda")
    gm.check(gm.lib.spvd_emd_approxmatch(gm._ptr(a), gm._ptr(b), B, n, m, gm._ptr(mt), gm._ptr(ws), gm._stream()))
    assert rel_err(mt.cpu().numpy(), match.cpu().numpy()) < 2e-3
    gc = torch.rand_like(cost)
    ra, rb = ref_emd.matchcost_backward(gc, a, b, match)
    (ours * gc).sum().backward()
    assert rel_err(ad.grad.cpu().numpy(), ra.cpu().numpy()) < 2e-3 and rel_err(bd.grad.cpu().numpy(), rb.cpu().numpy()) < 2e-3
