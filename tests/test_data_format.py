"""CPU tests of the host-side input format: sparse_quantize / sparse_collate_fn as the reference's
CPU scheduler and datasets call them (utils/schedulers.py:225-240, datasets/shapenet_pointflow_sparse.py:20-47,
datasets/partnet.py:118-168), the noisy training sample (datasets/utils.py:29-42) and the masked completion loss
(experiments/TrainCompletion.ipynb, CustomMSELoss) -- each against an independent restatement written here."""
import numpy as np
import torch

from spvd_b200 import data
from spvd_b200.sparse import SparseTensor


def brute_quantize(coords, vs):
    """first occurrence of every occupied cell, cells ordered by (x, y, z) after the per-axis min shift"""
    c = np.floor(coords / vs).astype(np.int64)
    seen, order = {}, []
    for i, row in enumerate(map(tuple, c)):
        if row not in seen:
            seen[row] = i
            order.append(row)
    order.sort()
    return np.array(order, dtype=np.int32), np.array([seen[r] for r in order])


def test_sparse_quantize_return_index_matches_brute_force():
    rng = np.random.RandomState(0)
    pts = rng.rand(500, 3).astype(np.float32) * 0.01
    pts[100:150] = pts[:50] + 1e-7                                   # collisions at pres = 1e-5
    c, idx = data.sparse_quantize(pts - pts.min(0, keepdims=True), 1e-5, return_index=True)
    bc, bidx = brute_quantize(pts - pts.min(0, keepdims=True), 1e-5)
    assert c.dtype == np.int32 and c.shape == bc.shape and c.shape[0] < 500
    assert np.array_equal(c, bc) and np.array_equal(idx, bidx)
    c2, idx2, inv = data.sparse_quantize(pts, (1e-5, 2e-5, 4e-5), return_index=True, return_inverse=True)
    assert np.array_equal(c2[inv], np.floor(pts / np.array((1e-5, 2e-5, 4e-5))).astype(np.int32))


def test_cpu_scheduler_call_sequence_matches_gpu_flavour():
    """SparseSchedulerCPU.torch2sparse (utils/schedulers.py:225-247) through compat's names gives the same voxel SET and
    features as the batched torch path (our sampler.torch2sparse, order='reference')."""
    from spvd_b200.compat import sparse_collate_fn, sparse_quantize
    from spvd_b200.sampler import torch2sparse
    g = torch.Generator().manual_seed(2)
    pts = torch.randn(3, 256, 3, generator=g).clamp(-3, 3)
    coords = (pts - pts.min(dim=1, keepdim=True)[0]).numpy()
    batch = []
    for b in range(3):
        c, ind = sparse_quantize(coords[b], 1e-5, return_index=True)
        batch.append({"pc": SparseTensor(pts[b][ind], torch.tensor(c))})
    st = sparse_collate_fn(batch)["pc"]
    ref = torch2sparse(pts, 1e-5, order="reference")
    assert st.C.dtype == torch.int32 and st.C.shape[1] == 4
    # the numpy path divides in float64 (voxel_size becomes a float64 array upstream), the torch path in float32: the same
    # point may land one 1e-5 cell apart -- match rows through their features and allow that single step
    assert st.F.shape == ref.F.shape
    a = {tuple(f): c for f, c in zip(st.F.tolist(), st.C.tolist())}
    b = {tuple(f): c for f, c in zip(ref.F.tolist(), ref.C.tolist())}
    assert a.keys() == b.keys()
    diff = np.array([np.subtract(a[k], b[k]) for k in a])
    assert np.abs(diff).max() <= 1 and (diff[:, 0] == 0).all() and (diff != 0).mean() < 0.05


def test_noisy_training_sample_and_collate():
    rng = np.random.RandomState(1)
    x0 = rng.randn(2048, 3).astype(np.float32)
    ns = data.NoiseSchedulerDDPM(1e-4, 0.02, 1000)
    eps = torch.randn(2048, 3, generator=torch.Generator().manual_seed(3))
    it = data.noisy_training_sample(x0, ns, t=371, eps=eps, label=12)
    ah = float(np.cumprod(1 - np.linspace(1e-4, 0.02, 1000, dtype=np.float32))[371])
    xt = np.sqrt(ah) * x0 + np.sqrt(1 - ah) * eps.numpy()
    assert it["t"] == 371 and it["label"] == 12
    assert it["input"].C.dtype == torch.int32 and it["input"].C.min() == 0
    # every row of the sample is one of the noised points, with its own noise vector and its own cell
    n = it["input"].F.shape[0]
    assert n <= 2048 and it["noise"].F.shape == (n, 3) and torch.equal(it["noise"].C, it["input"].C)
    d = np.abs(it["input"].F.numpy()[:, None, :] - xt[None, :50, :]).sum(-1)
    assert (d.min(0) < 1e-5).all()
    cells = np.floor((it["input"].F.numpy() - xt.min(0, keepdims=True)) / 1e-5).astype(np.int32)
    assert np.abs(cells - it["input"].C.numpy()).max() <= 1            # same cell up to one float rounding step
    back = (it["input"].F.numpy() - np.sqrt(1 - ah) * it["noise"].F.numpy()) / np.sqrt(ah)
    assert (np.abs(back[:64, None, :] - x0[None, :, :]).sum(-1).min(1) < 1e-4).all()    # x_t and eps of a row belong together
    b = data.collate([it, data.noisy_training_sample(x0[:1024], ns, t=5)])
    assert b["input"].C.shape[1] == 4 and set(b["input"].C[:, 0].tolist()) == {0, 1} and b["t"] == [371, 5]
    assert b["input"].F.shape[0] == b["noise"].F.shape[0] == n + 1024


def test_completion_sample_and_masked_loss():
    rng = np.random.RandomState(4)
    pts = rng.randn(512, 3).astype(np.float32) * 2 + 1
    keep = rng.rand(512) < 0.4
    it = data.completion_training_sample(pts, keep, data.NoiseSchedulerDDPM(), t=800)
    f, m = it["input"].F, it["mask"]
    assert f.shape[1] == 4 and torch.equal(f[:, 3].bool(), m)
    norm = (torch.tensor(pts) - torch.tensor(pts).mean(0)) / torch.tensor(pts).std()
    # kept points are the clean (normalised) ones
    kept = f[m][:, :3]
    assert all((norm - p).abs().sum(-1).min() < 1e-6 for p in kept[:20])
    preds = torch.randn(f.shape[0], 4)
    loss = data.MaskedMSELoss()(preds, (it["noise"], m))
    want = ((preds[~m, :3] - it["noise"][~m]) ** 2).mean()
    assert torch.allclose(loss, want)
    # gradient flows only into the noisy rows
    preds.requires_grad_(True)
    data.MaskedMSELoss()(preds, (it["noise"], m)).backward()
    assert preds.grad[m].abs().sum() == 0 and preds.grad[~m, :3].abs().sum() > 0 and preds.grad[:, 3].abs().sum() == 0
