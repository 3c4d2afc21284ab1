"""Host-side input format of the denoiser: what the reference's dataset classes and its
completion notebook feed the hot path.  numpy / torch-CPU only -- nothing here touches the CUDA library.

  * ``NoiseSchedulerDDPM``          datasets/utils.py:6-42  (q(x_t | x_0) sample of one cloud, random timestep)
  * ``noisy_training_sample``       datasets/shapenet_pointflow_sparse.py:20-47  (+ class id: ..._sparse_cond.py:29-58)
  * ``completion_training_sample``  datasets/partnet.py:118-168  (kept parts stay clean, 4th 0/1 feature channel)
  * ``collate``                     torchsparse.utils.collate.sparse_collate_fn as the datasets use it
  * ``MaskedMSELoss``               experiments/TrainCompletion.ipynb "CustomMSELoss": MSE over the noisy points only
"""
from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn as nn

from .schedule import DDPM
from .sparse import SparseTensor


# --- torchsparse.utils.quantize (numpy): first occurrence of every occupied voxel, voxels in ravel-hash order
def ravel_hash(x: np.ndarray) -> np.ndarray:
    x = (x - x.min(axis=0)).astype(np.uint64)
    xmax = x.max(axis=0).astype(np.uint64) + 1
    h = np.zeros(x.shape[0], dtype=np.uint64)
    for k in range(x.shape[1] - 1):
        h += x[:, k]
        h *= xmax[k + 1]
    return h + x[:, -1]


def sparse_quantize(coords, voxel_size=1, *, return_index=False, return_inverse=False):
    vs = np.asarray(voxel_size if isinstance(voxel_size, (tuple, list)) else (voxel_size,) * 3)
    coords = np.floor(coords / vs).astype(np.int32)
    _, ind, inv = np.unique(ravel_hash(coords), return_index=True, return_inverse=True)
    out = [coords[ind]] + ([ind] if return_index else []) + ([inv] if return_inverse else [])
    return out[0] if len(out) == 1 else out


def sparse_collate(inputs):
    coords = [torch.cat((torch.full((x.coords.shape[0], 1), k, dtype=torch.int), x.coords.int()), dim=1)
              for k, x in enumerate(inputs)]
    return SparseTensor(torch.cat([x.feats for x in inputs], 0), torch.cat(coords, 0), inputs[0].stride)


def collate(samples):
    """list of dicts -> dict: SparseTensors are batched (batch column prepended), tensors stacked, the rest listed."""
    out = {}
    for name in samples[0]:
        if not all(name in d for d in samples):
            continue                                  # optional keys (e.g. 'label') must be present in every sample
        v = [d[name] for d in samples]
        if isinstance(v[0], SparseTensor):
            out[name] = sparse_collate(v)
        elif isinstance(v[0], torch.Tensor):
            # upstream stacks; clouds that lost points to cell collisions have ragged row counts: concatenate those
            out[name] = torch.stack(v, 0) if all(t.shape == v[0].shape for t in v) else torch.cat(v, 0)
        else:
            out[name] = v
    return out


class NoiseSchedulerDDPM:
    """``xt, t, eps = sched(x0)``: one forward-diffusion sample at a uniformly random timestep (datasets/utils.py:29-42)."""

    def __init__(self, beta_min=0.0001, beta_max=0.02, n_steps=1000, mode="linear"):
        tab = DDPM(beta_min, beta_max, n_steps, mode)
        self.n_steps = n_steps
        self.beta, self.alpha, self.alpha_hat = tab.beta.numpy(), tab.alpha.numpy(), tab.alpha_hat.numpy()

    def __call__(self, x0, t=None, eps=None):
        t = random.randint(0, self.n_steps - 1) if t is None else int(t)
        eps = torch.randn(x0.shape) if eps is None else eps
        ah = self.alpha_hat[t]
        return np.sqrt(ah) * x0 + np.sqrt(1 - ah) * eps, t, eps


def noisy_training_sample(points, noise_scheduler, pres=1e-5, label=None, t=None, eps=None):
    """One item of ShapeNet15kPointCloudsSparseNoisy: {'input': SparseTensor(x_t), 't', 'noise': SparseTensor(eps)} with the
    reference's 1e-5 integer coordinates (one row per occupied cell, first occurrence wins)."""
    pts, t, noise = noise_scheduler(torch.as_tensor(points, dtype=torch.float32), t, eps)
    pts, noise = np.asarray(pts, dtype=np.float32), np.asarray(noise, dtype=np.float32)
    coords, idx = sparse_quantize(pts - pts.min(axis=0, keepdims=True), pres, return_index=True)
    coords = torch.tensor(coords, dtype=torch.int)
    item = {"input": SparseTensor(torch.tensor(pts[idx], dtype=torch.float), coords), "t": t,
            "noise": SparseTensor(torch.tensor(noise[idx], dtype=torch.float), coords)}
    if label is not None:
        item["label"] = int(label)
    return item


def completion_training_sample(points, keep_mask, noise_scheduler, pres=1e-5, add_real_info=True, t=None, eps=None):
    """One item of the PartNet completion dataset: the kept parts (keep_mask True) stay clean, the rest is noised; with
    ``add_real_info`` the mask rides along as a 4th feature channel (datasets/partnet.py:118-168)."""
    points = torch.as_tensor(points, dtype=torch.float32)
    points = (points - points.mean(dim=0)) / points.std()
    mask = torch.as_tensor(keep_mask, dtype=torch.bool)
    noisy, t, noise = noise_scheduler(points, t, eps)
    noisy = torch.as_tensor(noisy, dtype=torch.float32).clone()
    noisy[mask] = points[mask]
    c = noisy.numpy()
    coords, idx = sparse_quantize(c - c.min(axis=0, keepdims=True), pres, return_index=True)
    feats = torch.cat([noisy, mask.unsqueeze(-1).float()], dim=-1) if add_real_info else noisy
    return {"input": SparseTensor(feats[idx], torch.tensor(coords, dtype=torch.int)), "t": t,
            "noise": torch.as_tensor(noise, dtype=torch.float32)[idx], "mask": mask[idx]}


class MaskedMSELoss(nn.Module):
    """MSE between the first three predicted channels and the noise, over the points that were actually noised
    (mask False); ``target`` = (noise [N,3], mask [N])."""

    def forward(self, preds, target):
        noise, mask = target
        keep = ~mask.view(-1).bool()
        return nn.functional.mse_loss(preds[keep, :3], noise.view(-1, 3)[keep])
