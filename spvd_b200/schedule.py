"""DDPM / DDIM coefficient tables (utils/schedulers.py:38-112) -- plain torch-CPU + floats, no CUDA library.

Kept free of any ``spvd_b200`` import so that ``bench.py --impl reference`` (the CPU arm) can load this file by
path without mapping libspvd_b200.so into the process.
"""
from __future__ import annotations

import math

import torch


class DDPM:
    """utils/schedulers.py:38-89 (linear / 'warm0.1' beta schedules, sigma = sqrt(beta) or the posterior std)."""

    def __init__(self, beta_min=0.0001, beta_max=0.02, n_steps=1000, mode="linear", sigma="bt"):
        self.n_steps = n_steps
        if mode == "linear":
            beta = torch.linspace(beta_min, beta_max, n_steps)
        elif mode == "warm0.1":
            beta = beta_max * torch.ones(n_steps)
            w = int(0.1 * n_steps)
            beta[:w] = torch.linspace(beta_min, beta_max, w)
        else:
            raise NotImplementedError(mode)
        alpha = 1.0 - beta
        alpha_hat = torch.cumprod(alpha, dim=0)
        if sigma == "bt":
            sig = beta.sqrt()
        elif sigma == "coef_bt":
            prev = torch.ones_like(alpha_hat)
            prev[1:] = alpha_hat[:-1]
            sig = torch.sqrt(beta * (1 - prev) / (1 - alpha_hat))
        else:
            raise ValueError(sigma)
        self.beta, self.alpha, self.alpha_hat, self.sigma = beta, alpha, alpha_hat, sig
        self.steps = list(range(n_steps - 1, -1, -1))
        # x_{t-1} = c1[t] * (x_t - c2[t] * eps) + sigma[t] * z
        self.c1 = [1.0 / math.sqrt(float(a)) for a in alpha]
        self.c2 = [float((1 - a) / math.sqrt(1 - ah)) for a, ah in zip(alpha, alpha_hat)]
        self.sig = [float(s) for s in sig]

    def update(self, x_t, eps, t, i, noise_fn):
        z = noise_fn(x_t.shape) if t > 0 else None          # no noise on the last step
        out = (x_t - self.c2[t] * eps) * self.c1[t]
        return out if z is None else out.add_(z, alpha=self.sig[t])


class DDIM(DDPM):
    """utils/schedulers.py:91-112."""

    def __init__(self, beta_min=0.0001, beta_max=0.02, n_steps=1000, mode="linear", s_steps=100, s_mode="linear"):
        super().__init__(beta_min, beta_max, n_steps, mode)
        if s_mode != "linear":
            raise NotImplementedError(s_mode)
        inds = torch.floor(torch.linspace(0, n_steps - 1, s_steps + 1)).long()
        self.prev = list(reversed(inds[:-1].tolist()))
        self.steps = list(reversed(inds[1:].tolist()))

    def update(self, x_t, eps, t, i, noise_fn):
        ah, ah1 = float(self.alpha_hat[t]), float(self.alpha_hat[self.prev[i]])
        return math.sqrt(ah1) * ((x_t - math.sqrt(1 - ah) * eps) / math.sqrt(ah)) + math.sqrt(1 - ah1) * eps
