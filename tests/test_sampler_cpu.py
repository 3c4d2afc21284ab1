"""CPU: host logic of the sampler (pure torch, no kernels): the reference-order re-sparsification reproduces
utils/schedulers.py:249-257 + models/sparse_utils.py:217-258 (golden coords/feats), and the DDPM coefficient
tables reproduce utils/schedulers.py:64-89."""
import math

import numpy as np
import torch


def test_torch2sparse_reference_order_matches_golden(golden):
    from spvd_b200.sampler import torch2sparse
    for name in ("ops_b2_n256", "tiny_eval_spconv", "spvd_train_floor"):
        g = golden(name)
        st = torch2sparse(torch.from_numpy(g["pts"]), 1e-5, order="reference")
        assert np.array_equal(st.C.numpy(), g["coords"])
        assert np.array_equal(st.F.numpy(), g["feats"])
        keep = torch2sparse(torch.from_numpy(g["pts"]), 1e-5, order="keep")     # same set of rows, input order
        a = {tuple(r) for r in keep.C.numpy().tolist()}
        assert a == {tuple(r) for r in g["coords"].tolist()}


def test_ddpm_update_matches_reference_formula():
    from spvd_b200.sampler import DDPM
    s = DDPM()
    beta = torch.linspace(0.0001, 0.02, 1000)
    alpha = 1 - beta
    ahat = torch.cumprod(alpha, 0)
    x, eps, z = torch.randn(5, 3), torch.randn(5, 3), torch.randn(5, 3)
    for t in (999, 500, 1):
        want = 1 / math.sqrt(alpha[t]) * (x - (1 - alpha[t]) / math.sqrt(1 - ahat[t]) * eps) + beta[t].sqrt() * z
        got = s.update(x, eps, t, 0, lambda shape: z)
        assert torch.allclose(got, want, atol=1e-6)
    assert torch.allclose(s.update(x, eps, 0, 0, lambda shape: z), 1 / math.sqrt(alpha[0]) * (x - (1 - alpha[0]) / math.sqrt(1 - ahat[0]) * eps), atol=1e-6)
    assert s.steps[0] == 999 and s.steps[-1] == 0 and len(s.steps) == 1000
