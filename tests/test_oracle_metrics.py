"""CPU checks of the evaluation-distance oracle (analytic cases; the reference ships no assertions for these)."""
import torch

from oracle import metrics as om


def test_chamfer_analytic():
    a = torch.tensor([[[0., 0, 0], [1, 0, 0]]])
    b = torch.tensor([[[0., 0, 0.5], [1, 1, 0], [5, 5, 5]]])
    d1, d2, i1, i2 = om.chamfer(a, b)
    assert torch.allclose(d1, torch.tensor([[0.25, 1.0]])) and i1.tolist() == [[0, 1]]
    assert torch.allclose(d2, torch.tensor([[0.25, 1.0, 66.0]])) and i2.tolist() == [[0, 1, 1]]


def test_emd_two_point_known_answer():
    """metrics/PyTorchEMD/test_emd_loss.py:1-44 prints this case: two well-separated pairs match one to one."""
    a = torch.tensor([[[0., 0, 0], [10, 0, 0]]])
    b = torch.tensor([[[0.5, 0, 0], [10, 1, 0]]])
    match = om.approxmatch(a, b)
    assert torch.allclose(match, torch.eye(2)[None], atol=1e-4)
    assert torch.allclose(om.earth_mover_distance(a, b), torch.tensor([0.25 + 1.0]), atol=1e-3)
    # identical clouds cost nothing; the cost is symmetric in its arguments for equal sizes
    g = torch.Generator().manual_seed(0)
    x, y = torch.rand(2, 64, 3, generator=g), torch.rand(2, 64, 3, generator=g)
    assert float(om.earth_mover_distance(x, x).abs().max()) < 1e-6
    assert torch.allclose(om.matchcost(x, y, om.approxmatch(x, y)).sum(), om.matchcost(y, x, om.approxmatch(y, x)).sum(), rtol=0.2)
