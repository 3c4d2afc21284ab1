"""CPU: the oracle's functional model (oracle/model.py) against the golden vectors produced by the
reference's own Python (oracle/make_golden.py).  Pins module order, parameter names, skip-list
order, clamp-to-row-0, attention / FiLM glue and the torch2sparse ordering.

The goldens were produced by torch-CPU arithmetic, whose ``(c*1e-5)/0.1`` is a true division
(scale_mode="div"); the CUDA flavour ("mul", torch multiplies by the reciprocal on the GPU) lands a few
boundary points in the neighbouring voxel, so every comparison with
these files uses scale_mode="div".  test_scale_modes_differ_only_at_boundaries covers "mul"."""
import numpy as np
import pytest
import torch

from oracle import model as om
from conftest import rel_err


def test_conditional_and_superres_specs_match_the_reference_constructors():
    """parameter names/shapes of the conditional (ConditionalGeneration.ipynb cell 4) and 4-channel super-resolution
    (SuperResolution.ipynb cell 12) variants, and the product module tree, agree with oracle.model.param_spec"""
    import spvd_b200.denoiser as d
    for preset in ("COND_TINY", "COND_SPVD_L", "SUPERRES", "SPVD", "SPVD_L"):
        m = d.build_model(preset)
        spec = om.param_spec(preset)
        sd = m.state_dict()
        assert list(sd.keys()) == list(spec.keys())
        assert all(tuple(sd[k].shape) == tuple(spec[k]) for k in spec)
    assert om.num_params("COND_SPVD_L") == 88_103_366 + 55 * 256


def test_param_counts():
    # models/__init__.py:8,24 ("32.9M", "88.1M"); experiments/ModelArchitecture.ipynb:631
    assert om.num_params("SPVD") == 32_894_054
    assert om.num_params("SPVD_L") == 88_103_366


@pytest.mark.parametrize("name", ["tiny_eval_spconv", "tiny_eval_floor", "spvd_eval_spconv"])
def test_eval_forward_matches_reference(golden, name):
    g = golden(name)
    preset, mode, _, seed, _ = g["meta"]
    sd = om.make_state_dict(preset, seed=int(seed))
    with torch.no_grad():
        y = om.forward(sd, str(preset), torch.from_numpy(g["coords"]), torch.from_numpy(g["feats"]),
                       torch.from_numpy(g["t"]), downsample_mode=str(mode), scale_mode="div")
    assert rel_err(y.numpy(), g["out"]) < 1e-5


@pytest.mark.parametrize("name", ["tiny_train_floor", "spvd_train_floor"])
def test_train_forward_backward_matches_reference(golden, name):
    g = golden(name)
    preset, mode, _, seed, _ = g["meta"]
    sd = om.make_state_dict(preset, seed=int(seed))
    params = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    y = om.forward(params, str(preset), torch.from_numpy(g["coords"]), torch.from_numpy(g["feats"]),
                   torch.from_numpy(g["t"]), training=True, downsample_mode=str(mode), scale_mode="div")
    assert rel_err(y.detach().numpy(), g["out"]) < 1e-5
    loss = torch.nn.functional.mse_loss(y, torch.from_numpy(g["target"]))
    assert abs(loss.item() - float(g["loss"])) < 1e-6 * max(1.0, abs(float(g["loss"])))
    loss.backward()
    for k, n in zip(g["grad_names"], g["grad_norms"]):
        got = float(params[str(k)].grad.double().norm())
        assert abs(got - n) <= 2e-4 * n + 1e-6, (k, got, n)  # bias before a training BN has a zero gradient (rounding noise)
    assert rel_err(params["out_conv.2.weight"].grad.numpy(), g["grad_out_w"]) < 1e-4
    assert rel_err(params["stem_conv.voxel_conv.0.kernel"].grad.numpy(), g["grad_stem_k"]) < 1e-4


def test_torch2sparse_matches_reference(golden):
    for name in ["ops_b2_n256", "tiny_eval_floor"]:
        g = golden(name)
        coords, feats, _ = om.torch2sparse(torch.from_numpy(g["pts"]))
        assert np.array_equal(coords.numpy(), g["coords"])
        assert np.array_equal(feats.numpy(), g["feats"])


def test_point_voxel_ops_match_reference(golden):
    """models/sparse_utils.py:60-134 executed from the reference file vs the oracle's Geometry."""
    g = golden("ops_b2_n256")
    geo = om.Geometry(torch.from_numpy(g["coords"]), 1e-5, 0.1, "spconv", scale_mode="div")
    assert np.array_equal(geo.level(1).coords, g["vox_coords"])
    assert np.array_equal(geo.p2v[1], g["idx_query_s1"])
    assert np.array_equal(np.concatenate([geo.batch[:, None].astype(np.float32), geo.fcoords], 1), g["z_coords"])
    geo.downsample(1)
    geo.downsample(2)
    assert np.array_equal(geo.level(2).coords, g["s2_coords"])
    assert np.array_equal(geo.level(4).coords, g["s4_coords"])
    idx, w = geo.devox_idx_weights(4)
    assert np.array_equal(idx.numpy(), g["devox_idx_s4"])
    assert np.allclose(w.numpy(), g["devox_w_s4"], atol=1e-7)
    assert np.array_equal(geo.point_to_voxel_idx(4), g["p2v_idx_s4"].reshape(-1))
    from oracle import ops
    x0 = ops.scatter_mean(torch.from_numpy(g["feats"]), torch.from_numpy(geo.p2v[1]))
    assert rel_err(x0.numpy(), g["vox_feats"]) < 1e-6
    x2 = ops.conv_gather_gemm_scatter(x0, torch.from_numpy(g["w1"]), geo.level(1).down[3], geo.level(2).n)
    x4 = ops.conv_gather_gemm_scatter(x2, torch.from_numpy(g["w2"]), geo.level(2).down[3], geo.level(4).n)
    assert rel_err(x4.numpy(), g["s4_feats"]) < 1e-5
    z4 = ops.spdevoxelize(x4, idx, w)
    assert rel_err(z4.numpy(), g["devox_feats_s4"]) < 1e-5
    v4 = ops.scatter_mean(z4, torch.from_numpy(geo.point_to_voxel_idx(4)))
    assert rel_err(v4.numpy(), g["p2v_feats_s4"]) < 1e-5


def test_scale_modes_differ_only_at_boundaries():
    from oracle import ops
    c = np.arange(0, 900001, dtype=np.float32)
    mul = np.floor(ops.scale_point_coords(c, 1e-5, 0.1, "mul"))
    div = np.floor(ops.scale_point_coords(c, 1e-5, 0.1, "div"))
    exact = np.arange(0, 900001) // 10000
    bad = np.nonzero(mul != div)[0]
    assert 0 < len(bad) < 60 and np.all(np.abs(mul - div) <= 1)
    # every disagreement (with each other or with exact integer division) sits on a voxel boundary
    for arr in (mul, div):
        off = np.nonzero(arr != exact)[0]
        assert np.all((off % 10000 == 0) | ((off + 1) % 10000 == 0))
