"""GPU parity tests of the whole denoiser: spvd_b200.SPVUnet (CUDA kernels through the C ABI) against
(a) the golden vectors produced by the reference's own Python (tests/golden, scale_mode='div') and
(b) the CPU oracle run here on the same seeded inputs (scale_mode='mul', the CUDA flavour)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    import spvd_b200
    return spvd_b200


def make_model(sp, preset, seed, mode, scale_mode, impl):
    from oracle import model as om
    sd = om.make_state_dict(preset, seed=seed)
    m = sp.build_model(preset, downsample_mode=mode, scale_mode=scale_mode)
    m.load_state_dict(sd, strict=True)
    return m.cuda().set_conv_impl(impl), sd


@pytest.mark.parametrize("impl,tol", [("simt", 2e-5), ("umma", 1e-3)])
@pytest.mark.parametrize("name", ["tiny_eval_spconv", "tiny_eval_floor", "spvd_eval_spconv"])
def test_eval_forward_vs_reference_golden(sp, golden, name, impl, tol):
    g = golden(name)
    preset, mode, _, seed, _ = (str(v) for v in g["meta"])
    m, _ = make_model(sp, preset, int(seed), mode, "div", sp.ops.CONV_SIMT if impl == "simt" else sp.ops.CONV_AUTO)
    m.eval()
    x = sp.SparseTensor(torch.from_numpy(g["feats"]).cuda(), torch.from_numpy(g["coords"]).cuda())
    with torch.no_grad():
        y = m((x, torch.from_numpy(g["t"]).cuda()))
    assert y.shape == g["out"].shape
    assert rel_err(y.cpu().numpy(), g["out"]) < tol


@pytest.mark.parametrize("impl,tol", [("simt", 1e-4), ("umma", 3e-3)])
@pytest.mark.parametrize("name", ["tiny_train_floor", "spvd_train_floor"])
def test_train_step_vs_reference_golden(sp, golden, name, impl, tol):
    """fwd + bwd in training mode (batch-statistics BatchNorm).  TF32 operands in ~40 chained convs and
    their gradients: the loss must agree to 1e-3, per-parameter gradient norms to 3e-3 (measured ~1e-3)."""
    g = golden(name)
    preset, mode, _, seed, _ = (str(v) for v in g["meta"])
    m, _ = make_model(sp, preset, int(seed), mode, "div", sp.ops.CONV_SIMT if impl == "simt" else sp.ops.CONV_AUTO)
    m.train()
    x = sp.SparseTensor(torch.from_numpy(g["feats"]).cuda(), torch.from_numpy(g["coords"]).cuda())
    y = m((x, torch.from_numpy(g["t"]).cuda()))
    assert rel_err(y.detach().cpu().numpy(), g["out"]) < tol
    loss = torch.nn.functional.mse_loss(y, torch.from_numpy(g["target"]).cuda())
    assert abs(loss.item() - float(g["loss"])) < max(tol, 1e-5) * abs(float(g["loss"]))
    loss.backward()
    params = dict(m.named_parameters())
    total = float(np.sqrt((g["grad_norms"] ** 2).sum()))
    for k, n in zip(g["grad_names"], g["grad_norms"]):
        got = float(params[str(k)].grad.double().norm())
        assert abs(got - n) <= 3 * tol * n + 1e-5 * total, (k, got, n)
    assert rel_err(params["out_conv.2.weight"].grad.cpu().numpy(), g["grad_out_w"]) < 3 * tol
    assert rel_err(params["stem_conv.voxel_conv.0.kernel"].grad.cpu().numpy(), g["grad_stem_k"]) < 3 * tol


@pytest.mark.parametrize("mode", ["spconv", "floor"])
def test_config1_b4x2048_vs_oracle(sp, mode):
    """BASELINE.json configs[0]: SPVD, single forward, B=4 x 2048 synthetic points, t=999, eval."""
    from oracle import model as om
    g = torch.Generator().manual_seed(0)
    pts = torch.randn(4, 2048, 3, generator=g).clamp(-3, 3)
    coords, feats, _ = om.torch2sparse(pts)
    t = torch.full((4,), 999, dtype=torch.long)
    m, sd = make_model(sp, "SPVD", 0, mode, "mul", sp.ops.CONV_AUTO)
    m.eval()
    with torch.no_grad():
        want = om.forward(sd, "SPVD", coords, feats, t, downsample_mode=mode, scale_mode="mul")
        want_tf32 = om.forward(sd, "SPVD", coords, feats, t, downsample_mode=mode, scale_mode="mul", emulate_tf32=True)
        y_exec = m((sp.SparseTensor(feats.cuda(), coords.cuda()), t.cuda())).cpu().numpy()   # native executor
        m.use_executor = False
        y = m((sp.SparseTensor(feats.cuda(), coords.cuda()), t.cuda())).cpu().numpy()        # op-by-op python path
        m.set_conv_impl(sp.ops.CONV_SIMT)
        y32 = m((sp.SparseTensor(feats.cuda(), coords.cuda()), t.cuda())).cpu().numpy()
    assert rel_err(y32, want.numpy()) < 2e-5                 # fp32 kernels vs fp32 oracle
    assert rel_err(y, want_tf32.numpy()) < 5e-4              # tensor-core path vs TF32-emulating oracle (round 2: the nn.Linear
                                                             # layers also run as TF32 k=1 convs; the oracle emulates the convs only)
    assert rel_err(y, want.numpy()) < 1e-3                   # and within the stated tolerance of plain fp32
    assert rel_err(y_exec, want.numpy()) < 1e-3              # executor: nn.Linear layers also run as TF32 k=1 convs


def test_native_executor_matches_python_path_and_oracle(sp):
    """The one-call C executor (spvd_program_run) against the op-by-op Python path on the same weights and
    geometry, and against the TF32-emulating oracle."""
    from oracle import model as om
    g = torch.Generator().manual_seed(5)
    pts = torch.randn(3, 1024, 3, generator=g).clamp(-3, 3)
    coords, feats, _ = om.torch2sparse(pts)
    t = torch.tensor([999, 500, 3])
    for preset in ("SPVD_TINY", "SPVD"):
        m, sd = make_model(sp, preset, 1, "spconv", "mul", sp.ops.CONV_AUTO)
        m.eval()
        x = sp.SparseTensor(feats.cuda(), coords.cuda())
        with torch.no_grad():
            assert m.use_executor
            y_exec = m((x, t.cuda())).cpu().numpy()
            assert len(m._programs) == 1
            m.use_executor = False
            y_py = m((x, t.cuda())).cpu().numpy()
            want = om.forward(sd, preset, coords, feats, t, emulate_tf32=True).numpy()
            want32 = om.forward(sd, preset, coords, feats, t).numpy()
        assert rel_err(y_exec, y_py) < 1e-3        # Linear layers run as TF32 k=1 convs in the executor
        assert rel_err(y_exec, want32) < 1e-3
        assert rel_err(y_py, want) < 5e-4          # (nn.Linear layers are TF32 k=1 convs on this path too since round 2)


def test_split_planning_matches_one_shot_planning(sp):
    """Planning the coarse levels on a side stream under the first ops (plan_split_level) must give the same geometry
    and the same output as planning everything up front."""
    from oracle import model as om
    g = torch.Generator().manual_seed(11)
    pts = torch.randn(4, 2048, 3, generator=g).clamp(-3, 3)
    coords, feats, _ = om.torch2sparse(pts)
    t = torch.tensor([999, 640, 77, 0]).cuda()
    m, _ = make_model(sp, "SPVD", 3, "spconv", "mul", sp.ops.CONV_AUTO)
    m.eval()
    x = sp.SparseTensor(feats.cuda(), coords.cuda())
    outs, geos = {}, {}
    with torch.no_grad():
        for split in (0, 1, 2, 3, (1, 2), (1, 3, 4)):
            m.plan_split_level = split
            for _ in range(2):                      # twice: the second call reuses the other workspace of the pool
                outs[split] = m((x, t)).cpu()
            geo = m.plan_geometry(coords.cuda().float(), 4, 0, split)
            first = split[0] if isinstance(split, tuple) else split
            assert (geo._pending is not None) == (first > 0) and geo.levels_ready == (first or 5)
            geo.finish()
            geos[split] = {s: (lv.n, lv.coords.cpu(), lv.kmap3.map_t[:, :lv.n].cpu(),
                               None if lv.kmap_down is None else lv.kmap_down.map_t[:, :lv.kmap_down.n_out].cpu(),
                               getattr(lv, "dpairs", (0,) * 5)[3:])
                           for s, lv in geo.levels.items()}
    for split in (1, 2, 3, (1, 2), (1, 3, 4)):
        assert rel_err(outs[split].numpy(), outs[0].numpy()) < 5e-4      # run-to-run (atomics order + TF32 flips) is ~1e-4
        for s, ref in geos[0].items():
            got = geos[split][s]
            assert got[0] == ref[0] and torch.equal(got[1], ref[1]) and torch.equal(got[2], ref[2]) and got[4] == ref[4]
            assert (ref[3] is None and got[3] is None) or torch.equal(got[3], ref[3])


@pytest.mark.parametrize("preset", ["SPVD", "COND_SPVD_L"])
def test_time_embedding_kernel_vs_torch(sp, preset):
    """spvd_time_embed_fwd (one launch) against the module-by-module torch evaluation of emb_mlp / cond_emb."""
    import math
    import torch.nn.functional as F
    from spvd_b200.runtime import Program
    m, _ = make_model(sp, preset, 2, "spconv", "mul", sp.ops.CONV_AUTO)
    m.eval()
    B = 7
    t = torch.tensor([999, 998, 500, 37, 2, 1, 0]).cuda()
    cls = torch.tensor([0, 54, 3, 3, 17, 1, 9]).cuda() if hasattr(m, "cond_emb") else None
    prog = Program(m, 64 * B, B)
    te = prog.recs[0]
    out = torch.empty(B, te["c1"], device="cuda")
    from spvd_b200._C import lib, check
    check(lib.spvd_time_embed_fwd(t.data_ptr(), cls.data_ptr() if cls is not None else None, te["w"].data_ptr(),
                                  te["p0"].data_ptr() if te["p0"] is not None else None, B, te["c0"], te["c1"], 1,
                                  out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    with torch.no_grad():
        exponent = -math.log(10000) * torch.linspace(0, 1, m.n_temb // 2, device="cuda")
        e = t[:, None].float() * exponent.exp()[None, :]
        e = torch.cat([e.sin(), e.cos()], dim=-1)
        emb = m.emb_mlp[1](m.emb_mlp[0](e))
        if cls is not None:
            emb = emb + m.cond_emb(cls)
        want = F.silu(emb)
    assert rel_err(out.cpu().numpy(), want.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("preset,B,N,train", [("COND_TINY", 3, 512, False), ("COND_TINY", 4, 256, True),
                                                 ("SPVD_L", 2, 512, False), ("COND_SPVD_L", 2, 512, False),
                                                 ("SUPERRES", 2, 1024, False), ("SPVD", 1, 8192, False)])
def test_other_configs_vs_oracle(sp, preset, B, N, train):
    """BASELINE.json configs[2..4] as parity cases: SPVD-L, the conditional model (class embedding, 55 ShapeNet
    categories), the 4-channel super-resolution topology, and a high-resolution cloud (8192 points)."""
    from oracle import model as om
    g = torch.Generator().manual_seed(len(preset) + N)
    pts = torch.randn(B, N, 3, generator=g).clamp(-3, 3)
    coords, feats, _ = om.torch2sparse(pts)
    cfg = om.get_config(preset)
    if cfg["point_channels"] == 4:                      # xyz + a 0/1 mask channel (utils/completion_schedulers.py:43-59)
        feats = torch.cat([feats, (torch.rand(feats.shape[0], 1, generator=g) > 0.75).float()], 1)
    t = torch.randint(0, 1000, (B,), generator=g)
    c = torch.randint(0, cfg["n_classes"], (B,), generator=g) if cfg.get("n_classes") else None
    sd = om.make_state_dict(preset, seed=2)
    m = sp.build_model(preset)
    assert sum(p.numel() for p in m.parameters()) == om.num_params(preset)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train(train)
    inp = (sp.SparseTensor(feats.cuda(), coords.cuda()), t.cuda()) + ((c.cuda(),) if c is not None else ())
    if train:
        params = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
        want = om.forward(params, preset, coords, feats, t, c, training=True)
        want.square().mean().backward()
        y = m(inp)
        y.square().mean().backward()
        assert rel_err(y.detach().cpu().numpy(), want.detach().numpy()) < 2e-3
        gk = "cond_emb.weight"
        assert rel_err(dict(m.named_parameters())[gk].grad.cpu().numpy(), params[gk].grad.numpy()) < 1e-2
    else:
        with torch.no_grad():
            want = om.forward(sd, preset, coords, feats, t, c)
            y = m(inp)
        assert y.shape == (coords.shape[0], cfg["point_channels"])
        assert rel_err(y.cpu().numpy(), want.numpy()) < 1e-3
