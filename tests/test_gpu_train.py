"""GPU tests of the training-mode glue (csrc/train.cu, row TR of the scope table): every autograd Function of
spvd_b200.functional against the plain torch fp32 implementation of the same op (forward, input and parameter gradients),
and the whole training step: no ATen / cuBLAS / cuDNN compute kernel may run between the sparse convolutions.
Reference semantics: models/ddpm_unet_attn.py:24-65, models/modules/graph.py:19-41, models/modules/transformer.py:40-67,
102-125, models/utils.py:15-33, train_generation.py:98-112."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _cmp(a, b, tol, what=""):
    e = rel_err(a.detach().cpu().numpy(), b.detach().cpu().numpy())
    assert e < tol, (what, e)


@pytest.mark.parametrize("rows,c,act", [(5000, 64, True), (777, 3, True), (32, 128, True), (20000, 448, False), (300, 4, False)])
def test_batch_norm_train_fwd_bwd(rows, c, act):
    from spvd_b200 import functional as SF
    torch.manual_seed(0)
    x = (torch.randn(rows, c, device="cuda") * 1.7 + 0.3).requires_grad_(True)
    bn = torch.nn.BatchNorm1d(c).cuda().train()
    bn.weight.data.uniform_(0.5, 1.5)
    bn.bias.data.normal_()
    ref = torch.nn.BatchNorm1d(c).cuda().train()
    ref.load_state_dict(bn.state_dict())
    gy = torch.randn(rows, c, device="cuda")
    y = SF.batch_norm_act(x, bn, act)
    y.backward(gy)
    x2 = x.detach().clone().requires_grad_(True)
    y2 = ref(x2)
    y2 = F.silu(y2) if act else y2
    y2.backward(gy)
    _cmp(y, y2, 2e-6, "y")
    _cmp(x.grad, x2.grad, 2e-5, "gx")
    _cmp(bn.weight.grad, ref.weight.grad, 2e-5, "ggamma")
    _cmp(bn.bias.grad, ref.bias.grad, 2e-5, "gbeta")
    _cmp(bn.running_mean, ref.running_mean, 1e-6, "running_mean")
    _cmp(bn.running_var, ref.running_var, 1e-6, "running_var")


@pytest.mark.parametrize("rows,cin,cout,bias,res", [(4000, 64, 128, True, True), (3000, 3, 32, True, False), (3000, 32, 3, False, False),
                                                     (32, 128, 512, True, False), (700, 192, 576, False, False), (2500, 96, 64, True, True)])
def test_linear_fwd_bwd(rows, cin, cout, bias, res):
    """nn.Linear through the conv kernels: tcgen05 TF32 where cin, cout are multiples of 32 (tolerance of TF32 operands),
    fp32 SIMT otherwise."""
    from spvd_b200 import functional as SF
    torch.manual_seed(1)
    lin = torch.nn.Linear(cin, cout, bias=bias).cuda()
    x = torch.randn(rows, cin, device="cuda", requires_grad=True)
    r = torch.randn(rows, cout, device="cuda", requires_grad=True) if res else None
    gy = torch.randn(rows, cout, device="cuda")
    y = SF.linear(x, lin.weight, lin.bias, SF.LinearCache(), residual=r)
    y.backward(gy)
    got = (y, x.grad, lin.weight.grad, lin.bias.grad if bias else None, r.grad if res else None)
    lin.zero_grad()
    x2 = x.detach().clone().requires_grad_(True)
    r2 = r.detach().clone().requires_grad_(True) if res else None
    y2 = lin(x2) + (r2 if res else 0)
    y2.backward(gy)
    tol = 1e-3 if (cin % 32 == 0 and cout % 32 == 0) else 1e-5
    for a, b, n in zip(got, (y2, x2.grad, lin.weight.grad, lin.bias.grad if bias else None, r2.grad if res else None),
                       ("y", "gx", "gw", "gb", "gres")):
        if a is not None:
            _cmp(a, b, tol if n != "gres" else 1e-7, n)


def _ragged_offsets(counts):
    off = torch.zeros(len(counts) + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(torch.tensor(counts), 0)
    return off.cuda()


def test_film_fwd_bwd():
    from spvd_b200 import functional as SF
    torch.manual_seed(2)
    counts = [300, 1, 0, 777, 64, 2000]
    off = _ragged_offsets(counts)
    n, c, B = sum(counts), 96, len(counts)
    b = torch.repeat_interleave(torch.arange(B), torch.tensor(counts)).cuda()
    coords = torch.zeros(n, 4, dtype=torch.int32, device="cuda")
    coords[:, 0] = b.int()
    x = torch.randn(n, c, device="cuda", requires_grad=True)
    tt = torch.randn(B, 2 * c, device="cuda", requires_grad=True)
    gy = torch.randn(n, c, device="cuda")
    y = SF.film(x, tt, coords, off, max(counts))
    y.backward(gy)
    x2, tt2 = x.detach().clone().requires_grad_(True), tt.detach().clone().requires_grad_(True)
    sc, sh = torch.chunk(tt2[b], 2, dim=-1)
    y2 = (1 + sc) * x2 + sh
    y2.backward(gy)
    _cmp(y, y2, 1e-6, "y")
    _cmp(x.grad, x2.grad, 1e-6, "gx")
    _cmp(tt.grad, tt2.grad, 2e-5, "gtt")


def test_layernorm_fwd_bwd():
    from spvd_b200 import functional as SF
    torch.manual_seed(3)
    for rows, c in ((3000, 192), (77, 256), (1, 64)):
        ln = torch.nn.LayerNorm(c).cuda()
        ln.weight.data.uniform_(0.5, 1.5)
        ln.bias.data.normal_()
        x = (torch.randn(rows, c, device="cuda") * 2 + 1).requires_grad_(True)
        gy = torch.randn(rows, c, device="cuda")
        y = SF.layer_norm(x, ln)
        y.backward(gy)
        got = (y, x.grad, ln.weight.grad.clone(), ln.bias.grad.clone())
        ln.zero_grad()
        x2 = x.detach().clone().requires_grad_(True)
        y2 = ln(x2)
        y2.backward(gy)
        for a, b_, n in zip(got, (y2, x2.grad, ln.weight.grad, ln.bias.grad), ("y", "gx", "gg", "gb")):
            _cmp(a, b_, 2e-5, n)


@pytest.mark.parametrize("c,heads", [(192, 8), (256, 8), (64, 4)])
def test_attention_fwd_bwd(c, heads):
    """per-cloud attention over ragged row ranges against torch softmax attention run cloud by cloud (the reference's
    exp * mask / sum is the same function, models/utils.py:22-33)"""
    from spvd_b200 import functional as SF
    torch.manual_seed(4)
    counts = [130, 257, 1, 64, 33]
    off = _ragged_offsets(counts)
    n, d = sum(counts), c // heads
    qkv = torch.randn(n, 3 * c, device="cuda", requires_grad=True)
    go = torch.randn(n, c, device="cuda")
    o = SF.attention(qkv, off, len(counts), max(counts), heads)
    o.backward(go)
    q2 = qkv.detach().clone().requires_grad_(True)
    outs, r = [], 0
    for L in counts:
        blk = q2[r:r + L].view(L, 3, heads, d).permute(1, 2, 0, 3)          # [3, H, L, D]
        a = torch.softmax(blk[0] @ blk[1].transpose(-1, -2) / math.sqrt(d), dim=-1) @ blk[2]
        outs.append(a.permute(1, 0, 2).reshape(L, c))
        r += L
    o2 = torch.cat(outs, 0)
    o2.backward(go)
    _cmp(o, o2, 2e-5, "out")
    _cmp(qkv.grad, q2.grad, 5e-5, "gqkv")


def test_small_ops():
    from spvd_b200 import functional as SF, ops
    torch.manual_seed(5)
    a = torch.randn(1000, 48, device="cuda", requires_grad=True)
    b = torch.randn(1000, 3, device="cuda", requires_grad=True)
    g = torch.randn(1000, 51, device="cuda")
    y = SF.cat(a, b)
    y.backward(g)
    assert torch.equal(y, torch.cat([a, b], 1)) and torch.equal(a.grad, g[:, :48]) and torch.equal(b.grad, g[:, 48:])
    x = torch.randn(64, 128, device="cuda", requires_grad=True)
    gy = torch.randn(64, 128, device="cuda")
    s = SF.silu(x)
    s.backward(gy)
    x2 = x.detach().clone().requires_grad_(True)
    F.silu(x2).backward(gy)
    _cmp(s, F.silu(x2), 1e-6)
    _cmp(x.grad, x2.grad, 1e-5)
    # class embedding added to the time embedding
    table = torch.randn(55, 128, device="cuda", requires_grad=True)
    base = torch.randn(9, 128, device="cuda", requires_grad=True)
    idx = torch.tensor([0, 54, 3, 3, 3, 17, 0, 9, 54]).cuda()
    ge = torch.randn(9, 128, device="cuda")
    e = SF.embedding_add(base, table, idx)
    e.backward(ge)
    t2, b2 = table.detach().clone().requires_grad_(True), base.detach().clone().requires_grad_(True)
    (b2 + F.embedding(idx, t2)).backward(ge)
    _cmp(e, b2 + F.embedding(idx, t2), 1e-7)
    _cmp(table.grad, t2.grad, 1e-6)
    _cmp(base.grad, b2.grad, 1e-7)
    # sinusoidal timestep embedding (models/utils.py:15-19: linspace(0, 1, d/2) includes 1.0)
    t = torch.tensor([0, 1, 37, 500, 999]).cuda()
    for dim in (32, 64):
        exponent = -math.log(10000) * torch.linspace(0, 1, dim // 2, device="cuda")
        ang = t[:, None].float() * exponent.exp()[None, :]
        want = torch.cat([ang.sin(), ang.cos()], dim=-1)
        assert float((ops.timestep_embedding(t, dim) - want).abs().max()) < 2e-4        # sin/cos of arguments up to ~1e3
    # MSE loss + gradient
    p = torch.randn(65536, 3, device="cuda", requires_grad=True)
    tg = torch.randn(65536, 3, device="cuda")
    loss = SF.mse_loss(p, tg)
    loss.backward()
    p2 = p.detach().clone().requires_grad_(True)
    l2 = F.mse_loss(p2, tg)
    l2.backward()
    assert abs(float(loss) - float(l2)) < 1e-6 * float(l2)
    _cmp(p.grad, p2.grad, 1e-6)


def test_training_step_runs_no_library_kernels():
    """BASELINE north_star: 'the full SPVD denoiser forward/backward runs entirely on the new sm_100a kernels'.  Profile one
    training step (forward + MSE + backward) and check the names of the CUDA kernels: nothing from cuBLAS / cuDNN / CUTLASS /
    flash attention, no ATen batch-norm / layer-norm / index kernels (ATen elementwise adds of autograd's gradient
    accumulation are plumbing and allowed)."""
    import spvd_b200 as sp
    from spvd_b200 import functional as SF
    from oracle import model as om
    from torch.profiler import ProfilerActivity, profile
    m = sp.build_model("SPVD_TINY")
    m.load_state_dict(om.make_state_dict("SPVD_TINY", seed=0))
    m = m.cuda().train()
    g = torch.Generator().manual_seed(0)
    pts = torch.randn(4, 512, 3, generator=g).clamp(-3, 3)
    coords, feats, _ = om.torch2sparse(pts)
    x, t = sp.SparseTensor(feats.cuda(), coords.cuda()), torch.tensor([999, 500, 20, 3]).cuda()
    target = torch.randn(feats.shape[0], 3, device="cuda")
    SF.mse_loss(m((x, t)), target).backward()                  # warm-up (library build, caches)
    m.zero_grad()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        SF.mse_loss(m((x, t)), target).backward()
        torch.cuda.synchronize()
    names = [e.name for e in prof.events() if getattr(e, "device_type", None) is not None and "cuda" in str(e.device_type).lower()]
    names = [n for n in names if n and not n.lower().startswith("mem")]
    if not names:
        pytest.skip("the profiler recorded no CUDA kernels on this box")
    bad = [n for n in names if any(k in n.lower() for k in ("cublas", "cudnn", "cutlass", "gemm", "fmha", "flash", "batch_norm",
                                                             "layer_norm", "index", "softmax", "scatter", "gather"))
           and "spvd" not in n]
    assert not bad, sorted(set(bad))
    ours = [n for n in names if "spvd" in n or n.startswith("k_") or "::k_" in n]
    assert len(ours) > 0.8 * len(names), (len(ours), len(names), sorted(set(n for n in names if n not in ours))[:20])


def test_in_place_parameter_gradients_match_autograd_accumulation():
    """With a GradientAllReduce attached the backward kernels write every parameter gradient straight into its slot of the
    flat buffer (functional._claim): same gradients as autograd's own accumulation, every slot written exactly once, a
    second backward without zero_grad() accumulates, and the per-parameter ATen adds are gone."""
    import spvd_b200 as sp
    from spvd_b200 import functional as SF
    from spvd_b200.parallel import GradientAllReduce
    from oracle import model as om
    from torch.profiler import ProfilerActivity, profile
    m = sp.build_model("SPVD_TINY")
    m.load_state_dict(om.make_state_dict("SPVD_TINY", seed=0))
    m = m.cuda().train()
    g = torch.Generator().manual_seed(0)
    pts = torch.randn(4, 512, 3, generator=g).clamp(-3, 3)
    coords, feats, _ = om.torch2sparse(pts)
    x, t = sp.SparseTensor(feats.cuda(), coords.cuda()), torch.tensor([999, 500, 20, 3]).cuda()
    target = torch.randn(feats.shape[0], 3, device="cuda")
    bn_state = {k: v.clone() for k, v in m.state_dict().items()}

    def run(direct, twice=False):
        m.load_state_dict(bn_state)                                # same running statistics for both runs
        red.direct = direct
        red.zero_grad()
        SF.mse_loss(m((x, t)), target).backward()
        if twice:
            SF.mse_loss(m((x, t)), target).backward()
        torch.cuda.synchronize()
        return red.flat.clone(), len(red._written)

    red = GradientAllReduce(m.parameters())
    run(True)                                                       # warm-up
    ref, n_ref = run(False)
    ref2, _ = run(False)
    got, n_got = run(True)
    n_params = sum(1 for p in m.parameters() if p.requires_grad)
    assert n_ref == 0 and n_got > 0.9 * n_params, (n_ref, n_got, n_params)
    assert float(ref.abs().max()) > 0
    # scatter / weight-gradient kernels accumulate with atomics and TF32 rounding amplifies a last-bit difference to 2^-11:
    # two runs of the SAME path differ by `noise`; a missing or doubled slot would be an O(1) error
    noise = rel_err(ref2.cpu().numpy(), ref.cpu().numpy())
    tol = max(4 * noise, 1e-5)
    assert tol < 1e-2, noise
    assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < tol, noise
    floor = 1e-4 * float(ref.abs().max())                           # (a conv bias in front of a BatchNorm has a zero gradient:
    for p in m.parameters():                                        # rounding noise only) per slot: nothing missing or doubled
        o = red._offset_of(p) // 4
        a, b = got[o:o + p.numel()], ref[o:o + p.numel()]
        assert float((a - b).norm()) <= 0.05 * float(b.norm()) + floor, (tuple(p.shape), float(a.norm()), float(b.norm()))
    twice, _ = run(True, twice=True)
    assert rel_err(twice.cpu().numpy(), 2 * ref.cpu().numpy()) < tol, noise
    red.direct = True
    red.zero_grad()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        SF.mse_loss(m((x, t)), target).backward()
        torch.cuda.synchronize()
    names = [e.name for e in prof.events() if getattr(e, "device_type", None) is not None and "cuda" in str(e.device_type).lower()]
    adds = [n for n in names if "vectorized_elementwise_kernel" in n and "add" in n.lower()]
    if names:
        assert len(adds) < 0.25 * n_params, (len(adds), n_params)


@pytest.mark.parametrize("wd,clip,pending", [(0.0, 10.0, 1.0), (0.01, 0.05, 1.0), (0.0, 0.0, 1.0), (0.0, 0.05, 0.125), (0.0, 0.0, 0.5)])
def test_flat_adam_matches_torch_adam_and_clip(wd, clip, pending):
    """optim.FlatAdam (spvd_sumsq + spvd_adam_step over the flat buffers) against torch.optim.Adam after
    torch.nn.utils.clip_grad_norm_ (train_generation.py:106-112), several steps, odd parameter sizes.  ``pending`` is the
    factor still owed to the summed gradients (1 / world after the all-reduce): it must act before the clip."""
    from spvd_b200.optim import FlatAdam
    from spvd_b200.parallel import GradientAllReduce
    torch.manual_seed(3)
    make = lambda: torch.nn.Sequential(torch.nn.Linear(7, 33), torch.nn.SiLU(), torch.nn.Linear(33, 130), torch.nn.SiLU(),
                                       torch.nn.Linear(130, 3)).cuda()
    a, b = make(), make()
    b.load_state_dict(a.state_dict())
    red = GradientAllReduce(a.parameters())
    opt_a = FlatAdam(red, lr=3e-3, weight_decay=wd)
    opt_b = torch.optim.Adam(b.parameters(), lr=3e-3, weight_decay=wd)
    sched = torch.optim.lr_scheduler.StepLR(opt_a, 2, 0.5)
    sched_b = torch.optim.lr_scheduler.StepLR(opt_b, 2, 0.5)
    v0 = [p._version for p in a.parameters()]
    for it in range(6):
        x, y = torch.randn(64, 7, device="cuda") * (1 + it), torch.randn(64, 3, device="cuda")
        opt_a.zero_grad()
        F.mse_loss(a(x), y).backward()
        opt_a.step(max_norm=clip, pending_scale=pending)
        sched.step()
        opt_b.zero_grad()
        F.mse_loss(b(x), y).backward()
        for pb in b.parameters():
            pb.grad.mul_(pending)
        nb = torch.nn.utils.clip_grad_norm_(b.parameters(), clip) if clip > 0 else None
        opt_b.step()
        sched_b.step()
        if nb is not None:
            assert abs(float(opt_a.grad_norm) - float(nb)) <= 1e-5 * float(nb)
        for pa, pb in zip(a.parameters(), b.parameters()):
            _cmp(pa, pb, 2e-6, f"step {it}")
    assert all(p._version > v for p, v in zip(a.parameters(), v0))      # version-keyed caches see the in-place update
    assert all(p.data_ptr() >= opt_a.flat_p.data_ptr() for p in a.parameters())


def test_prefetched_geometry_gives_the_same_training_step():
    """model.prefetch_geometry plans batch j + 1 on a side stream while step j is issued (no host wait until its first use);
    the step computed on a prefetched geometry equals the step that plans its own, buffer sets alternate safely over
    several steps, and a geometry whose buffer set was re-planned is refused in backward."""
    import spvd_b200 as sp
    from spvd_b200 import functional as SF
    from oracle import model as om
    m = sp.build_model("SPVD_TINY")
    m.load_state_dict(om.make_state_dict("SPVD_TINY", seed=0))
    m = m.cuda().train()
    state = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    batches = []
    for j in range(4):
        pts = torch.randn(4, 384 + 64 * j, 3, generator=g).clamp(-3, 3)[:, :384]
        coords, feats, _ = om.torch2sparse(pts)
        batches.append((sp.SparseTensor(feats.cuda(), coords.cuda()), torch.tensor([999, 500, 20, 3]).cuda(),
                        torch.randn(feats.shape[0], 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(j))))

    def grads():
        return torch.cat([p.grad.flatten() for p in m.parameters()]).clone()

    ref = []
    for x, t, tg in batches:
        m.load_state_dict(state)
        m.zero_grad()
        loss = SF.mse_loss(m((x, t)), tg)
        loss.backward()
        ref.append((float(loss), grads()))
    # a FRESH model: its very first geometry builds are the prefetches (the buffer-set cursor must alternate from the
    # first call on -- it once handed the second set out twice in a row at start-up)
    m_ref = m
    m = sp.build_model("SPVD_TINY")
    m.load_state_dict(state)
    m = m.cuda().train()
    got = []
    geo = m.prefetch_geometry(batches[0][0], 4)
    for j, (x, t, tg) in enumerate(batches):
        geo_next = m.prefetch_geometry(batches[j + 1][0], 4) if j + 1 < len(batches) else None
        m.load_state_dict(state)
        m.zero_grad()
        loss = SF.mse_loss(m((x, t), geometry=geo), tg)
        loss.backward()
        got.append((float(loss), grads()))
        geo = geo_next
    torch.cuda.synchronize()
    sets = [ws for pool in m._plan_ws.values() for ws in pool[0]]
    assert len(sets) == 2 and sorted(ws.generation for ws in sets) == [2, 2], [ws.generation for ws in sets]
    del m_ref
    for (l0, g0), (l1, g1) in zip(ref, got):
        assert abs(l0 - l1) <= 1e-5 * abs(l0), (l0, l1)
        assert rel_err(g1.cpu().numpy(), g0.cpu().numpy()) < 2e-3       # run-to-run noise of the atomics (see the in-place test)
    # a stale geometry (its buffer set has been planned again twice since) must not be used silently
    x, t, tg = batches[0]
    geo_old = m.prefetch_geometry(x, 4)
    loss = SF.mse_loss(m((x, t), geometry=geo_old), tg)
    m.prefetch_geometry(batches[1][0], 4).finish()
    m.prefetch_geometry(batches[2][0], 4).finish()
    with pytest.raises(RuntimeError):
        loss.backward()
