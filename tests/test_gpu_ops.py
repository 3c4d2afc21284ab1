"""GPU parity tests, op by op: the CUDA path (through the C ABI, via spvd_b200.ops) against the CPU
oracle on the same seeded inputs.  Integer results (coordinates, indices, kernel maps) must be
bit-exact; fp32 features within 1e-5 relative (plain fp32 kernels) or 1e-3 (TF32 tensor-core conv,
the tolerance BASELINE.json states), and within 2e-5 of the oracle that emulates TF32 rounding."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    import spvd_b200
    return spvd_b200


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def random_coords(n, nb, extent, seed, lo=0):
    rng = np.random.default_rng(seed)
    c = np.empty((n, 4), dtype=np.int32)
    c[:, 0] = rng.integers(0, nb, n)
    c[:, 1:] = rng.integers(lo, lo + extent, (n, 3))
    return c


def cloud_points(B, N, seed):
    """integer point coordinates (b,x,y,z) at 1e-5 resolution like utils/schedulers.py:249-257"""
    from oracle import model as om
    g = torch.Generator().manual_seed(seed)
    pts = torch.randn(B, N, 3, generator=g).clamp(-3, 3)
    coords, feats, _ = om.torch2sparse(pts)
    return coords, feats


# ------------------------------------------------------------------------------------------ V1
@pytest.mark.parametrize("mode", ["mul", "div"])
def test_point_voxel_coords_rounding(sp, mode):
    from oracle import ops as oo
    c = np.concatenate([np.arange(0, 900001, 10000), np.arange(0, 900001, 10000) - 1,
                        np.random.default_rng(0).integers(0, 900000, 50000)]).clip(0).astype(np.float32)
    n = len(c) // 3 * 3
    pc = np.zeros((n // 3, 4), dtype=np.float32)
    pc[:, 1:] = c[:n].reshape(-1, 3)
    pc[:, 0] = np.arange(n // 3) % 7
    f, i = sp.ops.point_voxel_coords(dev(pc), 1e-5, 0.1, mode)
    want = oo.scale_point_coords(pc[:, 1:], 1e-5, 0.1, mode)
    assert np.array_equal(f.cpu().numpy()[:, 1:], want)            # bit-exact fp32
    assert np.array_equal(i.cpu().numpy()[:, 1:], np.floor(want).astype(np.int32))
    assert np.array_equal(i.cpu().numpy()[:, 0], pc[:, 0].astype(np.int32))


# ------------------------------------------------------------------------------- unique / sort
@pytest.mark.parametrize("n,nb,extent,lo", [(1, 1, 4, 0), (37, 2, 3, -1), (4096, 3, 12, -5), (4097, 4, 40, 0),
                                            (100000, 32, 61, 0), (300000, 5, 700, -300)])
def test_unique_coords(sp, n, nb, extent, lo):
    from oracle import ops as oo
    c = random_coords(n, nb, extent, n, lo)
    out, cnt = sp.ops.unique_coords(dev(c))
    want, inv = oo.unique_voxels(c)
    k = int(cnt.item())
    assert k == want.shape[0]
    assert np.array_equal(out[:k].cpu().numpy(), want)


def test_unique_coords_device_count_and_range_flag(sp):
    from oracle import ops as oo
    c = random_coords(5000, 2, 9, 1)
    n_dev = torch.tensor([1234], dtype=torch.int32).cuda()
    status = torch.zeros(1, dtype=torch.int32).cuda()
    out, cnt = sp.ops.unique_coords(dev(c), n_dev, status)
    want, _ = oo.unique_voxels(c[:1234])
    assert int(cnt.item()) == want.shape[0] and np.array_equal(out[: want.shape[0]].cpu().numpy(), want)
    assert int(status.item()) == 0
    c[17, 2] = 40000                                       # does not fit the 16-bit packed key
    sp.ops.unique_coords(dev(c), None, status)
    assert int(status.item()) & sp.ops.STATUS_COORD_RANGE


# ---------------------------------------------------------------------------------------- hash
@pytest.mark.parametrize("ks", [1, 2, 3])
def test_hash_query(sp, ks):
    from oracle import ops as oo
    target, _ = oo.unique_voxels(random_coords(20000, 4, 25, 5, -3))
    query = random_coords(30000, 5, 29, 6, -5)
    tab = sp.ops.HashTable(2 * target.shape[0], "cuda").build(dev(target))
    got = tab.query(dev(query), ks).cpu().numpy() - 1
    assert np.array_equal(got, oo.hash_query(query, target, ks))
    # (x,y,z,b) column order, the way sphashquery passes coordinates (models/sparse_utils.py:44,51)
    tab2 = sp.ops.HashTable(2 * target.shape[0], "cuda").build(dev(target[:, [1, 2, 3, 0]]), order=sp.ops.ORDER_XYZB)
    got2 = tab2.query(dev(query[:, [1, 2, 3, 0]]), ks, order=sp.ops.ORDER_XYZB).cpu().numpy() - 1
    assert np.array_equal(got2, got)


def test_kmap_subm3(sp):
    from oracle import ops as oo
    vox, _ = oo.unique_voxels(random_coords(50000, 8, 30, 7))
    n = vox.shape[0]
    d = dev(vox)
    tab = sp.ops.HashTable(2 * n, "cuda").build(d)
    map_t, mask = sp.ops.kmap_subm3(d, None, tab)
    want = oo.kmap_subm(vox, 3)                              # [n,27]
    got = map_t.cpu().numpy()
    assert np.array_equal(got[:, :n].T, want)
    assert np.all(got[:, n:] == -1)
    m = mask.cpu().numpy().astype(np.uint32)
    ld = got.shape[1]
    pad = np.full((ld, 27), -1, dtype=np.int32)
    pad[:n] = want
    bits = ((pad.reshape(ld // 128, 128, 27) >= 0).any(axis=1) * (1 << np.arange(27, dtype=np.uint64))).sum(axis=1)
    assert np.array_equal(m.astype(np.uint64), bits.astype(np.uint64))


@pytest.mark.parametrize("mode", ["spconv", "floor"])
def test_downsample_and_strided_maps(sp, mode):
    from oracle import ops as oo
    fine, _ = oo.unique_voxels(random_coords(40000, 6, 31, 9))
    n = fine.shape[0]
    d = dev(fine)
    coarse, cnt = sp.ops.downsample_coords(d, None, mode)
    want_c, want_parent, want_koff = oo.downsample_coords(fine, mode)
    nc = int(cnt.item())
    assert nc == want_c.shape[0]
    assert np.array_equal(coarse[:nc].cpu().numpy(), want_c)
    if mode == "spconv":
        assert (want_parent < 0).any()                       # the boundary drop really happens in this case
    tab = sp.ops.HashTable(2 * n, "cuda").build(coarse, cnt)
    md, kd, mu, ku, parent, koff = sp.ops.kmap_down2(d, None, tab, sp.ops.round128(n), want_parent=True)
    assert np.array_equal(parent.cpu().numpy(), want_parent)
    assert np.array_equal(koff.cpu().numpy(), want_koff)
    want_md = oo.kmap_down(nc, want_parent, want_koff)
    got_md = md.cpu().numpy()
    assert np.array_equal(got_md[:, :nc].T, want_md) and np.all(got_md[:, nc:] == -1)
    want_mu = np.full((n, 8), -1, dtype=np.int32)
    rows = np.nonzero(want_parent >= 0)[0]
    want_mu[rows, want_koff[rows]] = want_parent[rows]
    got_mu = mu.cpu().numpy()
    assert np.array_equal(got_mu[:, :n].T, want_mu) and np.all(got_mu[:, n:] == -1)


@pytest.mark.parametrize("parent,mode", [(False, "spconv"), (True, "spconv"), (True, "floor")])
def test_segmented_unique_ragged_clouds(sp, parent, mode):
    """per-cloud shared-memory sort: ragged clouds incl. an empty one and one that fills the capacity exactly"""
    from oracle import ops as oo
    rng = np.random.default_rng(3)
    sizes = [700, 0, 2048, 1, 1500, 33]
    rows = []
    for b, n in enumerate(sizes):
        c = np.empty((n, 4), dtype=np.int32)
        c[:, 0] = b
        c[:, 1:] = rng.integers(-6, 14, (n, 3))
        rows.append(c)
    c = np.concatenate(rows)
    if parent:
        c, _ = oo.unique_voxels(c)                       # a real (sorted, unique) fine level
    out, cnt, bc, bo = sp.ops.segmented_unique(dev(c), len(sizes), 2048, parent, mode)
    want = oo.downsample_coords(c, mode)[0] if parent else oo.unique_voxels(c)[0]
    k = int(cnt.item())
    assert k == want.shape[0] and np.array_equal(out[:k].cpu().numpy(), want)
    assert bc.cpu().tolist() == np.bincount(want[:, 0], minlength=len(sizes)).tolist()
    assert bo.cpu().tolist() == [0] + np.cumsum(bc.cpu().numpy()).tolist()
    status = torch.zeros(1, dtype=torch.int32).cuda()
    sp.ops.segmented_unique(dev(c), len(sizes), 512, parent, mode, status=status)      # hint too small -> flagged
    if max(np.bincount(c[:, 0])) > 512:
        assert int(status.item()) & 4


def test_empty_level_is_reported(sp):
    """spconv rule with max coordinate 0 on an axis drops everything: the count
    comes back 0 and Geometry raises a clear error instead of crashing inside a kernel."""
    c = np.zeros((50, 4), dtype=np.int32)
    c[:, 1] = np.arange(50)
    out, cnt = sp.ops.downsample_coords(dev(c), None, "spconv")
    assert int(cnt.item()) == 0
    out, cnt = sp.ops.downsample_coords(dev(c), None, "floor")
    assert int(cnt.item()) == 25


# ------------------------------------------------------------------------------ geometry pyramid
@pytest.mark.parametrize("mode", ["spconv", "floor"])
def test_geometry_matches_oracle(sp, mode):
    from oracle import model as om
    coords, feats = cloud_points(4, 2048, 3)
    og = om.Geometry(coords, 1e-5, 0.1, mode, "mul")
    geo = sp.Geometry(coords.float().cuda(), 1e-5, 0.1, strides=(1, 2, 4, 8, 16), p2v_strides=(1, 4, 16),
                      devox_strides=(1, 4, 16), n_batch=4, downsample_mode=mode)
    assert np.array_equal(geo.fcoords.cpu().numpy()[:, 1:], og.fcoords)
    for s in (1, 2, 4, 8):
        og.downsample(s)
    for s in (1, 2, 4, 8, 16):
        lv = geo.level(s)
        assert lv.n == og.level(s).n
        assert np.array_equal(lv.coords.cpu().numpy(), og.level(s).coords)
        assert np.array_equal(lv.kmap3.map_t.cpu().numpy()[:, : lv.n].T, og.kmap3(s))
        assert lv.batch_counts_host == np.bincount(og.level(s).coords[:, 0], minlength=4).tolist()
    for s in (1, 4, 16):
        assert np.array_equal(geo.p2v[s].cpu().numpy(), og.point_to_voxel_idx(s))
        idx, w = og.devox_idx_weights(s)
        assert np.array_equal(geo.devox[s][0].cpu().numpy(), idx.numpy())
        assert np.allclose(geo.devox[s][1].cpu().numpy(), w.numpy(), atol=2e-7)


def test_one_call_planner_matches_op_by_op_geometry(sp):
    """spvd_plan_build (one C call, persistent buffers) against the Geometry assembled from individual ops."""
    coords, _ = cloud_points(4, 2048, 8)
    pc = coords.float().cuda()
    kw = dict(strides=(1, 2, 4, 8, 16), p2v_strides=(1, 4, 16), devox_strides=(1, 4, 16))
    for mode in ("spconv", "floor"):
        ref = sp.Geometry(pc, 1e-5, 0.1, n_batch=4, downsample_mode=mode, **kw)
        ws = sp.geometry.PlanWorkspace(pc.shape[0], 4, kw["strides"], kw["strides"], kw["p2v_strides"], kw["devox_strides"], pc.device)
        # buffers are reused; hint 2048 = per-cloud smem sort path; the planner is replayed from CUDA graphs (one capture per
        # mode / hint / level grouping) or launched kernel by kernel; level groups on a side stream
        side = torch.cuda.Stream()
        for graphs, hint, split in ((False, 0, 0), (True, 0, 0), (True, 0, 0), (False, 2048, (1, 3)), (True, 2048, (1, 3)),
                                    (True, 2048, (1, 3)), (True, 2048, 0)):
            ws.use_graphs = graphs
            geo = ws.build(pc, 1e-5, 0.1, "mul", mode, max_cloud_rows=hint, split_level=split, side_stream=side).finish()
            _check_same_geometry(geo, ref, kw)


def _check_same_geometry(geo, ref, kw):
    if True:
        assert torch.equal(geo.fcoords, ref.fcoords)
        for s in kw["strides"]:
            a, b = geo.level(s), ref.level(s)
            assert a.n == b.n and torch.equal(a.coords, b.coords)
            assert torch.equal(a.kmap3.map_t, b.kmap3.map_t) and torch.equal(a.kmap3.mask, b.kmap3.mask)
            assert a.batch_counts_host == b.batch_counts_host
            assert torch.equal(geo.batch_offsets(s), ref.batch_offsets(s))
            if s < 16:
                assert torch.equal(a.kmap_down.map_t, b.kmap_down.map_t) and torch.equal(a.kmap_up.map_t, b.kmap_up.map_t)
                assert torch.equal(a.kmap_down.mask, b.kmap_down.mask) and torch.equal(a.kmap_up.mask, b.kmap_up.mask)
        for s in (1, 4, 16):
            assert torch.equal(geo.p2v[s], ref.p2v[s])
            assert torch.equal(geo.devox[s][0], ref.devox[s][0]) and torch.equal(geo.devox[s][1], ref.devox[s][1])


# ---------------------------------------------------------------------------- point <-> voxel ops
@pytest.mark.parametrize("c", [3, 32, 192])
def test_scatter_mean_and_devoxelize(sp, c):
    from oracle import model as om, ops as oo
    coords, _ = cloud_points(2, 1024, 5)
    og = om.Geometry(coords, 1e-5, 0.1, "floor", "mul")
    og.downsample(1)
    og.downsample(2)
    g = torch.Generator().manual_seed(c)
    n = coords.shape[0]
    src = torch.randn(n, c, generator=g)
    idx = torch.from_numpy(og.point_to_voxel_idx(4))
    nv = og.level(4).n
    a = src.clone().requires_grad_(True)
    want = oo.scatter_mean(a, idx, nv)
    gout = torch.randn(nv, c, generator=g)
    want.backward(gout)
    b = src.cuda().requires_grad_(True)
    got = sp.functional.scatter_mean(b, idx.int().cuda(), nv)
    got.backward(gout.cuda())
    assert rel_err(got.detach().cpu().numpy(), want.detach().numpy()) < 1e-6
    assert rel_err(b.grad.cpu().numpy(), a.grad.numpy()) < 1e-6
    # devoxelize
    idx8, w8 = og.devox_idx_weights(4)
    vf = torch.randn(nv, c, generator=g)
    a = vf.clone().requires_grad_(True)
    want = oo.spdevoxelize(a, idx8, w8)
    gp = torch.randn(n, c, generator=g)
    want.backward(gp)
    b = vf.cuda().requires_grad_(True)
    got = sp.functional.devoxelize(b, idx8.int().cuda().contiguous(), w8.cuda().contiguous())
    got.backward(gp.cuda())
    assert rel_err(got.detach().cpu().numpy(), want.detach().numpy()) < 1e-6
    assert rel_err(b.grad.cpu().numpy(), a.grad.numpy()) < 1e-5


# ------------------------------------------------------------------------------------------ conv
def _conv_case(sp, kind, cin, cout, seed, B=2, N=1024):
    from oracle import model as om, ops as oo
    coords, _ = cloud_points(B, N, seed)
    og = om.Geometry(coords, 1e-5, 0.1, "spconv", "mul")
    geo = sp.Geometry(coords.float().cuda(), 1e-5, 0.1, strides=(1, 2, 4), p2v_strides=(1,), devox_strides=(1,),
                      downsample_mode="spconv")
    g = torch.Generator().manual_seed(seed)
    if kind == "k3":
        s = 2
        og.downsample(1)
        m, n_in, n_out, kvol, km = og.kmap3(s), og.level(s).n, og.level(s).n, 27, geo.level(s).kmap3
    elif kind == "down":
        _, _, _, md, _ = og.downsample(1)
        m, n_in, n_out, kvol, km = md, og.level(1).n, og.level(2).n, 8, geo.level(1).kmap_down
    elif kind == "up":
        _, _, _, _, mu = og.downsample(1)
        m, n_in, n_out, kvol, km = mu, og.level(2).n, og.level(1).n, 8, geo.level(1).kmap_up
    else:
        n_in = n_out = og.level(1).n
        m, kvol, km = np.arange(n_in, dtype=np.int32)[:, None], 1, None
    x = torch.randn(n_in, cin, generator=g)
    w = torch.randn(kvol, cin, cout, generator=g) / (kvol * cin) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    return x, w, b, m, n_out, km, oo


SHAPES = [("k3", 32, 32), ("k3", 64, 96), ("k3", 320, 192), ("down", 32, 64), ("up", 256, 192), ("k1", 192, 256),
          ("k3", 448, 256)]


@pytest.mark.parametrize("kind,cin,cout", SHAPES + [("k3", 3, 32), ("k3", 20, 24)])
def test_conv_forward_simt(sp, kind, cin, cout):
    x, w, b, m, n_out, km, oo = _conv_case(sp, kind, cin, cout, 11)
    want = oo.conv_gather_gemm_scatter(x, w, m, n_out, b)
    got = sp.functional.sparse_conv(x.cuda(), w.cuda() if kind != "k1" else w[0].cuda(), b.cuda(), km, None, sp.ops.CONV_SIMT)
    assert rel_err(got.cpu().numpy(), want.numpy()) < 1e-5


@pytest.mark.parametrize("kind,cin,cout", SHAPES)
def test_conv_forward_tcgen05(sp, kind, cin, cout):
    x, w, b, m, n_out, km, oo = _conv_case(sp, kind, cin, cout, 13)
    want = oo.conv_gather_gemm_scatter(x, w, m, n_out, b)
    want_tf32 = oo.conv_gather_gemm_scatter(x, w, m, n_out, b, emulate_tf32=True)
    got = sp.functional.sparse_conv(x.cuda(), w.cuda() if kind != "k1" else w[0].cuda(), b.cuda(), km,
                                    sp.functional.PackedWeight(), sp.ops.CONV_UMMA).cpu().numpy()
    assert rel_err(got, want_tf32.numpy()) < 2e-5          # same operand rounding as the oracle's emulation
    assert rel_err(got, want.numpy()) < 1e-3                # BASELINE.json tolerance vs plain fp32


@pytest.mark.parametrize("pre", [False, True])      # fp32 rows converted while gathered / rows already f16
@pytest.mark.parametrize("kind,cin,cout", SHAPES + [("k3", 96, 64), ("k1", 32, 256), ("k3", 160, 128)])
def test_conv_forward_tcgen05_f16_operands(sp, kind, cin, cout, pre):
    """kind::f16 path: operands rounded to IEEE half (the significand of TF32), fp32 accumulation; 64-channel slabs with
    a half-empty last slab when cin % 64 == 32"""
    x, w, b, m, n_out, km, oo = _conv_case(sp, kind, cin, cout, 19)
    want = oo.conv_gather_gemm_scatter(x, w, m, n_out, b)
    want_f16 = oo.conv_gather_gemm_scatter(x.half().float(), w.half().float(), m, n_out, b)
    wp = sp.ops.pack_weights_f16(w.cuda())
    xd = x.cuda().half() if pre else x.cuda()
    # mt = 6: the persistent stream-K kernel, mt = 7: CTA pairs (cta_group::2); both for f16 rows only
    for nsplit, mt in ((0, 0), (1, 0), (5, 0), (0, 1), (3, 3)):
        got = sp.ops.conv_fwd(xd, w.cuda(), wp, km.map_t if km is not None else None, km.mask if km is not None else None,
                              n_out, b.cuda(), sp.ops.CONV_UMMA, nsplit=nsplit, mt=mt).cpu().numpy()
        assert rel_err(got, want_f16.numpy()) < 2e-5        # same operand rounding as the emulation
        assert rel_err(got, want.numpy()) < 1e-3            # BASELINE.json tolerance vs plain fp32


@pytest.mark.parametrize("bn,cout", [(32, 384), (64, 384), (96, 384), (128, 384), (192, 384), (256, 256)])
def test_conv_tcgen05_slice_widths(sp, bn, cout):
    x, w, b, m, n_out, km, oo = _conv_case(sp, "k3", 64, cout, 17)
    want = oo.conv_gather_gemm_scatter(x, w, m, n_out, b, emulate_tf32=True)
    wp = sp.ops.pack_weights(w.cuda())
    got = sp.ops.conv_fwd(x.cuda(), w.cuda(), wp, km.map_t, km.mask, n_out, b.cuda(), sp.ops.CONV_UMMA, bn=bn)
    assert rel_err(got.cpu().numpy(), want.numpy()) < 2e-5


@pytest.mark.parametrize("impl", ["simt", "umma"])
@pytest.mark.parametrize("kind,cin,cout", [("k3", 32, 64), ("down", 64, 128), ("up", 128, 64), ("k1", 64, 32)])
def test_conv_backward(sp, kind, cin, cout, impl):
    x, w, b, m, n_out, km, oo = _conv_case(sp, kind, cin, cout, 19)
    g = torch.Generator().manual_seed(1)
    gout = torch.randn(n_out, cout, generator=g)
    xa, wa, ba = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    oo.conv_gather_gemm_scatter(xa, wa, m, n_out, ba).backward(gout)
    xb = x.cuda().requires_grad_(True)
    wb = (w.cuda() if kind != "k1" else w[0].cuda()).requires_grad_(True)
    bb = b.cuda().requires_grad_(True)
    code = sp.ops.CONV_SIMT if impl == "simt" else sp.ops.CONV_AUTO
    sp.functional.sparse_conv(xb, wb, bb, km, sp.functional.PackedWeight(), code).backward(gout.cuda())
    tol = 1e-5 if impl == "simt" else 1e-3
    assert rel_err(xb.grad.cpu().numpy(), xa.grad.numpy()) < tol
    assert rel_err(wb.grad.cpu().numpy().reshape(wa.shape), wa.grad.numpy()) < tol      # umma: tcgen05 wgrad (TF32 operands)
    assert rel_err(bb.grad.cpu().numpy(), ba.grad.numpy()) < 1e-5


@pytest.mark.parametrize("mt", [1, 3])
@pytest.mark.parametrize("cout", [32, 128, 192, 256])
def test_conv_tcgen05_tile_configs(sp, mt, cout):
    """mt: 1 = one tile per CTA with a four-stage ring, 3 = lite (2-3 stages, two CTAs per SM: the product default)"""
    x, w, b, m, n_out, km, oo = _conv_case(sp, "k3", 64, cout, 29)
    want = oo.conv_gather_gemm_scatter(x, w, m, n_out, b, emulate_tf32=True)
    xd = sp.ops.round_tf32(x.cuda())
    res = torch.randn(n_out, cout, generator=torch.Generator().manual_seed(2))
    wp = sp.ops.pack_weights(w.cuda())
    got = sp.ops.conv_fwd(xd, w.cuda(), wp, km.map_t, km.mask, n_out, b.cuda(), sp.ops.CONV_UMMA, a_tf32=True, mt=mt,
                          residual=res.cuda())
    assert rel_err(got.cpu().numpy(), (want + res).numpy()) < 2e-5


@pytest.mark.parametrize("stride,cin,cout", [(1, 32, 32), (1, 96, 64), (2, 64, 128), (2, 192, 192)])
def test_conv_pair_major(sp, stride, cin, cout):
    """pair lists (spvd_kmap_pairs) are exactly the valid entries of the map, and the pair-major tcgen05
    convolution (vector-RED scatter epilogue) equals the output-stationary result"""
    from oracle import model as om, ops as oo
    coords, _ = cloud_points(4, 2048, 31)
    og = om.Geometry(coords, 1e-5, 0.1, "spconv", "mul")
    geo = sp.Geometry(coords.float().cuda(), 1e-5, 0.1, strides=(1, 2), p2v_strides=(1,), devox_strides=(1,))
    if stride == 2:
        og.downsample(1)
    km = geo.level(stride).kmap3
    n = km.n_out
    pin, pout, tk, nt, npairs = sp.ops.kmap_pairs(km.map_t, n)
    m = og.kmap3(stride)                                      # [n,27]
    T, P = int(nt.item()), int(npairs.item())
    assert P == int((m >= 0).sum())
    got = set()
    pin_h, pout_h, tk_h = pin.cpu().numpy(), pout.cpu().numpy(), tk.cpu().numpy()
    for t in range(T):
        sl = slice(t * 128, (t + 1) * 128)
        v = pout_h[sl] >= 0
        assert np.all((pin_h[sl] >= 0) == v)
        got.update(zip([int(tk_h[t])] * int(v.sum()), pin_h[sl][v].tolist(), pout_h[sl][v].tolist()))
    o_idx, k_idx = np.nonzero(m >= 0)
    assert got == set(zip(k_idx.tolist(), m[o_idx, k_idx].tolist(), o_idx.tolist())) and len(got) == P
    g = torch.Generator().manual_seed(5)
    x = oo.round_tf32(torch.randn(n, cin, generator=g))
    w = torch.randn(27, cin, cout, generator=g) / (27 * cin) ** 0.5
    b, res = torch.randn(cout, generator=g), torch.randn(n, cout, generator=g)
    want = oo.conv_gather_gemm_scatter(x, w, m, n, b, emulate_tf32=True) + res
    y = sp.ops.conv_fwd_pairs(x.cuda(), w.cuda(), sp.ops.pack_weights(w.cuda()), pin, pout, tk, T, n, b.cuda(), res.cuda())
    assert rel_err(y.cpu().numpy(), want.numpy()) < 2e-5
    # the same tiles with f16 operands (rows already f16, and fp32 rows converted while gathered)
    want16 = oo.conv_gather_gemm_scatter(x.half().float(), w.half().float(), m, n, b) + res
    wp16 = sp.ops.pack_weights_f16(w.cuda())
    for xd in (x.cuda().half(), x.cuda()):
        y16 = sp.ops.conv_fwd_pairs(xd, w.cuda(), wp16, pin, pout, tk, T, n, b.cuda(), res.cuda())
        assert rel_err(y16.cpu().numpy(), want16.numpy()) < 2e-5


@pytest.mark.parametrize("cin,cout", [(3, 32), (4, 64), (32, 3), (64, 4), (3, 3)])
def test_linear_tiny_width_simt(sp, cin, cout):
    """the streaming kernels of the 3 -> 32 point embedding and the 32 -> 3 output head (and the generic fallback)"""
    g = torch.Generator().manual_seed(cin * 100 + cout)
    n = 5000
    x, w, b, r = torch.randn(n, cin, generator=g), torch.randn(1, cin, cout, generator=g), torch.randn(cout, generator=g), torch.randn(n, cout, generator=g)
    want = x @ w[0] + b + r
    got = sp.ops.conv_fwd(x.cuda(), w.cuda(), None, None, None, n, b.cuda(), sp.ops.CONV_SIMT, residual=r.cuda())
    assert rel_err(got.cpu().numpy(), want.numpy()) < 1e-6


def test_conv_pair_major_simt_stem(sp):
    """the 3 -> 32 stem convolution on the pair list (fp32 SIMT, RED scatter) equals the oracle"""
    import ctypes
    from spvd_b200._C import lib, check
    from oracle import model as om, ops as oo
    coords, _ = cloud_points(4, 2048, 41)
    og = om.Geometry(coords, 1e-5, 0.1, "spconv", "mul")
    geo = sp.Geometry(coords.float().cuda(), 1e-5, 0.1, strides=(1,), p2v_strides=(1,), devox_strides=(1,))
    km = geo.level(1).kmap3
    n = km.n_out
    pin, pout, tk, nt, _ = sp.ops.kmap_pairs(km.map_t, n)
    g = torch.Generator().manual_seed(3)
    for cin, cout in ((3, 32), (4, 48)):
        x, w, b = torch.randn(n, cin, generator=g), torch.randn(27, cin, cout, generator=g), torch.randn(cout, generator=g)
        want = oo.conv_gather_gemm_scatter(x, w, og.kmap3(1), n, b)
        xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
        out = torch.empty(n, cout, device="cuda")
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        check(lib.spvd_conv_fwd_pairs_simt(P(xd), P(wd), P(pin), P(pout), P(tk), int(nt.item()), 13, n, 27, cin, cout, P(bd), P(out),
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        assert rel_err(out.cpu().numpy(), want.numpy()) < 1e-5


@pytest.mark.parametrize("a_tf32", [True, False])
def test_conv_pair_major_strided_and_transposed(sp, a_tf32):
    """the stride-2 conv and its transpose on ONE pair list (roles of the two row arrays swapped)"""
    from oracle import model as om, ops as oo
    coords, _ = cloud_points(4, 2048, 37)
    og = om.Geometry(coords, 1e-5, 0.1, "spconv", "mul")
    geo = sp.Geometry(coords.float().cuda(), 1e-5, 0.1, strides=(1, 2), p2v_strides=(1,), devox_strides=(1,))
    _, _, _, md, mu = og.downsample(1)
    nf, nc = og.level(1).n, og.level(2).n
    km = geo.level(1).kmap_down
    pin, pout, tk, nt, npairs = sp.ops.kmap_pairs(km.map_t, nc, kvol=8)
    T = int(nt.item())
    assert int(npairs.item()) == int((md >= 0).sum())
    g = torch.Generator().manual_seed(7)
    x = torch.randn(nf, 64, generator=g)
    w = torch.randn(8, 64, 96, generator=g) / (8 * 64) ** 0.5
    xr = oo.round_tf32(x) if a_tf32 else x
    want = oo.conv_gather_gemm_scatter(xr, w, md, nc, None, emulate_tf32=True)
    y = sp.ops.conv_fwd_pairs(xr.cuda(), w.cuda(), sp.ops.pack_weights(w.cuda()), pin, pout, tk, T, nc, a_tf32=a_tf32)
    assert rel_err(y.cpu().numpy(), want.numpy()) < 2e-5
    xc = torch.randn(nc, 96, generator=g)
    wt = torch.randn(8, 96, 32, generator=g) / 96 ** 0.5
    xcr = oo.round_tf32(xc) if a_tf32 else xc
    want_up = oo.conv_gather_gemm_scatter(xcr, wt, mu, nf, None, emulate_tf32=True)
    yu = sp.ops.conv_fwd_pairs(xcr.cuda(), wt.cuda(), sp.ops.pack_weights(wt.cuda()), pout, pin, tk, T, nf, a_tf32=a_tf32)
    assert rel_err(yu.cpu().numpy(), want_up.numpy()) < 2e-5


@pytest.mark.parametrize("nsplit", [1, 3, 16])
@pytest.mark.parametrize("a_tf32", [False, True])
def test_conv_tcgen05_split_and_async_gather(sp, nsplit, a_tf32):
    """split-K accumulation (vector RED into a zeroed output) and the cp.async gather of pre-rounded rows"""
    x, w, b, m, n_out, km, oo = _conv_case(sp, "k3", 96, 128, 23)
    want = oo.conv_gather_gemm_scatter(x, w, m, n_out, b, emulate_tf32=True)
    xd = x.cuda()
    if a_tf32:
        xd = sp.ops.round_tf32(xd)
        assert torch.equal(xd.cpu(), oo.round_tf32(x))          # cvt.rna.tf32 == the oracle's rounding
    wp = sp.ops.pack_weights(w.cuda())
    got = sp.ops.conv_fwd(xd, w.cuda(), wp, km.map_t, km.mask, n_out, b.cuda(), sp.ops.CONV_UMMA, nsplit=nsplit,
                          a_tf32=a_tf32)
    assert rel_err(got.cpu().numpy(), want.numpy()) < 2e-5


@pytest.mark.parametrize("kind,cin,cout", [("k3", 32, 32), ("k3", 192, 128), ("k3", 320, 192), ("down", 64, 128),
                                           ("up", 256, 192), ("k1", 192, 256), ("k1", 448, 256)])
def test_conv_wgrad_tcgen05(sp, kind, cin, cout):
    """weight gradient on the tensor cores (MN-major gathered operands) vs the fp32 oracle (autograd) and vs the oracle
    with TF32-rounded operands"""
    x, w, b, m, n_out, km, oo = _conv_case(sp, kind, cin, cout, 41)
    g = torch.Generator().manual_seed(3)
    gout = torch.randn(n_out, cout, generator=g)
    wa = w.clone().requires_grad_(True)
    oo.conv_gather_gemm_scatter(x, wa, m, n_out, None).backward(gout)
    wt = w.clone().requires_grad_(True)
    oo.conv_gather_gemm_scatter(oo.round_tf32(x), wt, m, n_out, None).backward(oo.round_tf32(gout))
    xr, gr = sp.ops.round_tf32(x.cuda()), sp.ops.round_tf32(gout.cuda())
    got = sp.ops.conv_wgrad_umma(xr, gr, km.pairs() if km is not None else None, n_out, w.shape[0]).cpu().numpy()
    assert rel_err(got, wt.grad.numpy()) < 2e-5
    assert rel_err(got, wa.grad.numpy()) < 1e-3


def test_weighted_scatter_mean_equals_scatter_mean(sp):
    """planner-side 1/count weights + one RED pass (spvd_scatter_wmean_fwd) against spvd_scatter_mean_fwd"""
    import ctypes
    from spvd_b200._C import lib, check
    g = torch.Generator().manual_seed(1)
    n, nv, c = 50000, 3000, 64
    idx = torch.randint(0, nv, (n,), generator=g).int()
    idx[:5000] = 0                                      # the crowded row-0 sink of unmatched points
    idx[idx == 5] = 6                                   # voxel 5 gets no point
    src = torch.randn(n, c, generator=g)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    S = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    d_idx, d_src = idx.cuda(), src.cuda()
    cnt, w = torch.empty(nv, dtype=torch.int32, device="cuda"), torch.empty(n, device="cuda")
    check(lib.spvd_point_mean_weights(P(d_idx), n, nv, P(cnt), P(w), S))
    counts = torch.bincount(idx.long(), minlength=nv)
    assert torch.equal(cnt.cpu().long(), counts) and torch.equal(w.cpu(), 1.0 / counts[idx.long()].float())
    out = torch.empty(nv, c, device="cuda")
    check(lib.spvd_scatter_wmean_fwd(P(d_src), P(d_idx), P(w), n, c, P(out), nv, S))
    want, _ = sp.ops.scatter_mean_fwd(d_src, d_idx, nv)
    assert rel_err(out.cpu().numpy(), want.cpu().numpy()) < 1e-6 and float(out[5].abs().max()) == 0.0
    # the same as a segmented reduction over points grouped by voxel (voxels 0 and 7 hold thousands / hundreds of points: many units)
    idx2 = idx.clone()
    idx2[5000:5100] = 7
    counts2 = torch.bincount(idx2.long(), minlength=nv)
    d_idx2 = idx2.cuda()
    gc, gs, gcur, go, gh = (torch.empty(k, dtype=torch.int32, device="cuda") for k in (nv + 2, nv, nv, n, 2 * (nv + n // 32 + 1)))
    check(lib.spvd_group_points_by_voxel(P(d_idx2), n, nv, None, P(gc), P(gs), P(gcur), P(go), P(gh), S))
    o, st = go.cpu().long(), gs.cpu().tolist()
    assert torch.equal(gc[:nv].cpu().long(), counts2) and int(gc[nv]) == n and sorted(o.tolist()) == list(range(n))
    assert all(bool((idx2.long()[o[st[v]: st[v] + int(counts2[v])]] == v).all()) for v in range(nv))
    n_units = int(gc[nv + 1])
    assert n_units == int(((counts2 + 31) // 32).sum())
    recs = gh.cpu()[: 2 * n_units].view(-1, 2).tolist()                       # (voxel, chunk) records, one per <= 32 points
    assert sorted(map(tuple, recs)) == sorted((v, u) for v in range(nv) for u in range((int(counts2[v]) + 31) // 32))
    out2 = torch.empty(nv, c, device="cuda")
    check(lib.spvd_segment_mean_fwd(P(d_src), P(gs), P(gc), P(go), P(gh), n, nv, c, P(out2), nv, S))
    want2, _ = sp.ops.scatter_mean_fwd(d_src, d_idx2, nv)
    assert rel_err(out2.cpu().numpy(), want2.cpu().numpy()) < 1e-5 and float(out2[5].abs().max()) == 0.0


def test_fused_glue(sp):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1000, 64, generator=g).cuda()
    sc, sh = torch.randn(64, generator=g).cuda(), torch.randn(64, generator=g).cuda()
    want = torch.nn.functional.silu(x * sc + sh)
    assert rel_err(sp.ops.affine_act(x, sc, sh, 1).cpu().numpy(), want.cpu().numpy()) < 1e-6
    from oracle import ops as oo
    r = sp.ops.affine_act(x, sc, sh, 3).cpu()
    assert torch.equal(r, oo.round_tf32(r)) and rel_err(r.numpy(), want.cpu().numpy()) < 5e-4
    coords = torch.zeros(1000, 4, dtype=torch.int32)
    coords[:, 0] = torch.arange(1000) // 250
    tt = torch.randn(4, 128, generator=g).cuda()
    b = coords[:, 0].long().cuda()
    want = (1 + tt[b, :64]) * x + tt[b, 64:]
    assert rel_err(sp.ops.film(x, tt, coords.cuda()).cpu().numpy(), want.cpu().numpy()) < 1e-6


def test_layernorm_attention_cat_kernels(sp):
    import ctypes
    from spvd_b200._C import lib, check
    g = torch.Generator().manual_seed(3)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    S = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for C_, heads in ((192, 8), (256, 8), (64, 4), (384, 8)):
        counts = torch.tensor([37, 0, 300, 129, 1], dtype=torch.int32)
        n = int(counts.sum())
        off = torch.zeros(6, dtype=torch.int32)
        off[1:] = torch.cumsum(counts, 0)
        x = torch.randn(n, C_, generator=g).cuda()
        gam, bet = torch.randn(C_, generator=g).cuda(), torch.randn(C_, generator=g).cuda()
        y = torch.empty_like(x)
        check(lib.spvd_layernorm_fwd(P(x), P(gam), P(bet), n, C_, 1e-5, 0, P(y), S))
        want = torch.nn.functional.layer_norm(x, (C_,), gam, bet, 1e-5)
        assert rel_err(y.cpu().numpy(), want.cpu().numpy()) < 1e-5
        qkv = torch.randn(n, 3 * C_, generator=g).cuda()
        out = torch.zeros(n, C_).cuda()
        offd = off.cuda()
        check(lib.spvd_attention_fwd(P(qkv), P(offd), 5, int(counts.max()), C_, heads, 0, P(out), S))
        d = C_ // heads
        ref = torch.zeros(n, C_)
        q4 = qkv.cpu().view(n, 3, heads, d)
        for b in range(5):
            lo, hi = int(off[b]), int(off[b + 1])
            if hi == lo:
                continue
            q, k, v = (q4[lo:hi, j].transpose(0, 1) for j in range(3))       # [H, L, d]
            a = torch.softmax(q @ k.transpose(-2, -1) * d ** -0.5, dim=-1)
            ref[lo:hi] = (a @ v).transpose(0, 1).reshape(hi - lo, C_)
        assert rel_err(out.cpu().numpy(), ref.numpy()) < 1e-5
    a, b = torch.randn(100, 64, generator=g).cuda(), torch.randn(100, 32, generator=g).cuda()
    y = torch.empty(100, 96).cuda()
    check(lib.spvd_cat_fwd(P(a), P(b), 100, 64, 32, P(y), S))
    assert torch.equal(y, torch.cat([a, b], 1))
    # cat + BN + SiLU with both results stored as f16 (the up-path blocks of the executor), and the f16 stores of the
    # other producers (act & 4)
    sc, sh = torch.randn(96, generator=g).cuda(), torch.randn(96, generator=g).cuda()
    raw16, y16 = torch.empty(100, 96, dtype=torch.float16).cuda(), torch.empty(100, 96, dtype=torch.float16).cuda()
    check(lib.spvd_cat_affine_f16_fwd(P(a), P(b), 100, 64, 32, P(sc), P(sh), 1, P(raw16), P(y16), S))
    assert torch.equal(raw16, y.half())
    assert torch.equal(y16, torch.nn.functional.silu(y * sc + sh).half()) or \
        rel_err(y16.float().cpu().numpy(), torch.nn.functional.silu(y * sc + sh).cpu().numpy()) < 1e-3
    h = torch.empty(100, 96, dtype=torch.float16).cuda()
    check(lib.spvd_bn_silu_fwd(P(y), P(sc), P(sh), 100, 96, 1 | 4, P(h), S))
    assert rel_err(h.float().cpu().numpy(), torch.nn.functional.silu(y * sc + sh).cpu().numpy()) < 1e-3
    big = torch.full((4, 8), 1e6).cuda()
    hb = torch.empty(4, 8, dtype=torch.float16).cuda()
    one, zero = torch.ones(8).cuda(), torch.zeros(8).cuda()
    check(lib.spvd_bn_silu_fwd(P(big), P(one), P(zero), 4, 8, 4, P(hb), S))
    assert float(hb.max()) == 65504.0                      # clamped to the largest finite half, never inf


def test_fused_ddpm_step_matches_reference_formulas(sp):
    """DDPM.update_rule (utils/schedulers.py:78-89) + torch2sparse (utils/schedulers.py:249-257) in two kernels"""
    from spvd_b200.sampler import DDPM, torch2sparse
    g = torch.Generator().manual_seed(9)
    B, N = 5, 777
    x, eps, z = (torch.randn(B * N, 3, generator=g) for _ in range(3))
    s = DDPM()
    for t in (999, 400, 0):
        want = s.update(x, eps, t, 0, lambda shape: z)
        want_sp = torch2sparse(want.view(B, N, 3), 1e-5, order="keep")           # CPU tensors -> torch path
        new, coords, pc = sp.ops.ddpm_step(x.cuda(), B, N, 1e-5, eps.cuda(), z.cuda() if t > 0 else None, s.c1[t], s.c2[t], s.sig[t])
        assert torch.allclose(new.cpu(), want, atol=1e-6, rtol=1e-6)
        cpu_from_gpu = torch2sparse(new.cpu().view(B, N, 3), 1e-5, order="keep")  # same points -> identical integers
        assert torch.equal(coords.cpu(), cpu_from_gpu.C)
        assert torch.equal(pc.cpu(), cpu_from_gpu.C.float())
        assert (coords.cpu() - want_sp.C).abs().max() <= 1                        # 1-ulp update differences move at most one cell
    st = torch2sparse(x.view(B, N, 3).cuda(), 1e-5)                              # fused re-sparsify only
    assert torch.equal(st.C.cpu(), torch2sparse(x.view(B, N, 3), 1e-5).C) and st.points_per_cloud == N


def test_cpu_tensors_are_refused(sp):
    with pytest.raises(RuntimeError):
        sp.ops.point_voxel_coords(torch.zeros(4, 4), 1e-5, 0.1)


# ------------------------------------------------------------------------------------------------ round 2
def _level_maps(sp, B=6, N=2048, seed=3):
    from oracle import model as om
    g = torch.Generator().manual_seed(seed)
    pts = torch.randn(B, N, 3, generator=g).clamp(-3, 3)
    coords, feats, _ = om.torch2sparse(pts)
    geo = sp.Geometry(coords.cuda().float(), 1e-5, 0.1, strides=(1, 2, 4), p2v_strides=(1,), devox_strides=(1,), n_batch=B)
    return geo


def test_kmap_sort_matches_host_sort():
    """spvd_kmap_sort: rows ordered by their active-offset bit set (stable), map / tile masks / permutation bit-exact
    against a torch argsort of the same masks -- for the 3^3 map and both stride-2 maps."""
    import spvd_b200 as sp
    from spvd_b200 import ops
    geo = _level_maps(sp)
    for s in (1, 2, 4):
        lv = geo.level(s)
        for km, kvol in ((lv.kmap3, 27), (lv.kmap_down, 8), (lv.kmap_up, 8)):
            if km is None:
                continue
            n = km.n_out
            ms, tm, perm = ops.kmap_sort(km.map_t, n, kvol)
            ks = torch.arange(kvol, device="cuda")
            mask = ((km.map_t[:, :n] >= 0).long() << ks[:, None]).sum(0)
            hp = torch.argsort(mask, stable=True)
            assert torch.equal(perm[:n].long(), hp) and bool((perm[n:] == -1).all())
            assert torch.equal(ms[:, :n], km.map_t[:, hp]) and bool((ms[:, n:] == -1).all())
            pm = torch.zeros(km.map_t.shape[1], dtype=torch.int64, device="cuda")
            pm[:n] = mask[hp]
            want = pm.view(-1, 128).cpu().numpy()
            want = np.bitwise_or.reduce(want, axis=1)
            assert np.array_equal(tm.cpu().numpy().astype(np.int64) & 0xffffffff, want)


@pytest.mark.parametrize("srt", [False, True])
@pytest.mark.parametrize("bn", [0, 64])
def test_conv_sp_fused_epilogue(srt, bn):
    """spvd_conv_sp_f16 (persistent sorted-tile kernel): conv + second K source + bias + residual + FiLM, fp32 result and the
    f16 BN+SiLU operand, against the one-tile tcgen05 kernel + torch glue on the same f16 operands; repeated to catch
    pipeline races (the map-row ring once let a few rows per launch gather a refilled segment)."""
    import spvd_b200 as sp
    from spvd_b200 import ops
    geo = _level_maps(sp, B=8)
    lv = geo.level(2)
    km, n = lv.kmap3, lv.n
    cin, cout, cin2 = 96, 128, 160
    g = torch.Generator(device="cuda").manual_seed(0)
    w = torch.randn(27, cin, cout, device="cuda", generator=g) / (27 * cin) ** 0.5
    w2 = torch.randn(1, cin2, cout, device="cuda", generator=g) / cin2 ** 0.5
    wp, wp2 = ops.pack_weights_f16(w), ops.pack_weights_f16(w2)
    x, x2 = torch.randn(n, cin, device="cuda", generator=g).half(), torch.randn(n, cin2, device="cuda", generator=g).half()
    bias, res = torch.randn(cout, device="cuda", generator=g), torch.randn(n, cout, device="cuda", generator=g)
    film = torch.randn(8, 2 * cout + 64, device="cuda", generator=g)[:, 32:32 + 2 * cout]
    sc, sh = torch.rand(cout, device="cuda", generator=g) + 0.5, torch.randn(cout, device="cuda", generator=g)
    ref = ops.conv_fwd(x, w, wp, km.map_t, km.mask, n, bias, ops.CONV_UMMA, mt=3)
    ref = ref + ops.conv_fwd(x2, w2, wp2, None, None, n, None, ops.CONV_UMMA) + res
    fv = film[lv.coords[:, 0].long()]
    ref32 = (1 + fv[:, :cout]) * ref + fv[:, cout:]
    ref16 = torch.nn.functional.silu(ref32 * sc + sh)
    kw = dict(map_t=km.map_t, tile_mask=km.mask)
    if srt:
        ms, tm, perm = ops.kmap_sort(km.map_t, n, 27)
        kw = dict(map_t=ms, tile_mask=tm, out_rows=perm)
    for rep in range(12):
        o32, o16 = ops.conv_sp(x, wp, 27, cin, cout, n, bias=bias, residual=res, in2=x2, w2_packed=wp2, film=film,
                               coords=lv.coords_cap, aff=(sc, sh), silu=True, want32=True, want16=True, bn=bn, stages=(0, 2, 3)[rep % 3], **kw)
        assert rel_err(o32.cpu().numpy(), ref32.cpu().numpy()) < 1e-5
        assert rel_err(o16.float().cpu().numpy(), ref16.cpu().numpy()) < 1e-3
        assert float((o32 - ref32).abs().max()) < 1e-3 * float(ref32.abs().max())      # no stray rows
    # identity (kernel size 1) and f16-only output
    wl = torch.randn(1, cin, cout, device="cuda", generator=g) / cin ** 0.5
    _, o16 = ops.conv_sp(x, ops.pack_weights_f16(wl), 1, cin, cout, n, bias=bias, want32=False, want16=True)
    want = x.float() @ wl[0].half().float() + bias
    assert rel_err(o16.float().cpu().numpy(), want.cpu().numpy()) < 1e-3


def test_executor_rebuilds_after_inplace_parameter_update():
    """a compiled program holds packed copies of the weights; an in-place update in eval mode must not leave it
    stale."""
    import spvd_b200 as sp
    from oracle import model as om
    m = sp.build_model("SPVD_TINY")
    m.load_state_dict(om.make_state_dict("SPVD_TINY", seed=0))
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(1)
    pts = torch.randn(2, 512, 3, generator=g).clamp(-3, 3)
    coords, feats, _ = om.torch2sparse(pts)
    x, t = sp.SparseTensor(feats.cuda(), coords.cuda()), torch.tensor([500, 20]).cuda()
    with torch.no_grad():
        y0 = m((x, t)).clone()
        m.stem_conv.voxel_conv[3].kernel.mul_(1.5)
        m.out_conv[0].running_var.mul_(2.0)
        y1 = m((x, t)).clone()
        m.use_executor = False
        y2 = m((x, t))
    assert rel_err(y1.cpu().numpy(), y0.cpu().numpy()) > 1e-2           # the update is visible ...
    assert rel_err(y1.cpu().numpy(), y2.cpu().numpy()) < 1e-3           # ... and matches the op-by-op path on the new weights
