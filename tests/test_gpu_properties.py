"""GPU, BASELINE.json's full size (B=32 x 2048 points): size-independent properties of the CUDA path that need no
oracle run -- sortedness / uniqueness / idempotence of the voxel sets, kernel-map symmetry, consistency of the
stride-2 maps and the pair lists, partition of unity of the trilinear weights, linearity of the tensor-core
convolution, constants preserved by scatter-mean -> devoxelize."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    import spvd_b200
    from spvd_b200.sampler import torch2sparse
    g = torch.Generator().manual_seed(0)
    pts = torch.randn(32, 2048, 3, generator=g).clamp_(-3, 3)
    st = torch2sparse(pts.cuda())
    m = spvd_b200.build_model("SPVD")
    geos = {mode: spvd_b200.Geometry(st.C.float(), 1e-5, 0.1, strides=(1, 2, 4, 8, 16), p2v_strides=(1, 4, 16),
                                     devox_strides=(1, 4, 16), n_batch=32, downsample_mode=mode) for mode in ("spconv", "floor")}
    return spvd_b200, st, geos


def keys(c):
    c = c.long()
    return (c[:, 0] << 48) | ((c[:, 1] + 32768) << 32) | ((c[:, 2] + 32768) << 16) | (c[:, 3] + 32768)


def test_levels_sorted_unique_idempotent_and_nested(full):
    sp, st, geos = full
    for mode, geo in geos.items():
        for s in (1, 2, 4, 8, 16):
            c = geo.level(s).coords
            k = keys(c)
            assert bool((k[1:] > k[:-1]).all())                      # strictly increasing = sorted and unique
            again, cnt = sp.ops.unique_coords(c.contiguous())
            assert int(cnt) == c.shape[0] and torch.equal(again[: c.shape[0]], c)      # unique is idempotent
            assert sum(geo.level(s).batch_counts_host) == c.shape[0]
            if s > 1:                                                # every coarse voxel is floor(fine/2) of some fine voxel
                fine = geo.level(s // 2).coords.clone()
                fine[:, 1:] >>= 1
                assert bool(torch.isin(k, keys(fine)).all())
        # 'floor' keeps every parent: sizes dominate 'spconv'
    assert all(geos["floor"].level(s).n >= geos["spconv"].level(s).n for s in (2, 4, 8, 16))
    assert geos["floor"].level(1).n == geos["spconv"].level(1).n


def test_point_to_voxel_indices_hit_the_right_voxel(full):
    sp, st, geos = full
    geo = geos["floor"]                                            # no dropped parents: every point finds its voxel
    f = geo.fcoords
    for s in (1, 4, 16):
        want = torch.cat([f[:, :1].int(), torch.floor(f[:, 1:] / s).int()], 1)
        assert torch.equal(geo.level(s).coords[geo.p2v[s].long()], want)


def test_submanifold_map_symmetry_and_pair_lists(full):
    sp, st, geos = full
    geo = geos["spconv"]
    for s in (1, 2, 4):
        km = geo.level(s).kmap3
        n = km.n_out
        m = km.map_t[:, :n]
        assert torch.equal(m[13], torch.arange(n, device=m.device, dtype=torch.int32))     # centre offset = identity
        for k in (0, 5, 12):                                       # map[26-k][map[k][o]] == o
            o = torch.nonzero(m[k] >= 0).flatten()
            assert torch.equal(m[26 - k][m[k][o].long()].long(), o)
        assert int((m >= 0).sum(0).min()) >= 1
        pin, pout, tk, nt, npairs = sp.ops.kmap_pairs(km.map_t, n)
        assert int(npairs) == int((m >= 0).sum())
        T = int(nt)
        valid = pout[: T * 128] >= 0
        kk = tk[:T].repeat_interleave(128)[valid].long()
        assert torch.equal(m[kk, pout[: T * 128][valid].long()], pin[: T * 128][valid])    # every pair is a map entry


def test_strided_maps_are_mutually_transposed(full):
    sp, st, geos = full
    for geo in geos.values():
        for s in (1, 2, 4, 8):
            lv = geo.level(s)
            down, up = lv.kmap_down.map_t[:, : lv.kmap_down.n_out], lv.kmap_up.map_t[:, : lv.n]
            k, p = torch.nonzero(down >= 0, as_tuple=True)
            child = down[k, p].long()
            assert torch.equal(up[k, child].long(), p)               # up[k][child] == parent
            assert int((up >= 0).sum()) == int((down >= 0).sum())
            assert int((up >= 0).sum(0).max()) <= 1                  # a fine voxel has exactly one parent (or none: dropped)


def test_trilinear_weights_partition_of_unity(full):
    sp, st, geos = full
    geo = geos["floor"]
    for s in (1, 4, 16):
        idx8, w8 = geo.devox[s]
        assert bool((w8 >= 0).all()) and bool(((idx8 >= 0) | (w8 == 0)).all())
        tot = w8.sum(1)          # = S / (S + 1e-8): 1 up to the reference's epsilon (the own voxel, corner 0, always exists)
        assert bool((tot <= 1 + 1e-5).all()) and bool((tot > 0).all()) and float((tot > 0.999).float().mean()) > 0.999
        n = geo.level(s).n
        ones = torch.ones(n, 32, device="cuda")
        assert torch.allclose(sp.functional.devoxelize(ones, idx8, w8), tot[:, None].expand(-1, 32), atol=1e-5)
        const = torch.full((st.C.shape[0], 8), 3.25, device="cuda")
        assert torch.allclose(sp.functional.scatter_mean(const, geo.p2v[s], n), torch.full((n, 8), 3.25, device="cuda"), atol=1e-6)


@pytest.mark.parametrize("stride,cin,cout", [(1, 32, 32), (2, 64, 64), (4, 192, 192), (16, 448, 256)])
def test_conv_linearity_and_path_agreement_at_full_size(full, stride, cin, cout):
    """conv(a x + b y) = a conv(x) + b conv(y) (on TF32-representable inputs the operand rounding commutes only
    approximately: tolerance 1e-3), and the output-stationary, pair-major and fp32 SIMT paths agree."""
    sp, st, geos = full
    geo = geos["spconv"]
    km = geo.level(stride).kmap3
    n = km.n_out
    g = torch.Generator(device="cuda").manual_seed(stride)
    x = sp.ops.round_tf32(torch.randn(n, cin, device="cuda", generator=g))
    y = sp.ops.round_tf32(torch.randn(n, cin, device="cuda", generator=g))
    w = torch.randn(27, cin, cout, device="cuda", generator=g) / (27 * cin) ** 0.5
    wp = sp.ops.pack_weights(w)
    conv = lambda v: sp.ops.conv_fwd(sp.ops.round_tf32(v), w, wp, km.map_t, km.mask, n, None, sp.ops.CONV_UMMA, a_tf32=True)
    lhs, rhs = conv(2.0 * x - 0.5 * y), 2.0 * conv(x) - 0.5 * conv(y)
    assert float((lhs - rhs).norm() / rhs.norm()) < 1e-3
    simt = sp.ops.conv_fwd(x, w, None, km.map_t, km.mask, n, None, sp.ops.CONV_SIMT)
    assert float((conv(x) - simt).norm() / simt.norm()) < 1e-3
    pin, pout, tk, nt, _ = sp.ops.kmap_pairs(km.map_t, n)
    pairs = sp.ops.conv_fwd_pairs(x, w, wp, pin, pout, tk, int(nt), n)
    assert float((pairs - conv(x)).norm() / simt.norm()) < 1e-5


def test_full_model_is_permutation_equivariant_over_points(full):
    """every op is a set function of the points: shuffling the input rows shuffles the output rows"""
    sp, st, geos = full
    from oracle import model as om
    m = sp.build_model("SPVD")
    m.load_state_dict(om.make_state_dict("SPVD", 0))
    m = m.cuda().eval()
    t = torch.randint(0, 1000, (32,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    perm = torch.randperm(st.C.shape[0], device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    with torch.no_grad():
        a = m((sp.SparseTensor(st.F, st.C), t))
        b = m((sp.SparseTensor(st.F[perm].contiguous(), st.C[perm].contiguous()), t))
    # atomics change the summation order of the scatter-means; a 1-ulp difference can flip a TF32 rounding downstream
    assert float((b - a[perm]).norm() / a.norm()) < 5e-4
