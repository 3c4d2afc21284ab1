"""Experiment + correctness check: persistent sorted-tile convolution (spvd_conv_sp_f16) and spvd_kmap_sort on the SPVD
layer shapes.  usage: python tools/_sp_sweep.py [t]"""
import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spvd_b200
from spvd_b200 import ops
from spvd_b200.sampler import torch2sparse, DDPM

t = int(sys.argv[1]) if len(sys.argv) > 1 else 500
dev = torch.device("cuda", 0)
strat = DDPM()
g = torch.Generator().manual_seed(1234)
x0 = torch.randn(32, 2048, 3, generator=g); x0 = x0 / x0.norm(dim=-1, keepdim=True) * math.sqrt(3.0)
ah = float(strat.alpha_hat[t])
x = math.sqrt(ah) * x0 + math.sqrt(1 - ah) * torch.randn(32, 2048, 3, generator=g)
st = torch2sparse(x.to(dev))
geo = spvd_b200.Geometry(st.C.float(), 1e-5, 0.1, strides=(1, 2, 4, 8, 16), p2v_strides=(1,), devox_strides=(1,), n_batch=32)
print({s: geo.level(s).n for s in (1, 2, 4, 8, 16)})
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return min(ts)


def host_sort(map_t, n):
    kvol, ld = map_t.shape
    ks = torch.arange(kvol, device=dev)
    mask = ((map_t[:, :n] >= 0).long() << ks[:, None]).sum(0)
    return torch.argsort(mask, stable=True).int(), mask


# ---- kmap_sort vs host
for s in (1, 2, 4, 8, 16):
    lv = geo.level(s)
    for km, kvol in ((lv.kmap3, 27), (lv.kmap_down, 8), (lv.kmap_up, 8)):
        if km is None:
            continue
        n = km.n_out
        ms, tm, perm = ops.kmap_sort(km.map_t, n, kvol)
        hp, mask = host_sort(km.map_t, n)
        assert torch.equal(perm[:n], hp), (s, kvol)
        assert bool((perm[n:] == -1).all())
        assert torch.equal(ms[:, :n], km.map_t[:, hp.long()])
        assert bool((ms[:, n:] == -1).all())
        pm = torch.zeros(km.map_t.shape[1], dtype=torch.int64, device=dev); pm[:n] = mask[hp.long()]
        acc = pm.view(-1, 128)[:, 0].clone()
        for j in range(1, 128):
            acc |= pm.view(-1, 128)[:, j]
        assert torch.equal(tm.long() & 0xffffffff, acc), (s, kvol)
    us = timeit(lambda: ops.kmap_sort(lv.kmap3.map_t, lv.n, 27))
    print(f"kmap_sort stride {s}: ok, {us:.1f} us (incl. allocations)")

# ---- fused epilogue correctness on one layer
lv = geo.level(4); km = lv.kmap3; n = lv.n
cin, cout, cin2 = 192, 128, 320
w = torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5
w2 = torch.randn(1, cin2, cout, device=dev) / cin2 ** 0.5
wp, wp2 = ops.pack_weights_f16(w), ops.pack_weights_f16(w2)
xin = torch.randn(n, cin, device=dev).half()
xin2 = torch.randn(n, cin2, device=dev).half()
bias = torch.randn(cout, device=dev)
res = torch.randn(n, cout, device=dev)
film = torch.randn(32, 2 * cout + 64, device=dev)[:, 32:32 + 2 * cout]     # a strided view: row stride 2*cout+64
sc, sh = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev)
ms, tm, perm = ops.kmap_sort(km.map_t, n, 27)
ref = ops.conv_fwd(xin, w, wp, km.map_t, km.mask, n, bias, ops.CONV_UMMA, mt=3)
ref = ref + ops.conv_fwd(xin2, w2, wp2, None, None, n, None, ops.CONV_UMMA) + res
b = lv.coords[:, 0].long()
fv = film[b]
ref_f = (1 + fv[:, :cout]) * ref + fv[:, cout:]
ref16 = torch.nn.functional.silu(ref_f * sc + sh)
for name, kw in (("sorted", dict(map_t=ms, tile_mask=tm, out_rows=perm)), ("lex", dict(map_t=km.map_t, tile_mask=km.mask))):
    for bn in (0, 64, 128):
        o32, o16 = ops.conv_sp(xin, wp, 27, cin, cout, n, bias=bias, residual=res, in2=xin2, w2_packed=wp2,
                               film=film, coords=lv.coords_cap, aff=(sc, sh), silu=True, want32=True, want16=True, bn=bn, **kw)
        e32 = float((o32 - ref_f).norm() / ref_f.norm()); e16 = float((o16.float() - ref16).norm() / ref16.norm())
        print(f"fused {name} bn={bn}: err32={e32:.2e} err16={e16:.2e}")
        assert e32 < 1e-5 and e16 < 1e-3
# identity (k = 1) path
wl = torch.randn(1, cin, cout, device=dev) / cin ** 0.5
wpl = ops.pack_weights_f16(wl)
o32, _ = ops.conv_sp(xin, wpl, 1, cin, cout, n, bias=bias)
refl = xin.float() @ wl[0].half().float() + bias
print("k=1 err", float((o32 - refl).norm() / refl.norm()))

# ---- timing
for s, cin, cout in [(1, 32, 32), (1, 64, 64), (1, 96, 64), (2, 64, 64), (2, 128, 128), (2, 192, 128), (4, 128, 128), (4, 320, 192), (4, 192, 192),
                     (8, 192, 192), (8, 384, 192), (16, 448, 256), (16, 256, 256)]:
    lv = geo.level(s); km = lv.kmap3; n = lv.n
    w = torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5
    wp = ops.pack_weights_f16(w)
    xin = torch.randn(n, cin, device=dev).half()
    bias = torch.randn(cout, device=dev)
    ref = ops.conv_fwd(xin, w, wp, km.map_t, km.mask, n, bias, ops.CONV_UMMA, mt=3)
    ms, tm, perm = ops.kmap_sort(km.map_t, n, 27)
    line = f"s={s:2d} n={n:6d} {cin:3d}->{cout:3d} |"
    old = timeit(lambda: ops.conv_fwd(xin, w, wp, km.map_t, km.mask, n, bias, ops.CONV_UMMA))
    line += f" old-auto:{old:6.1f}"
    if s <= 2:
        pin, pout, tk, nt, npairs = ops.kmap_pairs(km.map_t, n)
        T = int(nt.item())
        line += f" pairs:{timeit(lambda: ops.conv_fwd_pairs(xin, w, wp, pin, pout, tk, T, n, bias)):6.1f}"
    for name, kw in (("srt", dict(map_t=ms, tile_mask=tm, out_rows=perm)), ("lex", dict(map_t=km.map_t, tile_mask=km.mask))):
        for bn in sorted({0} | {b for b in (64, 96, 128, 192, 256) if cout % b == 0}):
            fn = lambda: ops.conv_sp(xin, wp, 27, cin, cout, n, bias=bias, bn=bn, **kw)
            o32, _ = fn()
            err = float((o32 - ref).norm() / ref.norm())
            line += f" {name}/bn{bn}:{timeit(fn):6.1f}" + ("" if err < 1e-5 else f"(ERR {err:.1e})")
    print(line, flush=True)
# strided maps (k = 8)
for s, cin, cout in [(1, 32, 64), (2, 64, 128), (4, 128, 192), (8, 192, 192)]:
    lv = geo.level(s); n_f, n_c = lv.n, geo.level(2 * s).n
    for kind, km, ci, co in (("down", lv.kmap_down, cin, cout), ("up", lv.kmap_up, cout, cin)):
        w = torch.randn(8, ci, co, device=dev) / (8 * ci) ** 0.5
        wp = ops.pack_weights_f16(w)
        xin = torch.randn(km.n_in, ci, device=dev).half()
        ref = ops.conv_fwd(xin, w, wp, km.map_t, km.mask, km.n_out, None, ops.CONV_UMMA, mt=3)
        ms, tm, perm = ops.kmap_sort(km.map_t, km.n_out, 8)
        old = timeit(lambda: ops.conv_fwd(xin, w, wp, km.map_t, km.mask, km.n_out, None, ops.CONV_UMMA))
        pin, pout, tk, nt, npairs = ops.kmap_pairs(km.map_t, km.n_out, 8)
        T = int(nt.item())
        pr = timeit(lambda: ops.conv_fwd_pairs(xin, w, wp, pin, pout, tk, T, km.n_out, None))
        fn = lambda: ops.conv_sp(xin, wp, 8, ci, co, km.n_out, map_t=ms, tile_mask=tm, out_rows=perm)
        o32, _ = fn()
        err = float((o32 - ref).norm() / ref.norm())
        print(f"{kind} s={s} {ci}->{co} n_out={km.n_out}: old-dense {old:6.1f} pairs {pr:6.1f} sp-sorted {timeit(fn):6.1f} err={err:.1e}", flush=True)
