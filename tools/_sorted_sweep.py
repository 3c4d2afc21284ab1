"""Experiment: mask-sorted tiles / weight multicast / two tiles per CTA for the f16 tcgen05 conv on the SPVD layer shapes.
usage: python tools/_sorted_sweep.py [t]"""
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


def sort_map(map_t, n, groups=1, g=0):
    """-> perm [n] int32, map_sorted [kvol, ld], tile_mask [ld/128] for offsets of group g (others dropped)"""
    kvol, ld = map_t.shape
    ks = torch.arange(kvol, device=dev)
    sel = (ks * groups // kvol) == g
    valid = (map_t[:, :n] >= 0) & sel[:, None]
    mask = (valid.long() << ks[:, None]).sum(0)
    perm = torch.argsort(mask, stable=True)
    nz = int((mask == 0).sum())          # rows without any offset of this group: moved to the end (their tiles do no MMA)
    perm = torch.cat([perm[nz:], perm[:nz]])
    m = perm.numel()
    ms = torch.full((kvol, ld), -1, dtype=torch.int32, device=dev)
    ms[:, :m] = torch.where(valid[:, perm], map_t[:, perm], torch.full_like(map_t[:, perm], -1))
    tm = torch.zeros(ld // 128, dtype=torch.int64, device=dev)
    pm = torch.zeros(ld, dtype=torch.int64, device=dev); pm[:m] = mask[perm]
    pm = pm.view(-1, 128)
    acc = pm[:, 0].clone()
    for j in range(1, 128):
        acc |= pm[:, j]
    return perm.int().contiguous(), ms.contiguous(), acc.int().contiguous(), m


for s, cin, cout in [(1, 32, 32), (1, 64, 64), (1, 96, 64), (2, 64, 64), (2, 128, 128), (2, 192, 128), (4, 128, 128), (4, 320, 192), (4, 192, 192),
                     (8, 192, 192), (8, 384, 192), (16, 448, 256), (16, 256, 256)]:
    lv = geo.level(s); km = lv.kmap3
    n = lv.n
    P = int((km.map_t[:, :n] >= 0).sum())
    w = torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5
    wp = ops.pack_weights_f16(w)
    xin = torch.randn(n, cin, device=dev).half()
    bias = torch.randn(cout, device=dev)
    ref = ops.conv_fwd(xin, w, wp, km.map_t, km.mask, n, bias, ops.CONV_UMMA, mt=3)
    perm, ms, tm, m = sort_map(km.map_t, n)
    issued_lex = int(sum(bin(v & 0x7ffffff).count("1") for v in km.mask[: (n + 127) // 128].tolist())) * 128
    issued_srt = int(sum(bin(v & 0x7ffffff).count("1") for v in tm[: (n + 127) // 128].tolist())) * 128
    line = f"s={s:2d} n={n:6d} {cin:3d}->{cout:3d} P={P} issued lex={issued_lex} sorted={issued_srt} |"
    for name, kw in (("lex", dict(map_t=km.map_t, mask=km.mask)), ("srt", dict(map_t=ms, mask=tm, out_rows=perm))):
        for mt in (3, 4, 2):
            try:
                fn = lambda: ops.conv_fwd(xin, w, wp, kw["map_t"], kw["mask"], n, bias, ops.CONV_UMMA, mt=mt, out_rows=kw.get("out_rows"))
                out = fn()
                err = float((out - ref).norm() / ref.norm())
                us = timeit(fn)
                line += f" {name}/mt{mt}:{us:6.1f}us" + ("" if err < 1e-5 else f"(ERR {err:.1e})")
            except RuntimeError as e:
                line += f" {name}/mt{mt}:fail"
    if s <= 4:
        pin, pout, tk, nt, npairs = ops.kmap_pairs(km.map_t, n)
        T = int(nt.item())
        us = timeit(lambda: ops.conv_fwd_pairs(xin, w, wp, pin, pout, tk, T, n, bias))
        line += f" pairs:{us:6.1f}us"
    print(line, flush=True)
    # split masks: 2 / 3 groups, sequential launches (first plain, others with residual = previous output)
    if s <= 4:
        for G in (2, 3):
            parts = [sort_map(km.map_t, n, G, gi) for gi in range(G)]
            def run():
                out = None
                for gi, (pp, mm, tt, m_) in enumerate(parts):
                    # group 0 holds the centre offset for G in (1,3); for G=2 centre k=13 is in group 0 (13*2//27 = 0)
                    o2 = ops.conv_fwd(xin, w, wp, mm, tt, n, bias if gi == 0 else None, ops.CONV_UMMA, mt=3, out_rows=pp, residual=out)
                    out = o2
                return out
            err = float((run() - ref).norm() / ref.norm())
            us = timeit(run)
            iss = sum(int(sum(bin(v & 0x7ffffff).count("1") for v in tt[: (m_ + 127) // 128].tolist())) * 128 for _, _, tt, m_ in parts)
            print(f"      split{G}: issued={iss} time={us:6.1f}us err={err:.1e}", flush=True)
