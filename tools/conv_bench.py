"""Time the tcgen05 sparse conv on the real layer shapes of SPVD at B=32x2048 with different tile configurations.
usage: python tools/conv_bench.py [t]   (t = diffusion timestep of the synthetic input, default 500)"""
import os, sys, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, spvd_b200
from spvd_b200 import ops
from spvd_b200.sampler import torch2sparse, DDPM
import math

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
layers = [(1, 32, 32), (1, 96, 64), (1, 64, 64), (2, 64, 64), (2, 192, 128), (2, 128, 128), (4, 128, 128), (4, 320, 192),
          (4, 192, 192), (8, 192, 192), (8, 384, 192), (16, 192, 192), (16, 448, 256), (16, 256, 256)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return min(ts)

for s, cin, cout in layers:
    lv = geo.level(s); km = lv.kmap3
    P = int((km.map_t[:, :lv.n] >= 0).sum())
    w = torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5
    wp = ops.pack_weights(w)
    xin = ops.round_tf32(torch.randn(lv.n, cin, device=dev))
    bias = torch.randn(cout, device=dev)
    res = []
    bns = [b for b in (32, 64, 96, 128, 192, 256) if cout % b == 0]
    for bn, mt, ns in itertools.product(bns, (2, 3, 4, 5), (0, 1, 2, 4)):
        if bn * min(mt, 2) > 512: continue
        try:
            us = timeit(lambda: ops.conv_fwd(xin, w, wp, km.map_t, km.mask, lv.n, bias, ops.CONV_UMMA, bn=bn, nsplit=ns, a_tf32=True, mt=mt))
        except RuntimeError as e:
            continue
        res.append((us, bn, mt, ns))
    res.sort()
    pair_us = None
    if s <= 4:
        pin, pout, tk, nt, npairs = ops.kmap_pairs(km.map_t, lv.n)
        T = int(nt.item())
        pair_us = timeit(lambda: ops.conv_fwd_pairs(xin, w, wp, pin, pout, tk, T, lv.n, bias))
        kp_us = timeit(lambda: ops.kmap_pairs(km.map_t, lv.n))
        print(f"     pair-major: {pair_us:7.1f}us  tiles={T}  (pair list build {kp_us:.1f}us)")
    auto = timeit(lambda: ops.conv_fwd(xin, w, wp, km.map_t, km.mask, lv.n, bias, ops.CONV_UMMA, a_tf32=True))
    fl = 2.0 * P * cin * cout
    print(f"s={s:2d} n={lv.n:6d} {cin:3d}->{cout:3d} P/n={P/lv.n:5.2f} auto={auto:7.1f}us ({fl/auto/1e6:6.1f} TF) best: " +
          " | ".join(f"{us:6.1f}us bn={bn} mt={mt} ns={ns}" for us, bn, mt, ns in res[:4]))
