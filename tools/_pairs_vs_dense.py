import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, spvd_b200
from spvd_b200 import ops
from spvd_b200.sampler import torch2sparse, DDPM
dev = torch.device("cuda", 0)
strat = DDPM()
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
for t in (50, 300, 600, 900, 999):
    g = torch.Generator().manual_seed(1234)
    x0 = torch.randn(32, 2048, 3, generator=g); x0 = x0 / x0.norm(dim=-1, keepdim=True) * math.sqrt(3.0)
    ah = float(strat.alpha_hat[t])
    x = math.sqrt(ah) * x0 + math.sqrt(1 - ah) * torch.randn(32, 2048, 3, generator=g)
    st = torch2sparse(x.to(dev))
    geo = spvd_b200.Geometry(st.C.float(), 1e-5, 0.1, strides=(1, 2, 4, 8, 16), p2v_strides=(1,), devox_strides=(1,), n_batch=32)
    for s, cin, cout in [(1, 32, 32), (1, 64, 64), (1, 96, 64), (2, 64, 64), (2, 128, 128), (2, 192, 128), (4, 128, 128), (4, 192, 192)]:
        lv = geo.level(s); km = lv.kmap3
        P = int((km.map_t[:, :lv.n] >= 0).sum())
        w = torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5
        wp = ops.pack_weights_f16(w)
        xin = torch.randn(lv.n, cin, device=dev).half()
        bias = torch.randn(cout, device=dev)
        dense = timeit(lambda: ops.conv_fwd(xin, w, wp, km.map_t, km.mask, lv.n, bias, ops.CONV_UMMA))
        pin, pout, tk, nt, npairs = ops.kmap_pairs(km.map_t, lv.n)
        T = int(nt.item())
        pairs = timeit(lambda: ops.conv_fwd_pairs(xin, w, wp, pin, pout, tk, T, lv.n, bias))
        print(f"t={t:3d} s={s} n={lv.n:6d} {cin:3d}->{cout:3d} P/n={P/lv.n:5.2f} dense={dense:6.1f} pairs={pairs:6.1f} {'PAIRS' if pairs < dense else 'dense'}")
