import os, sys, itertools, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, spvd_b200
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
for s, cin, cout in [(2, 128, 128), (2, 192, 128), (4, 320, 192), (4, 192, 192), (4, 128, 128), (8, 192, 192), (8, 384, 192), (16, 448, 256), (16, 256, 256)]:
    lv = geo.level(s); km = lv.kmap3
    w = torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5
    wp = ops.pack_weights_f16(w)
    xin = torch.randn(lv.n, cin, device=dev).half()
    bias = torch.randn(cout, device=dev)
    out = []
    for bn in [b for b in (96, 128, 192, 256) if cout % b == 0]:
        for mt in (3, 7):
            for ns in (0, 1, 2, 4, 8):
                if bn * (2 if mt == 2 else 1) > 512: continue
                try:
                    us = timeit(lambda: ops.conv_fwd(xin, w, wp, km.map_t, km.mask, lv.n, bias, ops.CONV_UMMA, bn=bn, nsplit=ns, a_tf32=True, mt=mt))
                except RuntimeError:
                    continue
                out.append((us, bn, mt, ns))
    auto = timeit(lambda: ops.conv_fwd(xin, w, wp, km.map_t, km.mask, lv.n, bias, ops.CONV_UMMA, a_tf32=True))
    print(f"s={s} n={lv.n} tiles={(lv.n+127)//128} {cin}->{cout} auto={auto:.1f}")
    for mt in (3, 7):
        print("   mt", mt, " ".join(f"bn{bn}/ns{ns}:{us:.0f}" for us, bn, m, ns in sorted(out, key=lambda r: (r[1], r[3])) if m == mt))
