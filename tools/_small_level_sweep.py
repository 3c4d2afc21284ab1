"""Sweep (slice width, split-K) of the f16 tcgen05 conv on the small levels (strides 8 / 16) and the k=1 layers there, against
the automatic choice.  usage: python tools/_small_level_sweep.py [t]"""
import os, sys, itertools, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spvd_b200
from spvd_b200 import ops
from spvd_b200.sampler import torch2sparse, DDPM

t = int(sys.argv[1]) if len(sys.argv) > 1 else 500
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1234)
x0 = torch.randn(32, 2048, 3, generator=g); x0 = x0 / x0.norm(dim=-1, keepdim=True) * math.sqrt(3.0)
ah = float(DDPM().alpha_hat[t])
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
    return sorted(ts)[len(ts) // 2]


if "points" in sys.argv[2:]:
    # the point-branch Linear layers: one row per POINT (65536), k = 1
    n = 32 * 2048
    for cin, cout in ((32, 256), (256, 192), (192, 32), (96, 64), (64, 32)):
        w = torch.randn(1, cin, cout, device=dev) / cin ** 0.5
        wp = ops.pack_weights_f16(w)
        xin = torch.randn(n, cin, device=dev).half()
        bias = torch.randn(cout, device=dev)
        call = lambda bn, ns: ops.conv_fwd(xin, w, wp, None, None, n, bias, ops.CONV_UMMA, bn=bn, nsplit=ns)
        auto = timeit(lambda: call(0, 0))
        res = sorted((timeit(lambda: call(bn, 1), 5), bn) for bn in (32, 64, 96, 128, 192, 256) if cout % bn == 0)
        mb = n * (cin * 2 + cout * 4) / 1e6
        print(f"points n={n} k=1 {cin:3d}->{cout:3d} | {mb:5.1f} MB | auto {auto:6.1f} us ({mb / auto * 1e3 / 1e3:.2f} TB/s) | " +
              "  ".join(f"{u:5.1f} (bn={b})" for u, b in res), flush=True)
    sys.exit(0)
layers = [(8, 27, 192, 192), (8, 27, 384, 192), (16, 27, 192, 192), (16, 27, 448, 256), (16, 27, 256, 256), (4, 27, 128, 128),
          (8, 1, 192, 576), (8, 1, 192, 192), (8, 1, 384, 192), (16, 1, 256, 768), (16, 1, 256, 256), (16, 1, 448, 256)]
for s, kvol, cin, cout in layers:
    lv = geo.level(s); n = lv.n
    km = lv.kmap3 if kvol == 27 else None
    w = torch.randn(kvol, cin, cout, device=dev) / (kvol * cin) ** 0.5
    wp = ops.pack_weights_f16(w)
    xin = torch.randn(n, cin, device=dev).half()
    bias = torch.randn(cout, device=dev)
    call = lambda bn, ns: ops.conv_fwd(xin, w, wp, km.map_t if km else None, km.mask if km else None, n, bias, ops.CONV_UMMA, bn=bn, nsplit=ns)
    auto = timeit(lambda: call(0, 0))
    res = []
    for bn in [b for b in (32, 64, 96, 128, 192, 256) if cout % b == 0]:
        for ns in (1, 2, 3, 4, 6, 8, 12, 16):
            try:
                res.append((timeit(lambda: call(bn, ns), 5), bn, ns))
            except RuntimeError:
                pass
    res.sort()
    print(f"s={s:2d} n={n:6d} k={kvol:2d} {cin:3d}->{cout:3d} | auto {auto:6.1f} us | best " + "  ".join(f"{u:5.1f} (bn={b},split={k})" for u, b, k in res[:5]), flush=True)
