import os, sys, math
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=9):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return min(ts)
tot = 0
for s, cin, cout in [(1, 32, 32), (1, 64, 64), (2, 64, 64), (2, 128, 128), (2, 192, 128), (4, 320, 192), (4, 192, 192), (4, 128, 128), (8, 192, 192), (8, 384, 192), (16, 448, 256), (16, 256, 256), (16, 192, 192)]:
    lv = geo.level(s); km = lv.kmap3
    w = torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5
    wp = ops.pack_weights(w)
    xin = ops.round_tf32(torch.randn(lv.n, cin, device=dev))
    bias = torch.randn(cout, device=dev)
    auto = timeit(lambda: ops.conv_fwd(xin, w, wp, km.map_t, km.mask, lv.n, bias, ops.CONV_UMMA, a_tf32=True))
    ref = ops.conv_fwd(xin, w, None, km.map_t, km.mask, lv.n, bias, ops.CONV_SIMT)
    got = ops.conv_fwd(xin, w, wp, km.map_t, km.mask, lv.n, bias, ops.CONV_UMMA, a_tf32=True)
    err = float((got - ref).abs().max() / ref.abs().max())
    tot += auto
    print(f"s={s} n={lv.n} {cin}->{cout} auto={auto:.1f} relerr={err:.1e}")
for rows, cin, cout in [(65536, 32, 256), (65536, 256, 192), (64035, 192, 128), (55435, 320, 192), (64035, 96, 64), (8259, 192, 576), (1849, 256, 768)]:
    w = torch.randn(1, cin, cout, device=dev) / cin ** 0.5
    wp = ops.pack_weights(w)
    xin = ops.round_tf32(torch.randn(rows, cin, device=dev))
    bias = torch.randn(cout, device=dev)
    res = torch.randn(rows, cout, device=dev)
    auto = timeit(lambda: ops.conv_fwd(xin, w, wp, None, None, rows, bias, ops.CONV_UMMA, a_tf32=True, residual=res))
    wp16 = ops.pack_weights_f16(w); x16 = xin.half()
    f16 = timeit(lambda: ops.conv_fwd(x16, w, wp16, None, None, rows, bias, ops.CONV_UMMA, residual=res))
    f16c = timeit(lambda: ops.conv_fwd(xin, w, wp16, None, None, rows, bias, ops.CONV_UMMA, residual=res))
    mb = rows * (cin * 2 + cout * 8) / 1e6
    tot += auto
    print(f"linear rows={rows} {cin}->{cout} tf32={auto:.1f} f16={f16:.1f} f16(convert)={f16c:.1f}  ideal(f16 in, res+out fp32)={mb/6.4e3*1e3/1e3:.1f}us")
print("total", round(tot, 1))
print("bn sweep for the point-level linears (f16 rows, residual)")
for rows, cin, cout in [(65536, 32, 256), (65536, 256, 192), (64035, 192, 128), (55435, 320, 192)]:
    w = torch.randn(1, cin, cout, device=dev) / cin ** 0.5
    wp16 = ops.pack_weights_f16(w); x16 = torch.randn(rows, cin, device=dev).half()
    bias = torch.randn(cout, device=dev); res = torch.randn(rows, cout, device=dev)
    out = []
    for bn in (32, 64, 96, 128, 192, 256):
        if cout % bn: continue
        out.append((bn, timeit(lambda: ops.conv_fwd(x16, w, wp16, None, None, rows, bias, ops.CONV_UMMA, residual=res, bn=bn))))
    print(rows, cin, cout, " ".join(f"bn{b}:{u:.1f}" for b, u in out))
