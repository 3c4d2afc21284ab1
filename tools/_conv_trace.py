"""Experiment: where does a CTA of the f16 tcgen05 conv spend its time?  Builds tools/trace/conv_trace.cu (-DSPVD_TRACE) on
the GPU box and prints per-phase clock statistics for a few SPVD layers.  usage: python tools/_conv_trace.py"""
import ctypes, os, subprocess, sys, math
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spvd_b200
from spvd_b200 import ops
from spvd_b200.sampler import torch2sparse, DDPM
sys.path.insert(0, os.path.join(ROOT, "tools"))

so = "/tmp/libconvtrace.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                       "-o", so, os.path.join(ROOT, "tools/trace/conv_trace.cu"), "-lcudart"])
lib = ctypes.CDLL(so)
dev = torch.device("cuda", 0)
t = 500
strat = DDPM()
g = torch.Generator().manual_seed(1234)
x0 = torch.randn(32, 2048, 3, generator=g); x0 = x0 / x0.norm(dim=-1, keepdim=True) * math.sqrt(3.0)
ah = float(strat.alpha_hat[t])
x = math.sqrt(ah) * x0 + math.sqrt(1 - ah) * torch.randn(32, 2048, 3, generator=g)
st = torch2sparse(x.to(dev))
geo = spvd_b200.Geometry(st.C.float(), 1e-5, 0.1, strides=(1, 2, 4, 8, 16), p2v_strides=(1,), devox_strides=(1,), n_batch=32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def sort_map(map_t, n):
    kvol, ld = map_t.shape
    ks = torch.arange(kvol, device=dev)
    valid = map_t[:, :n] >= 0
    mask = (valid.long() << ks[:, None]).sum(0)
    perm = torch.argsort(mask, stable=True)
    ms = torch.full((kvol, ld), -1, dtype=torch.int32, device=dev)
    ms[:, :n] = map_t[:, perm]
    pm = torch.zeros(ld, dtype=torch.int64, device=dev); pm[:n] = mask[perm]
    pm = pm.view(-1, 128)
    acc = pm[:, 0].clone()
    for j in range(1, 128):
        acc |= pm[:, j]
    return perm.int().contiguous(), ms.contiguous(), acc.int().contiguous()


def P(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


for s, cin, cout, srt, lite in [(2, 128, 128, False, 1), (2, 128, 128, True, 1), (2, 128, 128, True, 0), (4, 320, 192, False, 1), (4, 320, 192, False, 0),
                                (1, 64, 64, True, 1), (8, 192, 192, False, 1)]:
    lv = geo.level(s); km = lv.kmap3; n = lv.n
    w = torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5
    wp = ops.pack_weights_f16(w)
    xin = torch.randn(n, cin, device=dev).half()
    bias = torch.randn(cout, device=dev)
    out = torch.empty(n, cout, device=dev)
    map_t, mask, perm = km.map_t, km.mask, None
    if srt:
        perm, map_t, mask = sort_map(km.map_t, n)
    ntiles = (n + 127) // 128
    trace = torch.zeros(ntiles * 32, dtype=torch.int64, device=dev)
    for rep in range(2):
        trace.zero_()
        flush.fill_(1)
        torch.cuda.synchronize()
        rc = lib.conv_trace(P(xin), P(wp), P(map_t), map_t.shape[1], P(mask), P(perm), n, 27, cin, cout, P(bias), P(out), lite, P(trace))
        torch.cuda.synchronize()
        assert rc == 0
    T = trace.cpu().numpy().reshape(ntiles, 32).astype(np.float64)
    ref = ops.conv_fwd(xin, w, wp, km.map_t, km.mask, n, bias, ops.CONV_UMMA, mt=3)
    err = float((out - ref).norm() / ref.norm())
    it = T[:, 16]
    wall = (T[:, 11].max() - T[:, 10].min()) / 1e3
    start = (T[:, 10] - T[:, 10].min()) / 1e3
    life = (T[:, 11] - T[:, 10]) / 1e3
    def st_(a):
        return f"mean {a.mean():8.0f} med {np.median(a):8.0f} max {a.max():8.0f}"
    print(f"=== s={s} {cin}->{cout} sorted={srt} lite={lite} tiles={ntiles} err={err:.1e} kernel wall {wall:.1f}us; CTA life us: {st_(life)}; "
          f"CTA start us: med {np.median(start):.1f} max {start.max():.1f}; iters/tile mean {it.mean():.1f} max {it.max():.0f}")
    print("  clk: init(alloc+sync)      ", st_(T[:, 1] - T[:, 0]))
    print("  clk: rows staged+pdl+sync  ", st_(T[:, 2] - T[:, 1]))
    print("  clk: first stage issued    ", st_(T[:, 3] - T[:, 2]))
    print("  clk: producer loop (all)   ", st_(T[:, 4] - T[:, 2]))
    print("  clk: mma first data        ", st_(T[:, 5] - T[:, 2]))
    print("  clk: mma loop              ", st_(T[:, 6] - T[:, 5]))
    print("  clk: epilogue wait (4->7)  ", st_(T[:, 7] - T[:, 4]))
    print("  clk: epilogue body         ", st_(T[:, 8] - T[:, 7]))
    print("  clk: total                 ", st_(T[:, 9] - T[:, 0]))
    nz = it > 0
    print("  clk/iter: producer wait empty", st_(T[nz, 17] / it[nz]), "| mma wait full", st_(T[nz, 18] / it[nz]), "| mma loop/iter", st_((T[nz, 6] - T[nz, 5]) / it[nz]))
