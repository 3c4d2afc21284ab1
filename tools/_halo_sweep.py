"""Experiment + correctness: the halo convolution of tools/trace/conv_halo.cu (built here with nvcc, not part of the
product library) against the production tcgen05 kernel on the SPVD layer shapes.
usage: python tools/_halo_sweep.py [t] [trace]"""
import ctypes, os, subprocess, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spvd_b200
from spvd_b200 import ops
from spvd_b200.sampler import torch2sparse, DDPM

t = int(sys.argv[1]) if len(sys.argv) > 1 else 500
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = "/tmp/libconvhalo.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                       *(["-DSPVD_HALO_TRACE"] if "trace" in sys.argv[2:] else []),
                       "-o", so, os.path.join(ROOT, "tools/trace/conv_halo.cu"), "-lcudart"])
hl = ctypes.CDLL(so)
P = lambda x: ctypes.c_void_p(x.data_ptr()) if x is not None else None
S = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def kmap_halo(map_t, n, kvol, mt=2):
    ld, dv = map_t.shape[1], map_t.device
    cap = hl.spvd_halo_cap()
    groups = ((int(n) + 127) // 128 + mt - 1) // mt
    rows = torch.empty((groups, cap), dtype=torch.int32, device=dv)
    cnt = torch.empty(groups, dtype=torch.int32, device=dv)
    lidx = torch.empty((groups, kvol, mt * 128), dtype=torch.int16, device=dv)
    mx = torch.empty(1, dtype=torch.int32, device=dv)
    assert hl.spvd_kmap_halo(P(map_t), ld, int(n), None, kvol, mt, P(rows), P(cnt), P(lidx), P(mx), S()) == 0
    return rows, cnt, lidx, mx


def conv_halo(x16, wp, plan, mt, tile_mask, n_out, kvol, cin, cout, bias=None, residual=None):
    rows, cnt, lidx, _ = plan
    out = torch.empty((n_out, cout), dtype=torch.float32, device=x16.device)
    assert hl.spvd_conv_halo_f16(P(x16), P(wp), P(tile_mask), P(rows), P(cnt), P(lidx), mt, int(n_out), kvol, cin, cout, P(bias),
                                 P(residual), P(out), S()) == 0
    return out

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
    return min(ts)


for s, cin, cout in [(1, 64, 64), (2, 64, 64), (2, 128, 128), (2, 192, 128), (4, 128, 128), (4, 320, 192), (4, 192, 192),
                     (8, 192, 192), (8, 384, 192), (16, 448, 256), (16, 256, 256)]:
    lv = geo.level(s); km = lv.kmap3; n = lv.n
    w = torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5
    wp = ops.pack_weights_f16(w)
    xin = torch.randn(n, cin, device=dev).half()
    bias = torch.randn(cout, device=dev)
    res = torch.randn(n, cout, device=dev)
    ref = ops.conv_fwd(xin, w, wp, km.map_t, km.mask, n, bias, ops.CONV_UMMA, mt=3, residual=res)
    old = timeit(lambda: ops.conv_fwd(xin, w, wp, km.map_t, km.mask, n, bias, ops.CONV_UMMA, residual=res))
    line = f"s={s:2d} n={n:6d} {cin:3d}->{cout:3d} | old-auto:{old:6.1f}"
    for mt in (2,):
        plan = kmap_halo(km.map_t, n, 27, mt)
        mx = int(plan[3])
        pl_us = timeit(lambda: kmap_halo(km.map_t, n, 27, mt))
        if mx > hl.spvd_halo_cap():
            line += f" halo/mt{mt}: overflow ({mx})"
            continue
        fn = lambda: conv_halo(xin, wp, plan, mt, km.mask, n, 27, cin, cout, bias, res)
        errs = []
        for _ in range(3):
            out = fn()
            errs.append(float((out - ref).norm() / ref.norm()))
        if "trace" in sys.argv[2:]:
            fn(); torch.cuda.synchronize()
            buf = (ctypes.c_ulonglong * 32)()
            hl.spvd_halo_trace(buf)
            for b, nm in ((0, "slicer[other halo wait_empty slice arrive wait_acc]"), (8, "mma[other wait_w wait_a fences issue commits syncwarp]"), (16, "wload[other wait_wempty]")):
                nt = max(1, buf[24 + b // 8])
                print("   ", nm, "n=", buf[24 + b // 8], " ".join(f"{buf[b + i] / nt:.0f}" for i in range(8)))
        line += f" halo/mt{mt}:{timeit(fn):6.1f} (max halo {mx}, plan {pl_us:.0f}us" + ("" if max(errs) < 1e-5 else f", ERR {max(errs):.1e}") + ")"
    print(line, flush=True)
