"""Achieved HBM GB/s of the memory-bound kernels of the path (hash build / probe, kernel maps, point<->voxel queries,
scatter-mean, devoxelize, fused glue) at the BASELINE shapes (SPVD, B=32x2048), against the measured copy bandwidth
(MEASURED_PEAKS.json).  Algorithmic bytes = what the op must read + write once (SURVEY 8d / DESIGN section 4), time = CUDA
events around the op on torch's current stream, best of 9 after warm-up, with a 256 MiB L2 flush write before every
repetition.  usage: python tools/hbm_kernels_bench.py [t] > profiles/…txt"""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spvd_b200
from spvd_b200 import ops
from spvd_b200.sampler import torch2sparse, DDPM

t = int(sys.argv[1]) if len(sys.argv) > 1 else 500
dev = torch.device("cuda", 0)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
    src = "MEASURED_PEAKS.json hbm_gbs"
except Exception:
    peak, src = 6400.0, "fallback (B200_PROFILING.md copy bandwidth)"
strat = DDPM()
g = torch.Generator().manual_seed(1234)
x0 = torch.randn(32, 2048, 3, generator=g); x0 = x0 / x0.norm(dim=-1, keepdim=True) * math.sqrt(3.0)
ah = float(strat.alpha_hat[t])
x = math.sqrt(ah) * x0 + math.sqrt(1 - ah) * torch.randn(32, 2048, 3, generator=g)
st = torch2sparse(x.to(dev))
pc = st.C.float().contiguous()
geo = spvd_b200.Geometry(pc, 1e-5, 0.1, strides=(1, 2, 4, 8, 16), p2v_strides=(1, 4, 16), devox_strides=(1, 4, 16), n_batch=32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
NP = pc.shape[0]


def timeit(fn, reps=9):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3)
    return best


rows = []


def timeit_warm(fn, reps=5, inner=8):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            fn()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / inner)
    return best


def report(name, nbytes, fn, note=""):
    us = timeit(fn)
    warm = timeit_warm(fn)
    rows.append((name, nbytes / 1e6, us, nbytes / us / 1e3, nbytes / us / 1e3 / peak, warm, nbytes / warm / 1e3, note))


print(f"# SPVD B=32x2048, t={t}: voxels per level " + str({s: geo.level(s).n for s in (1, 2, 4, 8, 16)}))
print(f"# peak = {peak} GB/s ({src}); cold L2 (256 MiB flush before each repetition)")
for s in (1, 4, 16):
    lv = geo.level(s)
    n = lv.n
    # hash build: read n coords (16 B), write key+val per row (12 B) into a 2*NP-slot table (+ clearing it: 12 B x 2NP)
    tab = ops.HashTable(2 * NP, dev)
    report(f"hash build (stride {s})", n * 28 + 2 * NP * 12, lambda: tab.build(lv.coords_cap, lv.n_dev), "incl. clearing the 2N-slot table")
    # probe: K=1 lookup of every voxel: 16 B coord in, ~12 B slot read, 4 B out
    report(f"hash probe K=1 (stride {s})", n * 32, lambda: lv.table.query(lv.coords, 1) if hasattr(lv.table, "query") else None)
    # 3^3 kernel map: 16 B coord in, 27 probes (12 B each, L2-resident table), 27 x 4 B out
    report(f"kernel map 3^3 (stride {s})", n * (16 + 27 * 4) + 0, lambda: ops.kmap_subm3(lv.coords_cap, lv.n_dev, lv.table),
           "HBM bytes = coords in + 27 x 4 B map out; the 27 probes per voxel hit the L2-resident table")
    for c in (32, 256):
        if s == 1 and c == 256 or s == 16 and c == 32:
            continue
        z = torch.randn(NP, c, device=dev)
        idx = geo.p2v[s]
        report(f"scatter-mean fwd C={c} (points -> stride {s})", NP * (4 * c + 4) + n * 4 * c * 2,
               lambda: ops.scatter_mean_fwd(z, idx, n), "zero-fill + RED adds + divide pass over the voxel rows")
        xv = torch.randn(n, c, device=dev)
        i8, w8 = geo.devox[s]
        report(f"devoxelize fwd C={c} (stride {s} -> points)", NP * 64 + n * 4 * c + NP * 4 * c,
               lambda: ops.devoxelize_fwd(xv, i8, w8), "voxel rows are re-read from L2 (8 corners per point)")
    report(f"point->voxel query (stride {s})", NP * (16 + 4), lambda: ops.point_voxel_query(geo.fcoords, s, lv.table))
    report(f"devox query (stride {s})", NP * (16 + 64), lambda: ops.devox_query(geo.fcoords, s, lv.table))
for n, c in ((geo.level(1).n, 32), (geo.level(2).n, 128), (geo.level(4).n, 192)):
    xx = torch.randn(n, c, device=dev); sc = torch.randn(c, device=dev); sh = torch.randn(c, device=dev)
    report(f"BN+SiLU affine fp32 -> fp32 [{n},{c}]", n * c * 8, lambda: ops.affine_act(xx, sc, sh, 1))
print("# cold = one call after a 256 MiB flush write (the L2 then holds dirty flush lines whose write-back competes with the "
      "kernel, and the wrapper's torch.empty sits inside the timed region: ~5-8 us floor); warm = 8 calls back to back / 8")
print(f"{'kernel':52s} {'MB':>8s} {'cold us':>8s} {'GB/s':>7s} {'of peak':>8s} {'warm us':>8s} {'GB/s':>7s}")
for name, mb, us, gbs, frac, warm, wgbs, note in rows:
    print(f"{name:52s} {mb:8.2f} {us:8.1f} {gbs:7.0f} {frac:8.2f} {warm:8.1f} {wgbs:7.0f}  {note}")
