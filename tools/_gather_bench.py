"""Micro-benchmark of the row gather (tools/trace/gather_bench.cu): bytes per cycle and SM when warps gather 128-byte
pieces of random rows of an L2-resident f16 matrix into shared memory.  usage: python tools/_gather_bench.py"""
import ctypes, os, subprocess, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = "/tmp/libgatherbench.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                       "-o", so, os.path.join(ROOT, "tools/trace/gather_bench.cu"), "-lcudart"])
lib = ctypes.CDLL(so)
lib.gather_bench.restype = ctypes.c_float
dev = torch.device("cuda", 0)
P = lambda t: ctypes.c_void_p(t.data_ptr())
n_rows, cin = 28508, 192                       # the stride-4 level of the bench workload, 192 channels f16: 10.9 MB
mat = torch.randn(n_rows, cin, device=dev).half()
g = torch.Generator(device=dev).manual_seed(1)
rows_rand = torch.randint(0, n_rows, (1 << 20,), device=dev, generator=g, dtype=torch.int32)
# neighbour-like: runs of consecutive rows (what a kernel-map row looks like at a dense level)
base = torch.randint(0, n_rows - 160, (1 << 13,), device=dev, generator=g)
rows_run = (base[:, None] + torch.arange(128, device=dev)[None, :]).reshape(-1).int().contiguous()
clocks = torch.zeros(148 * 8 * 32, dtype=torch.int64, device=dev)
sink = torch.zeros(1, dtype=torch.int32, device=dev)
sm_clock = 1.965e9
print(f"matrix {n_rows} x {cin} f16 ({mat.numel() * 2 / 1e6:.1f} MB), 148 SMs; B/clk/SM = bytes gathered / (kernel time x {sm_clock / 1e9:.3f} GHz) / 148")
for name, rows in (("random rows", None), ("runs of 32 consecutive rows", rows_run)):
    print(f"--- {name}")
    for mode, mname in ((0, "cp.async.cg"), (1, "cp.async.ca"), (2, "ld.global+st.shared")):
        for cps, warps in ((1, 4), (2, 4), (1, 8), (2, 8), (1, 16), (2, 16), (1, 32)):
            line = f"{mname:20s} {cps} CTA/SM x {warps:2d} warps:"
            for depth in ((1, 2, 4, 8) if mode < 2 else (1,)):
                if warps * depth * 4096 * cps > 200 * 1024:
                    continue
                iters = 400
                ms = lib.gather_bench(P(mat), cin * 2, P(rows) if rows is not None else None, n_rows, iters, 148 * cps, warps, mode, depth, P(clocks), P(sink))
                if ms <= 0:
                    line += f"  d{depth}: err"
                    continue
                total = 148 * cps * warps * iters * 4096
                line += f"  d{depth}: {total / (ms * 1e-3 * sm_clock) / 148:5.1f}"
            print(line, flush=True)
