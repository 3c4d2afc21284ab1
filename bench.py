#!/usr/bin/env python
"""bench.py -- SPVD denoise steps/s at B=32 x 2048 points (BASELINE.json metric, configs[1]).

One "step" = one DDPM denoising step of the sampling loop at batch 32 x 2048 points: SPVD forward
(voxelize -> kernel maps -> sparse convs -> devoxelize) + DDPM update + re-sparsification
(utils/schedulers.py:148-170,207-212,249-257).  Open-loop replay:
step j denoises x_t = sqrt(abar_t) x0 + sqrt(1-abar_t) eps for a timestep t spread over 999..0, x0 =
points on a sphere of radius sqrt(3); with random-init weights the closed loop diverges, the replay
keeps the voxel occupancy of a real trajectory.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 is launched by torchrun (one rank per GPU); sampling shards clouds across ranks with no
collective (weak scaling: every rank denoises its own 32 x 2048 batch).
--impl reference times the CPU oracle (oracle/model.py -- the reference's torchsparse path cannot
run without CUDA and torchsparse is not installable here) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SPVD denoise steps/s, B=32x2048 pts"
UNIT = "steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--preset", default="SPVD")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--points", type=int, default=2048)
    ap.add_argument("--downsample-mode", default="spconv", choices=["spconv", "floor"])
    ap.add_argument("--cpu-baseline-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--train-steps", type=int, default=8, help="timed steps of the training sub-record (0 = skip)")
    ap.add_argument("--no-tf32-leg", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------- workload
def _schedule():
    """spvd_b200/schedule.py loaded by path: torch-only coefficient tables, so that the CPU arm never imports the
    spvd_b200 package (which would map libspvd_b200.so into the process)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_spvd_schedule", os.path.join(ROOT, "spvd_b200", "schedule.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload_name(args):
    return (f"{args.preset} DDPM sampling step (forward + update + re-sparsify), open-loop replay over t=999..0, "
            f"B={args.batch}x{args.points} points per GPU, random-init weights, downsample_mode={args.downsample_mode}")


def make_workload(n_steps, B, N, seed):
    """Host tensors: per step the noisy cloud x_t [B,N,3] and its timestep."""
    strat = _schedule().DDPM()
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(B, N, 3, generator=g)
    x0 = x0 / x0.norm(dim=-1, keepdim=True) * math.sqrt(3.0)
    xs, ts = [], []
    for j in range(n_steps):
        t = int(round(999 - j * 999 / max(1, n_steps - 1))) if n_steps > 1 else 999
        ah = float(strat.alpha_hat[t])
        eps = torch.randn(B, N, 3, generator=g)
        xs.append((math.sqrt(ah) * x0 + math.sqrt(1 - ah) * eps).contiguous())
        ts.append(t)
    return strat, xs, ts


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [v.strip() for v in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch.distributed as dist
    import spvd_b200
    from spvd_b200 import ops
    from spvd_b200.sampler import DDPMSparseSchedulerGPU

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N, K, W = args.batch, args.points, args.steps, args.warmup
    torch.manual_seed(0)                     # random-init weights of the named architecture (the module's own initialisers;
    model = spvd_b200.build_model(args.preset, downsample_mode=args.downsample_mode)   # nothing of oracle/ is on this arm)
    model = model.to(dev).eval()
    sched = DDPMSparseSchedulerGPU()
    strat, xs, ts = make_workload(K + W, B, N, seed=1234 + rank)
    shape = (B, N, 3)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_device(x_pts, t, i):
        st = sched.torch2sparse(x_pts, shape)
        tb = torch.full((B,), t, device=dev, dtype=torch.long)
        eps = model((st, tb))
        return sched.update_rule(st, eps, t, i, shape, dev)        # DDPM update + re-sparsify (fused kernels)

    # ---- device-resident leg (value)
    xs_dev = [x.to(dev) for x in xs]
    with torch.no_grad():
        for j in range(W):
            step_device(xs_dev[j], ts[j], j)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = ClockSampler(local) if rank == 0 else None
        ops.stats.reset(False)
        ev = []
        for j in range(W, W + K):
            if flush is not None:
                flush.fill_(j & 0xFF)                  # L2 flush: 256 MiB write between timed steps
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_device(xs_dev[j], ts[j], j)
            b.record()
            ev.append((a, b))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clk = clocks.stop() if clocks else None
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = ops.stats.launches
    ops.stats.reset(False)
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())

    # ---- roofline of the dominant kernel (tcgen05 sparse conv): an instrumented replay of the timed steps,
    #      CUDA events recorded around every convolution launch on the launching stream (inside the C executor)
    roof = None
    if rank == 0:
        sink = []
        model.profile_sink = sink
        with torch.no_grad():
            for j in range(W, W + K):
                if flush is not None:
                    flush.fill_(j & 0xFF)
                step_device(xs_dev[j], ts[j], j)
        torch.cuda.synchronize()
        model.profile_sink = None
        flops = ms = 0.0
        n = 0
        table = {}
        strides = model._plan["strides"]
        for evs, meta, pairs, sizes in sink:
            for ci, (kvol, cin, cout, rows_out, level, map_kind, kind, cin2) in enumerate(meta):
                dt = evs[2 * ci].elapsed_time(evs[2 * ci + 1])
                n_rows = B * N if rows_out == -1 else (B if rows_out == -2 else sizes[rows_out])
                P = n_rows if map_kind == 0 else pairs[(level, map_kind)]
                fl = 2.0 * P * cin * cout + 2.0 * n_rows * cin2 * cout       # + the 1x1 shortcut folded into the K loop
                e = table.setdefault((kind, kvol, cin, cout, strides[level] if level >= 0 else rows_out), [0, 0.0, 0.0])
                e[0] += 1
                e[1] += dt
                e[2] += fl
                if kind in ("umma", "sp"):
                    flops += fl
                    ms += dt
                    n += 1
        if os.environ.get("SPVD_BENCH_VERBOSE"):
            for key in sorted(table, key=lambda k: -table[k][1]):
                c, t_ms, fl = table[key]
                print(f"[conv] {key} launches={c} total_ms={t_ms:.3f} avg_us={t_ms / c * 1e3:.1f} "
                      f"alg_TFLOPs={fl / (t_ms * 1e-3) / 1e12:.1f}", file=sys.stderr)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        traffic, traffic_src = None, None
        try:        # mean DRAM bytes per launch of the tcgen05 conv kernels only, from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "conv_umma_traffic.json")))
            traffic = tj.get("dram_bytes_per_umma_launch", tj.get("dram_bytes_per_launch"))
            traffic_src = tj.get("source", "profiles/conv_umma_traffic.json")
        except Exception:
            pass
        ach = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        f16 = model.executor_operands == "f16"
        roof = {"bound": "tensor", "kernel": "k_conv_fwd_umma (tcgen05 kind::%s)" % ("f16" if f16 else "tf32"), "achieved": round(ach, 2),
                "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PF sustained",
                "note": ("f16 operands (same tensor-core rate as the bf16 peak), fp32 accumulate; " if f16 else
                         "TF32 operands (kind::tf32 runs at half the bf16 rate, so frac <= 0.5 by construction); ") + "durations from "
                        "CUDA events around each launch in an instrumented replay of the timed steps",
                "launches": n, "avg_launch_us": round(ms * 1e3 / max(n, 1), 2),
                "algorithmic_gflop_per_step": round(flops / K / 1e9, 1),
                "conv_share_of_step": round(ms / total_ms, 3) if total_ms else None}

    # ---- second roofline entry: the worst HBM-bound kernel of the path (devoxelize, C = 256 from the coarsest level),
    #      algorithmic bytes (Np*64 + Nv*4C + Np*4C) / CUDA-event time, cold L2
    roof_hbm = None
    if rank == 0:
        try:
            st = sched.torch2sparse(xs_dev[W], shape)
            geo = model.plan_geometry(st.coords_float, B, N).finish()
            s_top = max(geo.devox)
            idx8, w8 = geo.devox[s_top]
            nv, C = geo.level(s_top).n, 256
            xv = torch.randn(nv, C, device=dev)
            ops.devoxelize_fwd(xv, idx8, w8)
            dts = []
            for _ in range(10):
                if flush is not None:
                    flush.fill_(3)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ops.devoxelize_fwd(xv, idx8, w8)
                b.record()
                torch.cuda.synchronize()
                dts.append(a.elapsed_time(b))
            nbytes = B * N * 64 + nv * 4 * C + B * N * 4 * C
            hbm_peak = float(peaks.get("hbm_gbs", 6400.0))
            gbs = nbytes / (statistics.median(dts) * 1e-3) / 1e9
            roof_hbm = {"bound": "hbm", "kernel": f"k_devox_fwd4 (C={C}, stride {s_top} -> points)", "achieved": round(gbs, 1),
                        "peak": hbm_peak, "unit": "GB/s", "frac": round(gbs / hbm_peak, 4), "algorithmic_bytes": nbytes,
                        "note": "includes the wrapper's output allocation; median of 10 cold-L2 launches"}
        except Exception as e:          # a measurement aid must never take the bench line down
            roof_hbm = {"error": str(e)[:200]}

    # ---- the same device-resident leg with TF32 operands (the reference's precision class), for comparison
    value_tf32 = None
    if not args.no_tf32_leg:
        model.executor_operands = "tf32"
        with torch.no_grad():
            for j in range(W):
                step_device(xs_dev[j], ts[j], j)
            torch.cuda.synchronize()
            ev2 = []
            for j in range(W, W + K):
                if flush is not None:
                    flush.fill_(j & 0xFF)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                step_device(xs_dev[j], ts[j], j)
                b.record()
                ev2.append((a, b))
            torch.cuda.synchronize()
        t32 = sum(a.elapsed_time(b) for a, b in ev2)
        if world > 1:
            tt = torch.tensor([t32], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t32 = float(tt.item())
        value_tf32 = round(world * K / (t32 * 1e-3), 3)
        model.executor_operands = "f16"

    # ---- end-to-end leg through the public API with host buffers
    pin_in = [x.pin_memory() for x in xs]
    pin_out = torch.empty(shape, dtype=torch.float32).pin_memory()
    with torch.no_grad():
        for j in range(min(W, 2)):
            st = sched.torch2sparse(pin_in[j].to(dev, non_blocking=True), shape)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for j in range(W, W + K):
            x = pin_in[j].to(dev, non_blocking=True)                      # H2D of this step's input
            st = sched.torch2sparse(x, shape)
            tb = torch.full((B,), ts[j], device=dev, dtype=torch.long)
            eps = model((st, tb))
            new = sched.update_rule(st, eps, ts[j], j, shape, dev).F
            pin_out.copy_(new.view(shape), non_blocking=True)             # D2H of the step's result
            torch.cuda.current_stream().synchronize()                     # the caller owns x_{t-1} on the host
        e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())

    train = run_train_leg(args, dev, world, rank, xs_dev, flush) if args.train_steps > 0 else None

    out = None
    if rank == 0:
        nbytes = B * N * 3 * 4
        out = {"metric": METRIC, "value": round(world * K / (total_ms * 1e-3), 3), "unit": UNIT, "n_gpus": world,
               "steps": K, "warmup": W, "ms_per_step": round(total_ms / K, 4), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": ("f32 activations / accumulate, f16 tensor-core operands (11-bit significand, as TF32)"
                         if model.executor_operands == "f16" else "f32 (tf32 tensor-core operands, f32 accumulate)"),
               "data": "synthetic",
               "config": {"workload": workload_name(args),
                          "preset": args.preset, "batch_per_gpu": B, "points": N, "parallelism": f"dp{world} (clouds sharded, no collective)",
                          "l2": "256 MiB flush write between timed steps" if flush is not None else "no flush (per-step working set > L2)"},
               "e2e": {"value": round(world * K / e2e_s, 3), "unit": UNIT, "h2d_bytes_per_step": nbytes,
                       "d2h_bytes_per_step": nbytes},
               "gpu_launches": launches, "clocks": clk, "roofline": roof, "roofline_hbm": roof_hbm,
               "value_tf32_operands": value_tf32, "train": train}
    return out, rank, world, dict(model=model, dev=dev, sched=sched)


def run_train_leg(args, dev, world, rank, xs_dev, flush):
    """BASELINE config 3 as a sub-record: one optimisation step = forward + MSE + backward + NCCL gradient all-reduce
    (bucketed, launched from backward hooks) + clip_grad_norm_(10) + Adam, B per GPU as in the sampling leg
    (train_generation.py:98-112, utils/callbacks.py:58-63).  Device-timed, max over ranks."""
    import torch.distributed as dist
    import spvd_b200
    from spvd_b200.parallel import GradientAllReduce, train_step
    from spvd_b200.optim import FlatAdam
    from spvd_b200.sampler import torch2sparse
    B, N, K, W = args.batch, args.points, args.train_steps, 6     # (the first NCCL all-reduces of a new buffer set up connections,
                                                                  #  both geometry buffer sets capture their planner graphs)
    torch.manual_seed(0)
    model = spvd_b200.build_model(args.preset, downsample_mode=args.downsample_mode).to(dev).train()
    red = GradientAllReduce(model.parameters())
    opt = FlatAdam(red, lr=1e-4)          # flat-buffer Adam + clip + average on the library's kernels
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    ev, ar_ms = [], 0.0

    def make_batch(j):            # what the data loader hands over (train_generation.py:98-103): noisy clouds, t, target noise
        st = torch2sparse(xs_dev[j % len(xs_dev)])
        return st, torch.randint(0, 1000, (B,), device=dev, generator=g), torch.randn(st.F.shape, device=dev, generator=g)

    def run_steps(n, timed_from):
        """batch j + 1 is produced, and its geometry planned on a side stream (model.prefetch_geometry), while step j runs"""
        out = []
        batch = make_batch(0)
        geo = model.prefetch_geometry(batch[0], B)
        for j in range(n):
            if flush is not None:
                flush.fill_(j & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if j == timed_from:
                import gc
                gc.collect()          # (a full collection of Python's cyclic GC inside the 8 timed steps showed up as one 80 ms step)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
            a.record()
            nxt = make_batch(j + 1) if j + 1 < n else None
            geo_next = model.prefetch_geometry(nxt[0], B) if nxt is not None else None
            train_step(model, opt, batch, red, geometry=geo)
            b.record()
            if j >= timed_from:
                out.append((a, b))
            batch, geo = nxt, geo_next
        return out

    ev = run_steps(W + K, W)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total = sum(a.elapsed_time(b) for a, b in ev)
    if os.environ.get("SPVD_BENCH_VERBOSE"):
        print(f"[train rank {rank}] per-step ms: " + " ".join(f"{a.elapsed_time(b):.1f}" for a, b in ev), file=sys.stderr)
    # exposed all-reduce time: the same steps with the collective switched off (identical kernels otherwise)
    exposed = None
    if world > 1:
        red_group_active = red._active
        red._active = lambda: False
        ev2 = run_steps(K, 0)
        torch.cuda.synchronize()
        red._active = red_group_active
        local = sum(a.elapsed_time(b) for a, b in ev2)
        tt = torch.tensor([total, local], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total, local = float(tt[0]), float(tt[1])
        exposed = round(max(0.0, (total - local) / K), 3)
    if rank != 0:
        return None
    return {"metric": "SPVD training steps/s (fwd + bwd + gradient all-reduce + clip + Adam), B=%dx%d per GPU" % (B, N),
            "steps_per_s": round(K / (total * 1e-3), 3), "clouds_per_s": round(world * B * K / (total * 1e-3), 1),
            "ms_per_step": round(total / K, 3), "ms_per_step_median": round(sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2], 3),
            "steps": K, "warmup": W, "n_gpus": world,
            "pipeline": "batch j+1 (noise, t, re-sparsify) is produced and its geometry planned on a side stream while step j runs",
            "allreduce": {"bytes": int(red.flat.numel() * 4), "buckets": len(red.buckets), "exposed_ms_per_step": exposed,
                          "how": "step time with the NCCL all-reduce minus step time without it (same kernels), max over ranks"},
            "dtype": "f32 activations, tf32 tensor-core operands (forward, dgrad, wgrad), f32 accumulate"}


# -------------------------------------------------------------------------------- CPU oracle arm
def cpu_steps(args, n_steps, warmup, seed=1234, sd=None, keep=None):
    """The reference algorithm restated on the CPU (oracle/), timed on the host cores.  ``sd``: weights to use (default: the
    oracle's own seeded initialisation); ``keep``: list that receives (coords, feats, t, eps) of the first timed step."""
    from oracle import model as om
    B, N = args.batch, args.points
    strat, xs, ts = make_workload(n_steps + warmup, B, N, seed)
    sd = om.make_state_dict(args.preset, seed=0) if sd is None else sd
    times = []
    with torch.no_grad():
        for j in range(n_steps + warmup):
            t0 = time.perf_counter()
            coords, feats, _ = om.torch2sparse(xs[j])
            eps = om.forward(sd, args.preset, coords, feats, torch.full((B,), ts[j], dtype=torch.long),
                             downsample_mode=args.downsample_mode)
            new = strat.update(feats, eps, ts[j], j, lambda s: torch.randn(s))
            if new.shape[0] == B * N:
                om.torch2sparse(new.view(B, N, 3))
            dt = time.perf_counter() - t0
            if j >= warmup:
                times.append(dt)
                if keep is not None and not keep:
                    keep.append((coords, feats, ts[j], eps))
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return None, rank, world
    K, W = args.steps, args.warmup
    try:        # torchrun exports OMP_NUM_THREADS=1; this arm is alone on the host: use every core it may run on
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    times = cpu_steps(args, K, W)
    total = sum(times)
    v = round(K / total, 4)
    cores = torch.get_num_threads()
    sample = f"{K} full-size steps ({args.preset}, B={args.batch}x{args.points}) of the same open-loop replay, torch-CPU"
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": round(total / K * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(args),
                      "preset": args.preset, "batch_per_gpu": args.batch, "points": args.points,
                      "note": "reference's torchsparse path cannot run on CPU (GPUHashTable, models/sparse_utils.py:43) "
                              "and torchsparse is not installed: this arm is the CPU oracle port (oracle/model.py)"},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    return out, rank, world


def main():
    args = parse()
    if args.impl == "reference":
        out, rank, world = run_reference(args)
    else:
        out, rank, world, ctx = run_b200(args)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            n = max(1, args.cpu_baseline_steps)
            # the CPU leg runs the GPU arm's own weights, so its first step doubles as the parity check of what was timed
            model, dev = ctx["model"], ctx["dev"]
            sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
            keep = []
            times = cpu_steps(args, n, 0, sd=sd, keep=keep)
            out["cpu_baseline"] = {"value": round(n / sum(times), 4), "unit": UNIT, "cores": torch.get_num_threads(),
                                   "kind": "port",
                                   "sample": f"{n} full-size steps (t spread over 999..0) of the same workload on the "
                                             f"CPU oracle (oracle/model.py), os.cpu_count()={os.cpu_count()}"}
            coords, feats, t0, eps_cpu = keep[0]
            import spvd_b200
            with torch.no_grad():
                x = spvd_b200.SparseTensor(feats.to(dev), coords.to(dev))
                x.points_per_cloud = args.points
                eps_gpu = model((x, torch.full((args.batch,), t0, device=dev, dtype=torch.long))).cpu()
            err = float((eps_gpu - eps_cpu).norm() / eps_cpu.norm())
            out["parity_rel_err"] = round(err, 6)
            out["parity"] = {"against": "CPU oracle forward on the first timed step's input, same weights", "tolerance": 1e-3,
                             "ok": err < 1e-3}
            if err >= 1e-3:
                print(f"bench.py: PARITY FAILURE rel err {err:.3e} >= 1e-3", file=sys.stderr)
        elif rank == 0:
            out["cpu_baseline"] = None
    if rank == 0 and out is not None:
        print(json.dumps(out), flush=True)
    if world > 1 and args.impl != "reference":
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
