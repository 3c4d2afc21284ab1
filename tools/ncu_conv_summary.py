"""profiles/conv_umma_traffic.json from an `ncu --page raw --csv` export of the conv launches of one step.
usage: python tools/ncu_conv_summary.py gpurun_out/conv_r01g_raw.csv profiles/r01g_conv_ncu_full_raw.csv"""
import csv, json, sys

src, name = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
H = {h: i for i, h in enumerate(hdr)}
mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def g(r, k):
    try:
        return float(r[H[k]].replace(',', ''))
    except ValueError:
        return float('nan')


def unit(k):
    return units[H[k]]


TP = 'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'
out, tot_t, tot_b = [], 0.0, 0.0
for i, r in enumerate(data):
    du, dur = unit('gpu__time_duration.sum'), g(r, 'gpu__time_duration.sum')
    dur_us = {'ns': dur / 1e3, 'us': dur, 'ms': dur * 1e3}.get(du, dur)
    rd = g(r, 'dram__bytes_read.sum') * mult.get(unit('dram__bytes_read.sum'), 1)
    wr = g(r, 'dram__bytes_write.sum') * mult.get(unit('dram__bytes_write.sum'), 1)
    sm = g(r, 'launch__shared_mem_per_block_dynamic') * mult.get(unit('launch__shared_mem_per_block_dynamic'), 1) / 1024
    out.append(dict(id=i, kernel=r[H['Kernel Name']][:70], grid=r[H['Grid Size']], duration_us=round(dur_us, 2),
                    dram_read_MB=round(rd / 1e6, 3), dram_write_MB=round(wr / 1e6, 3), tensor_pipe_active_pct=round(g(r, TP), 2),
                    lts_throughput_pct=round(g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'), 2),
                    dram_throughput_pct=round(g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), 2),
                    regs=int(g(r, 'launch__registers_per_thread')), smem_dyn_KB=round(sm, 1)))
    tot_t += dur_us
    tot_b += rd + wr
tc = [o for o in out if any(k in o["kernel"] for k in ("k_conv_fwd_umma", "k_conv_pairs_persistent", "k_conv_sp_f16"))]
tc_bytes = sum((o["dram_read_MB"] + o["dram_write_MB"]) * 1e6 for o in tc)
date = sys.argv[3] if len(sys.argv) > 3 else "n/a"
json.dump({"source": f"{name} (ncu --set full --clock-control none --import-source on, all conv launches of one SPVD denoising step "
                     f"B=32x2048, t=997 geometry, cold caches, captured {date}; f16-operand kernels: one-tile 2-CTA/SM, persistent "
                     "pair-major, sorted-tile persistent with fused epilogue, split-K)",
           "dram_bytes_per_launch": int(tot_b / len(data)), "launches": len(data), "total_duration_us": round(tot_t, 1),
           "dram_bytes_per_umma_launch": int(tc_bytes / max(len(tc), 1)), "umma_launches": len(tc),
           "umma_total_duration_us": round(sum(o["duration_us"] for o in tc), 1),
           "per_launch": out}, open('profiles/conv_umma_traffic.json', 'w'), indent=1)
print(len(data), "launches,", round(tot_t, 1), "us,", int(tot_b / len(data)), "DRAM bytes per launch;", len(tc), "tcgen05 launches,",
      int(tc_bytes / max(len(tc), 1)), "DRAM bytes per tcgen05 launch")
