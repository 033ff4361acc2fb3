#!/usr/bin/env python3
"""Key metrics of every kernel in an .ncu-rep (raw page) + the stall mix of its source page: the text summaries
committed under profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second"]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print(f"== {d['Kernel Name']}")
    for k in KEYS:
        if k in d:
            print(f"   {k:72s} {d[k]:>18s} {u[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
if len(rows) > 3:
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    def f(r, k):
        try: return float(r[ix[k]])
        except Exception: return 0.0
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    tot = sum(f(r, '# Samples') for r in data) or 1.0
    print(f"   warp-state samples {tot:.0f} over {len(data)} SASS instructions (first kernel of the report):")
    for k in [h for h in hdr if h.startswith("stall_")]:
        v = sum(f(r, k) for r in data) / tot * 100
        if v >= 0.5: print(f"      {k:28s} {v:6.2f} %")
