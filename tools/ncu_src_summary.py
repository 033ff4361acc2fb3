#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: stall mix and hottest SASS regions."""
import csv, sys
path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else None
hi = int(sys.argv[3]) if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data)
keys = ['stall_barrier','stall_long_sb','stall_short_sb','stall_wait','stall_no_inst','stall_not_selected','stall_selected','stall_math','stall_mio','stall_branch_resolving','stall_dispatch','stall_lg','stall_membar']
if lo is None:
    print("total samples", tot, "instructions", len(data), "executed", sum(f(r,'Instructions Executed') for r in data)/1e6, "M")
    for k in keys:
        print(f"  {k:26s} {sum(f(r,k) for r in data)/tot*100:6.2f}%")
    step = 200
    for i in range(0, len(data), step):
        ch = data[i:i+step]
        s = sum(f(r,'# Samples') for r in ch); ex = sum(f(r,'Instructions Executed') for r in ch)
        if s/tot < 0.002: continue
        mix = {k: sum(f(r,k) for r in ch) for k in keys}
        top = sorted(mix.items(), key=lambda kv: -kv[1])[:3]
        print(f"{i:5d} {s/tot*100:5.1f}% smp {ex/1e6:8.1f}M inst  " + " ".join(f"{k[6:]}={v/s*100:.0f}%" for k,v in top) + f"   | {ch[0][1].strip()[:40]}")
else:
    for i in range(lo, hi):
        r = data[i]
        st = " ".join(f"{k[6:10]}={r[ix[k]]}" for k in keys if r[ix[k]] not in ('0',''))
        print(f"{i:5d} {r[1].strip()[:60]:60s} smp={r[ix['# Samples']]:>6s} ex={f(r,'Instructions Executed')/1e6:7.2f}M {st}")
