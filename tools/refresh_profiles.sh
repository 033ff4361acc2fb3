#!/bin/bash
# after tools/gpu_r02_final1.sh: turn what came back in gpurun_out/ into the committed summaries under profiles/
set -e
cd "$(dirname "$0")/.."
python tools/make_traffic_json.py > /dev/null
for k in tpf nii quad752 lat mf demap; do python tools/ncu_key_metrics.py gpurun_out/prof_dram_$k.ncu-rep > /tmp/k_$k.txt 2>&1; done
cp /tmp/k_tpf.txt profiles/r02_tpf_kernel_ncu.txt; cp /tmp/k_nii.txt profiles/r02_nii_kernel_ncu.txt
cp /tmp/k_quad752.txt profiles/r02_quad_kernel_n752_ncu.txt; cp /tmp/k_lat.txt profiles/r02_lat_kernel_ncu.txt; cp /tmp/k_mf.txt profiles/r02_waveform_ncu.txt; cp /tmp/k_demap.txt profiles/r02_demap_ncu.txt
cp gpurun_out/r02_final_gpu_tests.txt gpurun_out/r02_launches.csv gpurun_out/r02_mc_source_perf.txt gpurun_out/r02_latency.txt gpurun_out/r02_long.txt profiles/
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+1; break
kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); mu=hdr.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[start:]:
    if len(r)<=mv: continue
    try: v=float(r[mv].replace(',',''))
    except Exception: continue
    unit=r[mu]
    ms = v/1e6 if unit in ('ns','nsecond') else v/1e3 if unit in ('us','usecond') else v if unit in ('ms','msecond') else v*1e3
    agg[r[kn][:90]][0]+=1; agg[r[kn][:90]][1]+=ms
tot=sum(v[1] for v in agg.values())
out=["# ncu --metrics gpu__time_duration.sum --clock-control none -c 600, python bench.py --steps 2 --warmup 3 --frames 262144 (cold-cache, serialised: compare SHARES)",
     f"# total {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches (the first 600 of the process: headline decode, demapper / mapper / waveform blocks, 16QAM chain, non-parity modes, start of the host pipeline)"]
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]:
    out.append(f"{v[1]:10.2f} ms {v[1]/tot*100:6.2f} %  x{v[0]:4d}  {k}")
open('profiles/r02_launches_summary.txt','w').write("\n".join(out)+"\n")
PY
