#!/usr/bin/env python3
"""Builds profiles/r02_traffic.json from the `ncu --set full` captures that tools/gpu_dram.sh brings back in
gpurun_out/ (run here, no GPU needed: reads the reports' raw pages).  The file is stamped with the sha256 of
modulations_b200/libb200dvb.so AS IT IS NOW; tools/gpu_dram.sh records the hash of the library it profiled and
this script refuses to proceed when the two differ.  bench.py reports `roofline.traffic` from this file only while
the hash still matches the library it runs."""
import csv, hashlib, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            try:
                d[h] = float(v) * UNIT.get(u, 1.0)
            except ValueError:
                d[h] = v
        out.append(d)
    return out


def main():
    sha = hashlib.sha256(open(os.path.join(ROOT, "modulations_b200", "libb200dvb.so"), "rb").read()).hexdigest()
    prof_sha = open(os.path.join(OUT, "dram_lib_sha256.txt")).read().split()[0]
    if prof_sha != sha:
        sys.exit(f"the captures were taken on library {prof_sha[:12]}, the tree now holds {sha[:12]}: re-run tools/gpu_dram.sh")
    sys.path.insert(0, ROOT)
    from modulations_b200.build import source_sha256
    res = {"lib_sha256": sha, "src_sha256": source_sha256(), "how": "tools/gpu_dram.sh (ncu --set full --clock-control none, one launch each) + tools/make_traffic_json.py"}
    for key, rep, frames in (("tpf_kernel", "prof_dram_tpf.ncu-rep", 37888), ("nii_kernel", "prof_dram_nii.ncu-rep", 37888),
                             ("quad_kernel_n752", "prof_dram_quad752.ncu-rep", 4736)):
        p = os.path.join(OUT, rep)
        if not os.path.exists(p):
            continue
        d = raw(p)[0]
        res[key] = {
            "capture": f"profiles/r02_{key}_ncu.txt", "frames": frames, "kernel": d["Kernel Name"],
            "dram_bytes_per_frame": (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / frames,
            "dram_read_bytes_per_frame": d["dram__bytes_read.sum"] / frames,
            "dram_write_bytes_per_frame": d["dram__bytes_write.sum"] / frames,
            "duration_ms": d["gpu__time_duration.sum"] * (1.0 if d["gpu__time_duration.sum"] > 1e-3 else 1e3),
            "smsp__issue_active_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "pipe_alu_pct": d.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "pipe_fma_pct": d.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "lsu_wavefronts_pct": d.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "l2_hit_pct": d.get("lts__t_sector_hit_rate.pct"),
            "registers": d.get("launch__registers_per_thread"),
            "warp_instructions_per_frame": d.get("smsp__inst_executed.sum", 0) / frames,
        }
    p = os.path.join(OUT, "prof_dram_demap.ncu-rep")
    if os.path.exists(p):
        res["demap"] = {}
        nsym = 1 << 27
        for d, name in zip(raw(p), ("BPSK", "QPSK", "8PSK", "16QAM", "64QAM", "256QAM")):
            res["demap"][name] = {"kernel": d["Kernel Name"], "symbols": nsym,
                                  "dram_bytes_per_symbol": (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / nsym,
                                  "dram_throughput_pct": d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
    p = os.path.join(OUT, "prof_dram_mf.ncu-rep")
    if os.path.exists(p):
        for d in raw(p):
            k = "matched_filter" if "matched" in d["Kernel Name"] else "pulse_shape"
            res[k] = {"kernel": d["Kernel Name"], "dram_bytes": d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"],
                      "lsu_wavefronts_pct": d.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                      "dram_throughput_pct": d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
    json.dump(res, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
