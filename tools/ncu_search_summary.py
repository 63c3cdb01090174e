#!/usr/bin/env python3
"""Summarises an `ncu --set full` capture of the search kernel into the JSON bench.py reads for `roofline.traffic`
and `roofline_l1` (profiles/r02_search_ncu.json) and prints a markdown table row.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_search_summary.py raw.csv --frames 75776 --frame-passes 291700 [--out profiles/r02_search_ncu.json]

--frame-passes: refinement passes the captured launch executed (tools/bench_search.py prints it from
mcq_search_stats)."""
import argparse
import csv
import json
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("--frames", type=int, required=True)
    ap.add_argument("--frame-passes", type=float, required=True)
    ap.add_argument("--sms", type=int, default=148)
    ap.add_argument("--out")
    ap.add_argument("--row", type=int, default=0, help="which captured launch (row of the raw page)")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw_csv)))
    hdr, units, vals = rows[0], rows[1], rows[2 + a.row]
    m = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            continue
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0,
                 "ns": 1e-9}.get(u, 1.0)
        m[h] = x * scale

    def get(pat):
        for k, v in m.items():
            if re.search(pat, k):
                return v
        return None
    fp = a.frame_passes
    wf_total = get(r"TriageCompute\.l1tex__data_pipe_lsu_wavefronts\.avg$") * a.sms
    wf_shared = get(r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$")
    wf_lgds = get(r"TriageCompute\.l1tex__data_pipe_lsu_wavefronts_mem_lgds\.avg$") * a.sms
    out = {
        "kernel": m and next((v for h, v in zip(hdr, vals) if h == "Kernel Name"), None),
        "frames_per_launch": a.frames, "frame_passes_per_launch": fp,
        "duration_ms_under_ncu": get(r"^gpu__time_duration\.sum$") * 1e3,
        "dram_bytes_read": get(r"^dram__bytes_read\.sum$"), "dram_bytes_write": get(r"^dram__bytes_write\.sum$"),
        "dram_bytes_per_launch": get(r"^dram__bytes_read\.sum$") + get(r"^dram__bytes_write\.sum$"),
        "l1_wavefronts_per_frame_pass": wf_total / fp,
        "l1_wavefronts_global_per_frame_pass": wf_lgds / fp,
        "l1_wavefronts_shared_per_frame_pass": wf_shared / fp,
        "warp_instructions_per_frame_pass": get(r"^smsp__inst_executed\.sum$") / fp,
        "lsu_data_pipe_pct": get(r"^l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed$"),
        "issue_active_pct": get(r"^sm__issue_active\.avg\.pct_of_peak_sustained_elapsed$"),
        "registers_per_thread": get(r"^launch__registers_per_thread$"),
        "warps_active_per_sm": get(r"^sm__warps_active\.avg\.per_cycle_active$"),
        "l1_hit_rate_global_ld_pct": get(r"^l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate\.pct$"),
        "lts_hit_rate_pct": get(r"^lts__t_sector_hit_rate\.pct$"),
        "stalls_per_issue": {k.split("issue_stalled_")[1].split("_per_issue")[0]: round(v, 3) for k, v in m.items()
                             if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                             and v >= 0.05},
    }
    print(json.dumps(out, indent=1))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    main()
