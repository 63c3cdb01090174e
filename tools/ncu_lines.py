#!/usr/bin/env python3
"""Per-CUDA-source-line summary of an ncu report's source page.
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_lines.py src.csv <units>
`units` = the number the counts are divided by (e.g. frames x passes)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
F = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[2]
col = {n: hdr.index(n) for n in ("Instructions Executed", "L1 Wavefronts Shared", "L1 Tag Requests Global",
                                 "L2 Theoretical Sectors Global", "# Samples")}


def num(v):
    try:
        return float(v)
    except ValueError:
        return 0.0


tot = [0.0] * 5
for r in rows[3:]:
    if len(r) <= max(col.values()) or r[0] == "":
        continue
    v = [num(r[c]) for c in col.values()]
    if v[0] > 0:
        tot = [a + b for a, b in zip(tot, v)]
        print(f"{r[0]:>4} inst={v[0] / F:8.1f} shwf={v[1] / F:7.1f} gtag={v[2] / F:7.1f} sec={v[3] / F:7.1f} "
              f"smp={int(v[4]):6d} | {r[1][:100]}")
print("total inst=%.1f shwf=%.1f gtag=%.1f sec=%.1f samples=%d" % (tot[0] / F, tot[1] / F, tot[2] / F, tot[3] / F, tot[4]))
