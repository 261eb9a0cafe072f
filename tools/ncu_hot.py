#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv` export (gzipped): SASS instructions ranked by warp-stall samples, with the
dominant stall reason and average active threads."""
import csv
import gzip
import sys

rows = list(csv.reader(gzip.open(sys.argv[1], "rt")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                body.append(rows[j])
            j += 1
        col = {h: k for k, h in enumerate(hdr)}
        s_all = col["Warp Stall Sampling (All Samples)"]
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        total = sum(int(r[s_all] or 0) for r in body)
        print("==", name[:120], "| samples", total, "| instructions", len(body))
        agg = {}
        for h in stalls:
            agg[h] = sum(int(r[col[h]] or 0) for r in body)
        print("   stall mix:", ", ".join(f"{h[6:]} {100 * v / max(total, 1):.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        for idx, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][s_all] or 0))[:top]:
            dom = max(stalls, key=lambda h: int(r[col[h]] or 0))
            print(f"   {100 * int(r[s_all] or 0) / max(total, 1):5.1f}%  #{idx:5d} thr {r[col['Avg. Threads Executed']]:>5}  exec {r[col['Instructions Executed']]:>10}  {dom[6:]:10s} {r[col['Source']].strip()[:90]}")
        i = j
    else:
        i += 1
