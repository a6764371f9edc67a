#!/usr/bin/env python
"""per-source-line totals from `ncu -i rep --page source --csv --print-source sass,cuda`:
   tools/ncu_source_lines.py <csv> [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None; out = []
for x in rows:
    if len(x) > 6 and x[0] == "Line No":
        hdr = x; continue
    if hdr and len(x) == len(hdr) and x[0] != "":
        out.append(x)
ci = {n: i for i, n in enumerate(hdr)}
ie = ci["Instructions Executed"]; sm = ci["# Samples"]
tot = sum(int(x[ie]) for x in out); ts = sum(int(x[sm]) for x in out)
print("total warp instructions %d, samples %d" % (tot, ts))
for x in sorted(out, key=lambda x: -int(x[sm]))[:top]:
    print("%5s inst %5.1f%% samp %5.1f%%  %s" % (x[0], 100.0 * int(x[ie]) / tot, 100.0 * int(x[sm]) / max(ts, 1), x[1][:120]))
