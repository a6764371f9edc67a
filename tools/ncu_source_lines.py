#!/usr/bin/env python
"""per-source-line totals from `ncu -i rep --page source --csv --print-source sass,cuda`:
   tools/ncu_source_lines.py <csv> [top] [function substring]"""
import csv, sys, os
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
fn = fl = None; hdr = None
acc = defaultdict(lambda: defaultdict(lambda: [0, 0, ""]))
for x in rows:
    if len(x) >= 2 and x[0] == "File Path": fl = os.path.basename(x[1]); continue
    if len(x) >= 2 and x[0] == "Function Name": fn = x[1].split("(")[0]; continue
    if len(x) > 6 and x[0] == "Line No":
        hdr = x; ci = {n: i for i, n in enumerate(hdr)}; ie = ci["Instructions Executed"]; sm = ci["# Samples"]; continue
    if hdr and len(x) == len(hdr) and x[0] != "":
        a = acc[fn][(fl, x[0])]
        a[0] += int(x[ie]); a[1] += int(x[sm]); a[2] = x[1]
for f, d in acc.items():
    if want not in f: continue
    tot = sum(v[0] for v in d.values()); ts = sum(v[1] for v in d.values())
    print("== %s: warp instructions %d, samples %d" % (f, tot, ts))
    for (fl, ln), v in sorted(d.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%14s:%-5s inst %5.1f%% samp %5.1f%%  %s" % (fl[:14], ln, 100.0 * v[0] / max(tot, 1), 100.0 * v[1] / max(ts, 1), v[2][:110]))
