"""tools/sass_hot.py REPORT.ncu-rep KERNEL [top]: hot spots of one kernel from an `ncu --set full --import-source on`
report: per-SASS-instruction stall samples / executions / active threads (ncu --page source --print-source sass)."""
import csv, subprocess, sys, collections, io
rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the first kernel block only
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = []; blocks.append(cur); continue
    if cur is not None: cur.append(r)
h, data = blocks[0][0], [r for r in blocks[0][1:] if len(r) == len(blocks[0][0])]
iS, iI, iT = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
stalls = [c for c in h if c.startswith("stall_") and "Not" not in c]
tot_s = sum(int(r[iS]) for r in data); tot_i = sum(int(r[iI]) for r in data)
print(kern, len(data), "SASS instructions; samples", tot_s, "warp instructions", tot_i,
      "threads/inst %.1f" % (sum(int(r[iT]) for r in data) / max(tot_i, 1)))
agg = collections.Counter()
for r in data:
    for c in stalls: agg[c[6:]] += int(r[h.index(c)])
print("stall reasons:", dict(agg.most_common(8)))
for n in sorted(sorted(range(len(data)), key=lambda n: -int(data[n][iS]))[:top_n]):
    r = data[n]; st = {c[6:]: int(r[h.index(c)]) for c in stalls if int(r[h.index(c)]) > 0}
    print(n, r[1].strip()[:58].ljust(58), r[iS].rjust(4), r[iI].rjust(7), "%.0f" % (int(r[iT]) / max(int(r[iI]), 1)), st)
