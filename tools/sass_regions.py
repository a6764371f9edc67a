"""tools/sass_regions.py REPORT KERNEL [launch-index]: runs of SASS instructions with the same execution count / active lanes
(finds code a warp runs several times in divergent groups)"""
import csv, subprocess, sys, collections, io
rep, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = []; blocks.append(cur); continue
    if cur is not None: cur.append(r)
h, data = blocks[which][0], [r for r in blocks[which][1:] if len(r) == len(blocks[which][0])]
iS, iI, iT = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
tot_s = sum(int(r[iS]) for r in data); tot_i = sum(int(r[iI]) for r in data)
print(kern, len(data), "SASS; samples", tot_s, "warp inst", tot_i)
prev = None; start = 0; acc_s = 0; acc_i = 0
def flush(a, b, key, s, i):
    if i / tot_i > 0.01 or s / max(tot_s, 1) > 0.01:
        print("%5d-%5d n=%4d exec~%9d lanes %4.1f  inst %5.1f%% samples %5.1f%%" % (a, b - 1, b - a, key[0], key[1], 100 * i / tot_i, 100 * s / tot_s))
for n, r in enumerate(data):
    i = int(r[iI]); t = int(r[iT]) / max(i, 1)
    key = (i, t)
    if prev is None or abs(i - prev[0]) > 0.03 * max(prev[0], 1) or abs(t - prev[1]) > 1.5:
        if prev is not None: flush(start, n, prev, acc_s, acc_i)
        prev = key; start = n; acc_s = 0; acc_i = 0
    acc_s += int(r[iS]); acc_i += i
flush(start, len(data), prev, acc_s, acc_i)
