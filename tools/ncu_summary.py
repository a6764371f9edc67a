#!/usr/bin/env python
"""Summarise ncu outputs into profiles/: a launch list (gpu__time_duration per kernel, shares)
and the key counters of a --set full capture.  Usage:
  tools/ncu_summary.py launches <launches.csv> <out.md>
  tools/ncu_summary.py full <report.ncu-rep> <out.md>"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        try:
            d[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.1f | %.3f |\n" % (k, len(v), sum(v) / 1e3, sum(v) / tot))


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        names = [r[hdr.index("Kernel Name")].split("(")[0] for r in rows[2:]]
        f.write("| metric | unit | " + " | ".join("`%s`" % n for n in names) + " |\n|---|---|" + "---:|" * len(names) + "\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write("| %s | %s | " % (k, units[i]) + " | ".join(r[i] for r in rows[2:]) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
