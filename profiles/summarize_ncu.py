#!/usr/bin/env python
"""Summaries of ncu output for profiles/ (run in the build container; ncu -i needs no GPU).

  summarize_ncu.py launches <launches.csv> "<command line>"   -> per-kernel launch list (shares of kernel time)
  summarize_ncu.py full <report.ncu-rep>                       -> one CSV row per captured launch, selected metrics
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name


def launches(path, command):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head, rows = rows[0], rows[1:]
    k, v, u = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows:
        ns = float(r[v].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[u], 1.0)
        t = tot.setdefault(short(r[k]), [0, 0.0])
        t[0] += 1
        t[1] += ns
    total = sum(t[1] for t in tot.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none --csv, command: %s" % command)
    print("# Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  %d launches, %.1f ms of kernel time."
          % (len(rows), total / 1e6))
    for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-78s launches %4d  total %10.1f us  avg %8.1f us  share %5.1f%%" % (name[:78], n, ns / 1e3, ns / 1e3 / n, 100 * ns / total))


def full(path):
    if path.endswith(".csv"):          # `ncu -i X.ncu-rep --page raw --csv` exported on the GPU box (the report itself can exceed what travels back)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, rows = rows[0], rows[1], rows[2:]
    cols = [head.index("Kernel Name")] + [head.index(m) for m in METRICS]
    w = csv.writer(sys.stdout)
    w.writerow(["Kernel Name"] + METRICS)
    w.writerow([units[c] for c in cols])
    for r in rows:
        w.writerow([r[c] for c in cols])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        full(sys.argv[2])
