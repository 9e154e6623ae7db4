#!/usr/bin/env python
"""Print a compact table from a bench.py JSON line (helper for profiles/)."""
import json
import sys

d = json.load(open(sys.argv[1]))
for k in ('value', 'ms_per_step', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline'):
    print(k, d.get(k))
print()
for k, v in d['ops'].items():
    print("%-26s %8.3f ms  %7.1f GB/s  frac %.3f   %.2e el/s" % (k, v['ms'], v['gbs'], v['frac'], v['elements_per_s']))
