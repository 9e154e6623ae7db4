#!/usr/bin/env python
"""Per basic block share of executed instructions and of warp-stall samples from `ncu -i X.ncu-rep --page source --csv`.

    ncu -i rep.ncu-rep --page source --csv > src.csv;  python profiles/source_blocks.py src.csv <kernel substring> [section #]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
want = [s for s in sections if sys.argv[2] in s["name"]]
sec = want[int(sys.argv[3]) if len(sys.argv) > 3 else 0]
head, body = sec["rows"][0], [r for r in sec["rows"][1:] if len(r) > 10]
isrc, iex, ism = head.index("Source"), head.index("Instructions Executed"), head.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[iex]) for r in body)
tots = sum(int(r[ism]) for r in body)
print(sec["name"])
print("warp instructions", tot, " samples", tots, " SASS lines", len(body))
acc = accs = start = 0
for i, r in enumerate(body):
    acc += int(r[iex])
    accs += int(r[ism])
    op = r[isrc].strip()
    if "BAR" in op or "BRA" in op or "EXIT" in op or i == len(body) - 1:
        if acc / tot > 0.004 or accs / max(tots, 1) > 0.004:
            print("%4d-%4d %-46s inst %5.1f%%  samples %5.1f%%" % (start, i, op[:46], 100 * acc / tot, 100 * accs / max(tots, 1)))
        acc = accs = 0
        start = i + 1
