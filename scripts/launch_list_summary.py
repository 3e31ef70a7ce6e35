#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: launch_list_summary.py <launches.csv> "<command line that was profiled>" """
import csv, re, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*$", "", r[4]).replace("void ", "")
    name = name if len(name) < 72 else name[-72:]
    tot[name][0] += 1
    tot[name][1] += float(r[-1]) / 1000.0
allus = sum(v[1] for v in tot.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print(f"(first {len(rows)} kernel launches of the process; cold-cache, serialised: compare shares, not absolutes)\n")
print(f"{'kernel':72s} {'launches':>8s} {'total us':>12s} {'share':>7s}")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {n:8d} {us:12.1f} {100 * us / allus:6.1f}%")
