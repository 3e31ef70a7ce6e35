#!/usr/bin/env python
"""Top source lines by warp-stall samples from an ncu report captured with --import-source on.
usage: profile_hotlines.py <report.ncu-rep> [n]"""
import csv, io, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur, hdr = None, None
agg = defaultdict(lambda: defaultdict(float))
src = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None or len(r) < len(hdr):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    key = (cur, line)
    src[key] = r[1].strip()[:110]
    for i, h in enumerate(hdr):
        if h == "# Samples" or (h.startswith("stall_") and "Not Issued" not in h) or h == "Instructions Executed":
            try:
                agg[key][h] += float(r[i])
            except ValueError:
                pass
tot = sum(v["# Samples"] for v in agg.values()) or 1
print(f"total samples {tot:.0f}")
keyf = (lambda kv: -kv[1]["Instructions Executed"]) if len(sys.argv) > 3 and sys.argv[3] == "inst" else (lambda kv: -kv[1]["# Samples"])
for key, v in sorted(agg.items(), key=keyf)[:n]:
    st = sorted(((x, h[6:]) for h, x in v.items() if h.startswith("stall_") and x > 0), reverse=True)[:3]
    print(f"{100 * v['# Samples'] / tot:5.1f}%  {key[0]}:{key[1]:<4d} inst {v['Instructions Executed']:>9.0f}  " +
          ",".join(f"{h} {x:.0f}" for x, h in st) + f"   | {src[key]}")
