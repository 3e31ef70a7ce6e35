#!/usr/bin/env python
"""Counts the SASS mnemonics that identify the Blackwell paths (tcgen05 / TMEM / TMA / bulk copies / mbarriers / packed
fp32) per kernel of wekws_b200/libwekws_b200.so, plus registers / spills from the ptxas logs.  Needs only cuobjdump
(no GPU).  usage: python scripts/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UBLKPF", "SYNCS", "FFMA2", "FADD2", "FFMA", "MUFU",
        "LDS", "STS", "LDG", "STG", "SHFL", "LDCU", "ELECT"]


def main():
    so = os.path.join(ROOT, "wekws_b200", "libwekws_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()
    counts, order, cur = {}, [], None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", ln)
        if m and cur:
            counts[cur][m.group(1)] += 1
    regs = {}
    for log in glob.glob(os.path.join(ROOT, "wekws_b200", "csrc", "*.ptxas.log")):
        name = None
        for ln in open(log):
            m = re.search(r"Compiling entry function '(\S+)'", ln)
            if m:
                name = m.group(1)
            m = re.search(r"(\d+) bytes spill stores", ln)
            if m and name:
                regs.setdefault(name, {})["spill"] = int(m.group(1))
            m = re.search(r"Used (\d+) registers", ln)
            if m and name:
                regs.setdefault(name, {})["regs"] = int(m.group(1))
    print("# SASS mnemonics per kernel in wekws_b200/libwekws_b200.so (cuobjdump -sass, sm_100a), round 2 -- scripts/sass_mnemonics.py")
    print("# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, UTMALDG = TMA tensor load,")
    print("# UBLKCP = cp.async.bulk, UBLKPF = bulk L2 prefetch, SYNCS = mbarrier ops, FFMA2 / FADD2 = packed f32x2, LDCU = uniform constant load,")
    print("# ELECT = elect_one_sync around the tcgen05 issue; regs / spill bytes from ptxas -v\n")
    for fn in order:
        c = counts[fn]
        short = re.sub(r"wekws::\(anonymous namespace\)::|\(anonymous namespace\)::|wekws::", "", demangle(fn))
        short = re.sub(r"\([^()]*\)$", "", short).replace("void ", "").replace("<unnamed>::", "")
        r = regs.get(fn, {})
        cols = " ".join(f"{k}={c[k]}" for k in KEEP if c[k])
        print(f"{short:34s} regs={r.get('regs', '?'):>3} spill={r.get('spill', '?'):>3}  {cols}")


if __name__ == "__main__":
    main()
