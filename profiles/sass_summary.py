#!/usr/bin/env python
"""Opcode summary of every kernel in tensortoolkit_b200/libqlb200.so (cuobjdump -sass): which pipes / copy mechanisms each
kernel really uses.  usage: python profiles/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "tensortoolkit_b200/libqlb200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["DMMA", "DFMA", "DADD", "DMUL", "LDGSTS", "LDG", "STG", "LDS", "STS", "SYNCS", "BAR", "ATOMG", "RED", "UBLKCP", "UTMALDG", "UTMASTG",
         "UTCMMA", "LDTM", "MULTIMEM", "USETMAXREG", "MEMBAR", "NOP"]
fn, counts = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"qlb200::\(anonymous namespace\)::", "", fn)
        fn = re.sub(r"\(qlb200::GemmParams\)|\(.*\)$", "", fn)
        counts[fn] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and fn:
        op = m.group(1)
        counts[fn]["_total"] += 1
        for w in WATCH:
            if op.split(".")[0] == w or (w == "MULTIMEM" and "MULTIMEM" in op):
                counts[fn][w] += 1
print(f"# SASS opcode counts per kernel of {lib} (sm_100a); columns = static instruction counts")
print(f"# tcgen05 (UTCMMA / LDTM) does not appear: it has no FP64 kind; the FP64 tensor path is mma.sync -> DMMA.8x8x4")
cols = [w for w in WATCH if any(c[w] for c in counts.values())]
print(f"{'kernel':78s} {'total':>7s} " + " ".join(f"{c:>8s}" for c in cols))
for fn, c in counts.items():
    print(f"{fn[:78]:78s} {c['_total']:7d} " + " ".join(f"{c[w]:8d}" for w in cols))
