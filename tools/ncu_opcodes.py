#!/usr/bin/env python
"""Dynamic opcode histogram of the profiled kernel: tools/ncu_opcodes.py file.ncu-rep [top]"""
import csv, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
hist, stall = collections.Counter(), collections.Counter()
tot = 0
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    src = r[hdr.index("Source")].split()
    if not src:
        continue
    op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
    op = op.rstrip(";")
    n = int(r[hdr.index("Instructions Executed")] or 0)
    hist[op] += n
    stall[op] += int(r[hdr.index("Warp Stall Sampling (All Samples)")] or 0)
    tot += n
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("total warp instructions", tot)
for op, n in hist.most_common(top):
    print("%-28s %12d %5.1f%%   stall samples %d" % (op, n, 100.0 * n / tot, stall[op]))
