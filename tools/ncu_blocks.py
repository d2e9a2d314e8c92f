#!/usr/bin/env python
"""Basic-block view of an ncu report's SASS page: runs of instructions with the same executed
count, their share of all executed warp instructions and their dominant opcodes.
    tools/ncu_blocks.py file.ncu-rep [min_share_pct]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
rows = list(csv.reader(out.splitlines()))
hdr, data = None, []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) >= len(hdr):
        data.append(r)
isrc, ie, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ie] or 0) for r in data)
tst = sum(int(r[ist] or 0) for r in data)
blocks, cur = [], None
for i, r in enumerate(data):
    n = int(r[ie] or 0)
    if cur and cur[2] == n:
        cur[1] = i
    else:
        cur = [i, i, n]
        blocks.append(cur)
print("static instructions %d, executed warp instructions %d" % (len(data), tot))
for b in blocks:
    ln = b[1] - b[0] + 1
    share = 100.0 * ln * b[2] / tot
    if share < minshare:
        continue
    ops = {}
    st = 0
    for r in data[b[0]:b[1] + 1]:
        s = r[isrc].split()
        op = (s[1] if s[0].startswith("@") and len(s) > 1 else s[0]).rstrip(";")
        ops[op] = ops.get(op, 0) + 1
        st += int(r[ist] or 0)
    top = " ".join("%s:%d" % kv for kv in sorted(ops.items(), key=lambda x: -x[1])[:6])
    print("rows %4d-%4d len %3d exec %9d share %5.1f%% stall %5.1f%%  %s" % (b[0], b[1], ln, b[2], share, 100.0 * st / max(tst, 1), top))
