#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` dump by SASS opcode: executed warp instructions and stall samples."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iI, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops = defaultdict(lambda: [0, 0])
tot = [0, 0]
for r in rows[2:]:
    if len(r) <= iN:
        continue
    src = r[iS].strip()
    t = src.split()
    op = t[0]
    if op.startswith("@"):
        op = t[1]
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:3]) if op.startswith(("LDS", "STS", "LDG", "STG")) else "")
    n, smp = int(r[iI] or 0), int(r[iN] or 0)
    ops[op][0] += n; ops[op][1] += smp
    tot[0] += n; tot[1] += smp
print("total warp-inst %d  samples %d" % tuple(tot))
for op, (n, smp) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%-16s %12d %6.2f%%   samples %7d %6.2f%%" % (op, n, 100.0*n/tot[0], smp, 100.0*smp/max(tot[1], 1)))
