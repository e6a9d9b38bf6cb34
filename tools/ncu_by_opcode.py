#!/usr/bin/env python
"""Dynamic instruction mix of one kernel from an `ncu --page source --csv --print-source sass` export:
warp instructions executed and stall samples per opcode.
    python tools/ncu_by_opcode.py src.csv [top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0])
tot = [0, 0]
for r in rows[2:]:
    try:
        ins = int(r[col["Instructions Executed"]] or 0)
        smp = int(r[col["# Samples"]] or 0)
    except (ValueError, IndexError):
        continue
    t = r[col["Source"]].split()
    if not t:
        continue
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    op = re.sub(r"[;,]", "", op)
    key = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LD", "ST", "MUFU", "ATOM")) and "." in op else "")
    agg[key][0] += ins; agg[key][1] += smp
    tot[0] += ins; tot[1] += smp
print("warp instructions %d, samples %d" % tuple(tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-14s %6.2f%% ins  %6.2f%% smp" % (k, 100 * v[0] / tot[0], 100 * v[1] / max(tot[1], 1)))
