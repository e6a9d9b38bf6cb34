#!/usr/bin/env python
"""Aggregates an `ncu --page source --csv --print-source sass` export per device function of
swb_route_kernel (function address ranges come from `cuobjdump -elf` of the profiled library).
    python tools/ncu_by_function.py src.csv [libswmm_b200.so]"""
import csv
import re
import subprocess
import sys

src = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "stormwater-management-model_b200/csrc/libswmm_b200.so"
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
funcs = []
for ln in elf.splitlines():
    m = re.match(r"\s+0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+.*route_kernel\S*\$(\S+)", ln)
    if m:
        funcs.append((int(m.group(1), 16), int(m.group(2), 16), m.group(3)))
rows = list(csv.reader(open(src)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][0], 16)
agg = {}
tot_s = tot_i = 0
for r in data:
    try:
        a = int(r[0], 16) - base
        smp = int(r[col["# Samples"]] or 0)
        ins = int(r[col["Instructions Executed"]] or 0)
        thr = int(r[col["Thread Instructions Executed"]] or 0)
    except (ValueError, IndexError):
        continue
    name = "<kernel body>"
    for off, sz, n in funcs:
        if off <= a < off + sz:
            name = n
            break
    d = agg.setdefault(name, {"s": 0, "i": 0, "t": 0, "st": {}})
    d["s"] += smp; d["i"] += ins; d["t"] += thr
    for k in ("stall_long_sb", "stall_no_inst", "stall_barrier", "stall_wait", "stall_selected",
              "stall_short_sb", "stall_branch_resolving", "stall_math", "stall_lg"):
        d["st"][k] = d["st"].get(k, 0) + int(r[col[k]] or 0)
    tot_s += smp; tot_i += ins
print(f"total samples {tot_s}, warp instructions {tot_i}")
for name, d in sorted(agg.items(), key=lambda x: -x[1]["s"])[:14]:
    st = ", ".join(f"{k[6:]} {v * 100 // max(d['s'], 1)}%" for k, v in sorted(d["st"].items(), key=lambda x: -x[1])[:4])
    print(f"{d['s'] * 100.0 / tot_s:5.1f}% samples {d['i'] * 100.0 / tot_i:5.1f}% instr  "
          f"lanes {d['t'] / max(d['i'], 1):4.1f}  {name[:44]:44s} {st}")
