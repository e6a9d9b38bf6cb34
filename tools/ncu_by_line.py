#!/usr/bin/env python
"""Dynamic profile per source line: joins an `ncu --page source --csv --print-source sass` export with
`nvdisasm -g -c` of the profiled cubin (same build!) by instruction offset inside swb_route_kernel.
    python tools/ncu_by_line.py src.csv sass.txt [top] [func-substring]
Prints, per (function, file:line of the innermost inlined location): samples, long-scoreboard samples,
warp instructions."""
import collections
import csv
import re
import sys

src, sass = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
only = sys.argv[4] if len(sys.argv) > 4 else None
kern = sys.argv[5] if len(sys.argv) > 5 else "_Z16swb_route_kernel"
lines = open(sass).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text." + kern)][0]
lre = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
fre = re.compile(r"^\$?(_Z\w+?)(?:\$(\w+))?:\s*$")
ire = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")
off2loc = {}
func, cur = "kernel", None
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.lstrip().startswith(".section"):
        break
    s = l.strip()
    m = fre.match(s)
    if m and not s.startswith(".L"):
        func = m.group(2) or "kernel"
        continue
    m = lre.search(l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = ire.match(l)
    if m:
        off2loc[int(m.group(1), 16)] = (func, cur, m.group(2))
rows = list(csv.reader(open(src)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][0], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in data:
    try:
        a = int(r[0], 16) - base
        smp = int(r[col["# Samples"]] or 0)
        ins = int(r[col["Instructions Executed"]] or 0)
        lsb = int(r[col["stall_long_sb"]] or 0)
    except (ValueError, IndexError):
        continue
    f, loc, _ = off2loc.get(a, ("?", None, ""))
    f = re.sub(r"^_ZN3swb\d+", "", f)[:28]
    if only and only not in f:
        continue
    k = (f, loc)
    agg[k][0] += smp; agg[k][1] += lsb; agg[k][2] += ins
    tot[0] += smp; tot[1] += lsb; tot[2] += ins
print("total samples %d long_sb %d warp-instr %d" % tuple(tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% smp  %5.2f%% lsb  %5.2f%% ins  %-28s %s" % (100 * v[0] / tot[0], 100 * v[1] / max(tot[1], 1),
                                                           100 * v[2] / tot[2], k[0], "%s:%d" % k[1] if k[1] else "-"))
