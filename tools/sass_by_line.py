#!/usr/bin/env python
"""Static SASS footprint per (device function, source region).
    cuobjdump -xelf all libswmm_b200.so && nvdisasm -g -c swb_api.sm_100a.cubin > sass.txt
    python tools/sass_by_line.py sass.txt [min_instr]
Buckets instructions of every out-of-line function of swb_route_kernel by the source file and a
25-line window of the innermost inlined line: shows which inlined helpers bloat the hot functions."""
import collections
import re
import sys

lines = open(sys.argv[1]).read().split("\n")
min_instr = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
start = [i for i, l in enumerate(lines) if l.startswith(".text._Z16swb_route_kernel")][0]
fre = re.compile(r"^\$?(_Z\w+?)(?:\$(\w+))?:\s*$")
lre = re.compile(r'//## File "([^"]+)", line (\d+)')
func, cur = "kernel", None
hist = collections.defaultdict(collections.Counter)
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = fre.match(l.strip())
    if m and not l.strip().startswith(".L"):
        func = m.group(2) or "kernel"
        continue
    m = lre.search(l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", l):
        hist[func][cur] += 1
for f, h in sorted(hist.items(), key=lambda kv: -sum(kv[1].values())):
    tot = sum(h.values())
    if tot < min_instr:
        continue
    print(f"{f}: {tot} instr")
    agg = collections.Counter()
    for k, c in h.items():
        if k:
            agg[(k[0], k[1] // 25 * 25)] += c
    for k, c in agg.most_common(18):
        print(f"    {k[0]}:{k[1]}-{k[1] + 24}  {c}")
