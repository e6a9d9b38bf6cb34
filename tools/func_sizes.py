#!/usr/bin/env python
"""Sizes of the device functions inside swb_route_kernel (from `cuobjdump -elf`)."""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "stormwater-management-model_b200/csrc/libswmm_b200.so"
out = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
rows = []
for ln in out.splitlines():
    m = re.match(r"\s+0x[0-9a-f]+\s+0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+.*route_kernel\S*\$(\S+)", ln)
    if m:
        rows.append((int(m.group(1), 16), m.group(2)))
    m = re.match(r"\s+[0-9a-f]+\s+[0-9a-f]+\s+([0-9a-f]+)\s+.*PROGBITS.*\.text\.(\S*route_kernel\S*)", ln)
    if m:
        rows.append((int(m.group(1), 16), "TOTAL .text " + m.group(2)[:30]))
for sz, n in sorted(rows, reverse=True)[:16]:
    print(f"{sz / 1024:8.1f} KB {sz // 16:7d} instr  {n[:80]}")
