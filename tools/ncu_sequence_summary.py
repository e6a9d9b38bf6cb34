#!/usr/bin/env python
"""Sums an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,...` log of one
bench launch sequence (all kernels of swb_run_steps) per kernel name and in total.
    python tools/ncu_sequence_summary.py sequence.csv launch_plain.json out.json"""
import collections
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0,
        "s": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3, "inst": 1.0, "": 1.0,
        "register/thread": 1.0}
per = collections.defaultdict(lambda: collections.defaultdict(float))
count = collections.Counter()
ids = set()
for r in rows[1:]:
    name = r[col["Kernel Name"]].split("(")[0]
    metric, unit, val = r[col["Metric Name"]], r[col["Metric Unit"]], float(r[col["Metric Value"]].replace(",", ""))
    scale = UNIT.get(unit, 1.0)
    if metric == "launch__registers_per_thread":
        per[name]["registers"] = val
        continue
    per[name][metric] += val * scale
    if metric == "gpu__time_duration.sum":
        count[name] += 1
tot = collections.defaultdict(float)
for name, d in per.items():
    for k, v in d.items():
        if k != "registers":
            tot[k] += v
plain = json.load(open(sys.argv[2]))
out = {
    "members": plain["members"], "grid": plain["grid"], "routing_steps": plain["routing_steps"],
    "dram_bytes_read": tot["dram__bytes_read.sum"], "dram_bytes_write": tot["dram__bytes_write.sum"],
    "conduit_updates_in_launch": plain["conduit_updates_in_launch"],
    "gpu_time_ms": tot["gpu__time_duration.sum"], "warp_instructions": tot["smsp__inst_executed.sum"],
    "kernels_in_launch": int(sum(count.values())),
    "per_kernel": {n: {"launches": count[n], "time_ms": round(d["gpu__time_duration.sum"], 3),
                       "share": round(d["gpu__time_duration.sum"] / max(tot["gpu__time_duration.sum"], 1e-9), 4),
                       "dram_GB": round((d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / 1e9, 3),
                       "warp_instructions": d["smsp__inst_executed.sum"], "registers": d.get("registers")}
                   for n, d in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"])},
    "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum "
              "--profile-from-start off --clock-control none, python tools/profile_launch.py (tools/ncu_capture_r2.sh): "
              "sums over every kernel of ONE swb_run_steps launch sequence; per-kernel times are serialised and cold-cache",
    "same_launch_without_ncu": plain,
}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k not in ("per_kernel", "same_launch_without_ncu", "source")}))
for n, d in out["per_kernel"].items():
    print("%-60s %s" % (n[:60], d))
