#!/usr/bin/env python3
"""DRAM traffic of the dominant kernel from an ncu metrics pass over ONE bench step (one pipeline, so that a launch is a whole bounce):
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:wf_trace_kernel \\
      --csv --log-file gpurun_out/traffic.csv python tools/sweep.py --spp 64 --reps 0 wf_streams=1
  python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/ncu_traffic.json c3 "C3, 64 spp, one step (...)" [path to record as source]
Merges {key: {"dram_bytes_per_launch": mean over the captured launches, the same over the bounce launches alone, ...}} into the JSON;
bench.py reports dram_bytes_per_launch as roofline.traffic."""
import csv
import json
import os
import sys

src, dst, key = sys.argv[1], sys.argv[2], sys.argv[3]
workload = sys.argv[4] if len(sys.argv) > 4 else ""
recorded_as = sys.argv[5] if len(sys.argv) > 5 else src
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ii, ki, mi, ui, vi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}
launch = {}
for r in rows[1:]:
    if "wf_trace_kernel" not in r[ki]:
        continue
    d = launch.setdefault(r[ii], {"primary": r[ki].split("<")[1].split(">")[0].replace(" ", "").endswith(",1")})   # <COUNT, QN, PRIMARY>
    d[r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
all_ = list(launch.values())
bounce = [d for d in all_ if not d["primary"]]


def dram(ds):
    return sum(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in ds)


rec = {"launches": len(all_), "dram_bytes_per_launch": dram(all_) / max(len(all_), 1),
       "bounce_launches": len(bounce), "dram_bytes_per_bounce_launch": dram(bounce) / max(len(bounce), 1),
       "seconds_under_ncu": sum(d.get("gpu__time_duration.sum", 0.0) for d in all_), "source": recorded_as, "workload": workload}
out = json.load(open(dst)) if os.path.exists(dst) else {}
out[key] = rec
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps({key: rec}))
