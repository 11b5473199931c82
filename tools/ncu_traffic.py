#!/usr/bin/env python3
"""DRAM traffic of the dominant kernel from an ncu metrics pass over ONE bench step:
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:wf_trace_kernel \\
      --csv --log-file gpurun_out/traffic.csv python tools/sweep.py --spp 64 --reps 0
  python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/ncu_traffic.json
Writes {"dram_bytes_per_launch": mean over the captured launches, ...}; bench.py reports it as roofline.traffic."""
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}
tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
n = 0
for r in rows[1:]:
    if r[mi] in tot:
        tot[r[mi]] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
        n += r[mi] == "gpu__time_duration.sum"
out = {"kernel": "wf_trace_kernel", "launches": n, "dram_bytes_read": tot["dram__bytes_read.sum"], "dram_bytes_write": tot["dram__bytes_write.sum"],
       "dram_bytes_per_launch": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / max(n, 1),
       "seconds_under_ncu": tot["gpu__time_duration.sum"], "source": sys.argv[1]}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
