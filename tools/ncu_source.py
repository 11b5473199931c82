#!/usr/bin/env python3
"""Per-SASS-instruction view of the first kernel in an ncu source-page CSV:
ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv ; python tools/ncu_source.py src.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hdr = rows[1]
end = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
body = rows[2:end[1]] if len(end) > 1 else rows[2:]
ia, ie, it, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[ie]) for r in body)
totsamp = sum(int(r[isamp]) for r in body)
print("total warp-inst", tot, "samples", totsamp, "static instructions", len(body))
for k, r in enumerate(body):
    e = int(r[ie])
    if 100.0 * e / tot < minpct and 100.0 * int(r[isamp]) / max(totsamp, 1) < minpct:
        continue
    print("%4d %-72s exec %9d (%4.1f%%) thr/inst %4.1f samples %5d (%4.1f%%)" % (k, r[ia].strip()[:72], e, 100 * e / tot, int(r[it]) / max(e, 1), int(r[isamp]), 100 * int(r[isamp]) / max(totsamp, 1)))
