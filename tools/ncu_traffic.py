#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report: per kernel the DRAM bytes (read + write) of one launch
(median over the captured launches).  bench.py reads it for `roofline.traffic`.
   python tools/ncu_traffic.py <report.ncu-rep> <config, e.g. C3> <source label> [commit]"""
import csv
import json
import os
import re
import statistics
import subprocess
import sys

rep, config, source = sys.argv[1], sys.argv[2], sys.argv[3]
commit = sys.argv[4] if len(sys.argv) > 4 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True,
                                                              text=True).stdout.strip()
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].split("::")[-1].strip()    # e.g. schur_syrk_kernel, k1_kernel<1, 0, 1>
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[col[m]].replace(",", "")) * scale[units[col[m]]]
    per.setdefault(name, []).append(tot)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
try:
    cur = json.load(open(path))
except (OSError, ValueError):
    cur = {}
for k, v in per.items():
    cur[k] = {"config": config, "bytes": statistics.median(v), "launches": len(v), "source": source, "commit": commit}
json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps({k: cur[k]["bytes"] for k in per}, indent=1))
