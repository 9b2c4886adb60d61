#!/usr/bin/env python
"""Per-source-line warp-stall samples of one kernel from an `ncu --set full --import-source on` report.
   python tools/ncu_source_hot.py <report.ncu-rep> <kernel regex> [top N]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{pat}", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hdr]
samp = h.index("# Samples")
lines = []
for r in rows[hdr + 1:]:
    if len(r) > samp and r[0].isdigit():
        lines.append((int(r[samp]) if r[samp].isdigit() else 0, int(r[0]), r[1].strip()))
tot = sum(x[0] for x in lines) or 1
print(f"# {rows[1][1][:100]}: {tot} samples")
for s, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{s:7d} {100.0 * s / tot:5.1f}%  L{ln:<4d} {src[:110]}")
