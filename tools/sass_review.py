#!/usr/bin/env python
"""Static review of the built library's SASS, one line per kernel: instructions, registers / spills (ptxas logs of the
build), indirect branches (BRX: jump tables -- the two round-2 performance bugs), local-memory accesses, divergent
regions (BSSY), FP64 tensor instructions (DMMA), TMA bulk copies (UBLKCP), mbarrier instructions (SYNCS).
   python tools/sass_review.py > profiles/rNN_sass_review.txt"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "rsba_b200", "lib", "librsba_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, name = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        kern[name] = []
    elif name and re.search(r"/\*[0-9a-f]{4,5}\*/", line):
        kern[name].append(line)
regs = {}
for log in glob.glob(os.path.join(ROOT, "rsba_b200", "csrc", "build", "*.ptxas.log")):
    cur = None
    for line in open(log):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
        m = re.search(r"(\d+) bytes spill stores", line)
        if m and cur:
            d = regs.setdefault(cur, {})
            d["spill"] = max(d.get("spill", 0), int(m.group(1)))   # (device sub-functions report their own, zero, frames)
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            regs.setdefault(cur, {})["regs"] = int(m.group(1))
demangle = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.relpath(lib, ROOT)}: {len(kern)} kernels")
print(f"{'kernel':58s} {'instr':>6s} {'regs':>4s} {'spill':>5s} {'BRX':>3s} {'LDL':>3s} {'STL':>3s} {'BSSY':>4s} {'DMMA':>4s} {'UBLKCP':>6s} {'SYNCS':>5s}")
rows = []
for (k, lines), dm in zip(kern.items(), demangle):
    short = re.sub(r"\(anonymous namespace\)::|rsba::|void ", "", dm).split("(")[0]
    c = lambda pat: sum(1 for l in lines if re.search(pat, l))
    r = regs.get(k, {})
    rows.append((short[:58], len(lines), r.get("regs", -1), r.get("spill", -1), c(r"\bBRX\b"), c(r"\bLDL"), c(r"\bSTL"),
                 c(r"\bBSSY"), c(r"\bDMMA"), c(r"\bUBLKCP"), c(r"\bSYNCS")))
for row in sorted(rows, key=lambda r: -r[1]):
    print("%-58s %6d %4d %5d %3d %3d %3d %4d %4d %6d %5d" % row)
