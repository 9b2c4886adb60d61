#!/usr/bin/env python
"""Summarise ncu outputs into the small text files kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv          > profiles/rNN_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [regex]      > profiles/rNN_<kernel>_full.txt

`launches` aggregates a `--metrics gpu__time_duration.sum --csv` log per kernel (count, total, share).
`full` prints the roofline-relevant raw metrics of every captured launch of a `--set full` report.
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
    "sm__sass_inst_executed_op_global_red.sum",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"^void |\(anonymous namespace\)::|rsba::|<unnamed>::", "", r["Kernel Name"]).split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r["Metric Unit"], 1.0)
        a = agg.setdefault(k, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v); n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {n} launches, {tot / 1e6:.3f} ms of kernel time (ncu: cold cache, serialised -> compare shares)")
    print(f"{'kernel':44s} {'n':>6s} {'total ms':>10s} {'share':>7s} {'avg us':>10s} {'min us':>10s} {'max us':>10s}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:44]:44s} {a[0]:6d} {a[1] / 1e6:10.3f} {a[1] / tot * 100:6.1f}% {a[1] / a[0] / 1e3:10.1f} {a[2] / 1e3:10.1f} {a[3] / 1e3:10.1f}")


def full(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for row in rows[2:]:
        if pat and not re.search(pat, row[name_col]):
            continue
        print(f"## {row[name_col].split('(')[0]}  (id {row[0]})")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:86s} {row[i]:>16s} {units[i]}")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
