"""Per-task time stamps of the K3 task-graph kernel (rsba_cuda_reduced_solve's trace_out) on a banded or dense
tile pattern: saves the raw trace and prints where the time of the critical chain goes.
Usage: python tools/trace_k3.py [T] [bandwidth|dense] [merge_levels] [out.npz]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rsba_b200.api as api  # noqa: E402
from bench_k3 import spd_band  # noqa: E402

NAMES = ["factor", "trsm", "update", "back_fin", "back_tile"]


def summarize(trace, tasks):
    t0 = trace[:, 0].min()
    out = {"span_us": float(trace[:, 2].max() - t0) / 1e3}
    for ty, nm in enumerate(NAMES):
        m = tasks[:, 0] == ty
        if not m.any():
            continue
        wait = (trace[m, 1] - trace[m, 0]) / 1e3
        run = (trace[m, 2] - trace[m, 1]) / 1e3
        run_ck = (trace[m, 5] - trace[m, 4])
        out[nm] = {"n": int(m.sum()), "run_us_median": float(np.median(run)), "run_us_max": float(run.max()),
                   "run_cycles_median": float(np.median(run_ck)), "wait_us_median": float(np.median(wait)),
                   "wait_us_max": float(wait.max())}
    fac = np.flatnonzero(tasks[:, 0] == 0)
    f = trace[fac]
    out["factor_phase_cycles_median"] = {
        "load": float(np.median(f[:, 8] - f[:, 4])), "factorise": float(np.median(f[:, 9] - f[:, 8])),
        "invert": float(np.median(f[:, 10] - f[:, 9])), "store_and_forward": float(np.median(f[:, 11] - f[:, 10])),
        "panels_sum_warp0": float(np.median(f[:, 12])), "phase1_sum": float(np.median(f[:, 13]))}
    ends = np.sort(trace[fac, 2] - t0) / 1e3
    out["factor_end_us_sorted_tail"] = [float(v) for v in ends[-24:]]
    nf = (tasks[:, 0] <= 2)
    out["factorisation_span_us"] = float(trace[nf, 2].max() - t0) / 1e3
    out["backward_span_us"] = float(trace[~nf, 2].max() - trace[nf, 2].max()) / 1e3 if (~nf).any() else 0.0
    out["sms_used"] = int(len(set(trace[:, 6].tolist())))
    return out


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 125
    what = sys.argv[2] if len(sys.argv) > 2 else "3"
    merge = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    dense = what == "dense"
    bw = T if dense else int(what)
    A = spd_band(T, bw)
    b = np.random.default_rng(1).normal(size=T * 96)
    pairs = [(a, c) for a in range(T) for c in range(a, min(T, a + bw + 1))]
    pa, pb = [p[0] for p in pairs], [p[1] for p in pairs]
    r = api.reduced_solve(A, b, T, pa, pb, dense=dense, merge_levels=merge, repeats=3, want_trace=True)
    dag = api.plan_task_graph(T, pa, pb, dense=dense, merge_levels=merge)
    out = summarize(r["trace"], dag["tasks"])
    out.update(T=T, pattern=what, merge_levels=merge, ms=r["ms"])
    print(json.dumps(out))
    if len(sys.argv) > 4:
        np.savez_compressed(sys.argv[4], trace=r["trace"], tasks=dag["tasks"], sources=dag["sources"])


if __name__ == "__main__":
    main()
