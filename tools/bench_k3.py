"""K3 alone on the GPU: the reduced-system solve of a video-like (banded) or fully covisible (dense) tile
pattern, task-graph kernel vs level-batched launches.  Usage: python tools/bench_k3.py [T] [bandwidth|dense]"""
import json
import sys
import os
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rsba_b200.api as api  # noqa: E402

TB = 96


def spd_band(T, bw, seed=0):
    rng = np.random.default_rng(seed)
    n = T * TB
    A = np.zeros((n, n))
    w = (bw + 1) * TB
    for a in range(T):
        lo, hi = a * TB, min(n, a * TB + w)
        G = rng.normal(size=(hi - lo, 16))
        A[lo:hi, lo:hi] += G @ G.T
    A += np.eye(n) * (1.0 + 0.01 * np.abs(np.diag(A)).mean())
    return A


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 125
    what = sys.argv[2] if len(sys.argv) > 2 else "3"
    dense = what == "dense"
    bw = T if dense else int(what)
    A = spd_band(T, bw)
    b = np.random.default_rng(1).normal(size=T * TB)
    pairs = [(a, c) for a in range(T) for c in range(a, min(T, a + bw + 1))]
    pa, pb = [p[0] for p in pairs], [p[1] for p in pairs]
    want = np.linalg.solve(A, b) if T <= 60 else None
    out = {"T": T, "pattern": "dense" if dense else f"band {bw}"}
    for mode, merge in (("levels", 1), ("dag", 1), ("dag", 4), ("dag", 16)):
        r = api.reduced_solve(A, b, T, pa, pb, dense=dense, mode=mode, merge_levels=merge, repeats=5)
        key = mode if mode == "levels" else f"dag_merge{merge}"
        out[key + "_ms"] = round(r["ms"], 4)
        if want is not None:
            out[key + "_relerr"] = float(np.linalg.norm(r["x"] - want) / np.linalg.norm(want))
        else:
            out[key + "_resid"] = float(np.abs(A @ r["x"] - b).max())
    plan = api.plan_reduced_system(T, pa, pb, dense=dense)
    out["tile_flops"] = plan["flops"]
    out["levels"] = plan["n_levels"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
