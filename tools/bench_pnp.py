#!/usr/bin/env python
"""Throughput of the batched RS-PnP kernel (rsba_cuda_pnp_batch): RANSAC hypotheses per second for
one frame with n 2-D/3-D matches, minimal samples of 6, 10 LM iterations each + inlier scoring.
   python tools/bench_pnp.py [--hyp 16384] [--points 1000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rsba_b200.api as api  # noqa: E402
from rsba_b200.scene import make_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--hyp", type=int, default=16384)
ap.add_argument("--points", type=int, default=1000)
args = ap.parse_args()

sc = make_scene(6, max(args.points * 6 // 4, 600), 4, name="pnp-bench")
sel = np.flatnonzero(sc.obs_frame == 3)[:args.points]
pts = sc.points_true[sc.obs_point[sel]].astype(np.float32).astype(np.float64)
xy = sc.obs_xy[sel].astype(np.float32).astype(np.float64)
xy[::4] += 40.0                                            # 25 % wrong matches
n = pts.shape[0]
rng = np.random.default_rng(0)
idx = rng.integers(0, n, (args.hyp, 6)).astype(np.int32)
pose0 = sc.poses_true[3] + rng.normal(0, 2e-3, 12)
poses = np.tile(pose0, (args.hyp, 1))
with api.Problem(0) as pb:
    pb.pnp_batch(sc.cam, sc.shutter, sc.scanlines, pts, xy, idx[:256], poses[:256])      # warm-up
    t0 = time.perf_counter()
    out = pb.pnp_batch(sc.cam, sc.shutter, sc.scanlines, pts, xy, idx, poses, inlier_threshold=3.0)
    wall = time.perf_counter() - t0
    kern = pb.stage_ms("pnp")
print(json.dumps({"workload": f"{args.hyp} RANSAC hypotheses x 6 points, 10 LM iterations, inlier sweep over {n} matches",
                  "kernel_ms": kern, "hypotheses_per_s_kernel": args.hyp / (kern * 1e-3),
                  "wall_ms_incl_host_copies": wall * 1e3, "hypotheses_per_s_e2e": args.hyp / wall,
                  "mean_iterations": float(out["iterations"].mean()), "best_inliers": int(out["inliers"].max()),
                  "points": int(n)}))
