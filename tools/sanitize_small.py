#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rsba_b200.api as api  # noqa: E402
from rsba_b200.scene import make_scene  # noqa: E402

sc = make_scene(40, 1500, 8, name="sanitize")
priors = [(1, 5.0, 0.9, k, k - 1) for k in range(1, sc.num_frames)]
with api.Problem(0) as pb:
    pb.load_scene(sc)
    pb.set_motion_priors([p[0] for p in priors], [p[1] for p in priors], [p[2] for p in priors],
                         [p[3] for p in priors], [p[4] for p in priors])
    pb.set_loss(3.0)
    cost, r, J, v = pb.evaluate()
    ok, err = pb.validate()
    s = pb.solve(api.default_options(max_num_iterations=4))
    print("cost", cost, "->", s.final_cost, "iterations", s.iterations, "valid", int(ok.sum()), "/", ok.size)
assert s.usable == 1 and s.final_cost < cost

# second pass: free interFrameRatio + GoodPosePrior blocks + free intrinsics (the pseudo-frame border), and the
# iterative re-projection sweep
rng = np.random.default_rng(3)
vals = np.array([sc.poses[f, 6 * w:6 * w + 6] + rng.normal(0, 1e-3, 6) for f in range(1, sc.num_frames) for w in (0, 1)])
frames = [f for f in range(1, sc.num_frames) for _ in (0, 1)]
which = [w for _ in range(1, sc.num_frames) for w in (0, 1)]
with api.Problem(0) as pb:
    pb.set_intrinsics_free(True)
    pb.load_scene(sc)
    pb.set_motion_priors([2] * len(priors), [8.0] * len(priors), [1.0] * len(priors), [p[3] for p in priors],
                         [p[4] for p in priors])
    pb.set_inter_frame_ratio_free(True, 1.0)
    pb.set_pose_priors(frames, which, [10.0] * len(frames), [3.0] * len(frames), vals)
    s2 = pb.solve(api.default_options(max_num_iterations=4))
    xy, ok2 = pb.reproject(sc.obs_frame[:2000], sc.obs_point[:2000])
    print("cost", s2.initial_cost, "->", s2.final_cost, "ratio", pb.inter_frame_ratio(), "reprojected", int(ok2.sum()))
assert s2.usable == 1 and s2.final_cost < s2.initial_cost
