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
