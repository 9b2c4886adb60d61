#!/usr/bin/env python
"""Cold call at a BASELINE config: fresh handle -> set_scene -> solve(K) -> get_parameters -> destroy, wall clock, with
the library's own phase trace (RSBA_CUDA_TRACE=1) on stderr.   python tools/cold_call.py [C3] [iterations] [repeats]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rsba_b200.api as api
from rsba_b200.scene import make_config

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sc = make_config(cfg)
with api.Problem(0) as warm:      # CUDA context + module load are paid once per process, not per BA call
    warm.load_scene(make_config("C1"))
    warm.solve(api.default_options(max_num_iterations=1))
for r in range(reps):
    t0 = time.perf_counter()
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        t1 = time.perf_counter()
        s = pb.solve(api.default_options(max_num_iterations=iters, function_tolerance=0.0, gradient_tolerance=0.0,
                                         parameter_tolerance=0.0))
        t2 = time.perf_counter()
        pb.get_parameters()
    t3 = time.perf_counter()
    print(f"cold call {cfg} rep {r}: total {1e3 * (t3 - t0):.1f} ms = load_scene {1e3 * (t1 - t0):.1f} + solve({s.iterations} it) "
          f"{1e3 * (t2 - t1):.1f} + get/destroy {1e3 * (t3 - t2):.1f}", flush=True)
