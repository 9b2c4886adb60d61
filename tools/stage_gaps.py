#!/usr/bin/env python
"""Where does an LM iteration's time go besides the big kernels?  Runs linearize_and_step passes at C3 with the
per-kernel events on and prints every stage timer next to the enclosing stage's total."""
import statistics
import sys

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import rsba_b200.api as api
from rsba_b200.scene import make_config

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
sc = make_config(cfg)
with api.Problem(0) as pb:
    pb.load_scene(sc)
    opt = api.default_options()
    opt.max_num_iterations = 1
    samples = {k: [] for k in api.STAGES}
    for _ in range(8):
        pb.linearize_and_step(1e4, opt, want_S=False, fetch=False)
        for k in api.STAGES:
            samples[k].append(pb.stage_ms(k))
    med = {k: statistics.median(v[2:]) for k, v in samples.items()}
    for k, v in med.items():
        print(f"{k:14s} {v:8.4f} ms")
    inner = med["point_blocks"] + med["frame_blocks"] + med["phi_build"] + med["schur_syrk"] + med["schur_reduce"]
    print(f"schur stage {med['schur']:.4f} = timed kernels {inner:.4f} + other {med['schur'] - inner:.4f}; finalize {med['finalize']:.4f}")
