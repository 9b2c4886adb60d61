"""Generates tests/golden/*.npz by running the REFERENCE's own headers (oracle/_ref, compiled
verbatim from /root/reference/src) on fixed inputs.  Needs /root/reference, so it runs only in
the build container; the .npz files are committed and travel to the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from helpers import edge_scene, small_scene  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dump(name, scene, keep=None):
    res, J, valid = oracle.evaluate(scene, impl="ref")
    sel = slice(None) if keep is None else keep
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        cam=scene.cam, shutter=np.int32(scene.shutter), scanlines=scene.scanlines,
        interpolate_rotation=np.int32(scene.interpolate_rotation), poses=scene.poses, points=scene.points,
        obs_xy=scene.obs_xy[sel], obs_frame=scene.obs_frame[sel], obs_point=scene.obs_point[sel],
        residuals=res[sel], jacobian=J[sel], valid=valid[sel])
    print(name, "obs", res[sel].shape[0], "invalid", int((valid[sel] == 0).sum()))


def dump_prior_functors():
    """RsConstVeloPrior / RsConstAccelerationPrior (video_bundler_rs_inter.h:55-173) under Jet autodiff: residuals,
    Jacobian w.r.t. the four pose blocks and the d residual / d interFrameRatio column, on fixed inputs."""
    rng = np.random.default_rng(20240917)
    rows = []
    for kind in (1, 2):
        for ratio in (1.0, 0.7, 2.5, 1e-3, 0.0):
            if kind == 2 and ratio == 0.0:
                continue                       # the acceleration functor divides by the ratio
            for _ in range(4):
                fk, fp = rng.normal(0, 0.3, 12), rng.normal(0, 0.3, 12)
                scale = float(rng.uniform(0.1, 30.0))
                ok, r, J, col = oracle.motion_prior_eval_ref(kind, scale, ratio, fk, fp)
                rows.append(dict(kind=kind, ratio=ratio, scale=scale, fk=fk, fp=fp, ok=ok, r=r, J=J, col=col))
    os.makedirs(os.path.join(HERE, "priors"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "priors", "prior_functors.npz"),
                        **{k: np.array([row[k] for row in rows]) for k in rows[0]})
    print("prior_functors", len(rows), "cases")


if __name__ == "__main__":
    oracle.build()
    assert oracle.ref_available(), "needs /root/reference to build oracle/_ref"
    c1 = small_scene()
    dump("c1_first600", c1, keep=np.arange(0, c1.num_obs, 8)[:600])
    for sh in (0, 1, 2):
        for ir in (0, 1):
            dump(f"edge_s{sh}_r{ir}", edge_scene(shutter=sh, interpolate_rotation=bool(ir)))
    dump_prior_functors()
