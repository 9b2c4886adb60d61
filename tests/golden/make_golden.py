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


if __name__ == "__main__":
    oracle.build()
    assert oracle.ref_available(), "needs /root/reference to build oracle/_ref"
    c1 = small_scene()
    dump("c1_first600", c1, keep=np.arange(0, c1.num_obs, 8)[:600])
    for sh in (0, 1, 2):
        for ir in (0, 1):
            dump(f"edge_s{sh}_r{ir}", edge_scene(shutter=sh, interpolate_rotation=bool(ir)))
