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


def mat_test_fixtures():
    """The INPUT tables of the reference's own tests (src/rsba/test/mat_test.cc): TEST(SfM, reprojection) poses /
    points / cameras (:170-229) and TEST(SfM, Distortion) cameras / image points (:145-167).  Values only."""
    eps, pi2 = np.finfo(np.float64).eps, np.pi / 2          # _EPS = __DBL_EPSILON__ (mat/core.h:12), M_PI_2
    pose = np.array([[0, 0, 0, 0, 0, 0], [0, 0, 0, 1, 1, 1], [0, 0, pi2, 20, 20, 20], [0, pi2, pi2, -2, 20, 20],
                     [pi2, pi2, pi2, -2, -2, 20], [-1, -1, -1, -2, -2, -2], [-pi2, -1, -1, -20, -2, -2],
                     [0.5, -pi2, -1, -2, -20, -2], [0.5, 0.5, -pi2, 0.2, -2, -20], [eps] * 6, [-eps] * 6], dtype=np.float64)
    pt = np.array([[10, 10, 10], [100, 0, 1], [0, 100, 1], [0, 0, 100], [-100, 0, 1], [0, -100, 1], [0, 0, -100],
                   [0, 0, -1], [0, 0, 0], [1, 1, 1], [-1, -1, -1], [0.1, 0.1, 0.1], [100, 100, 100],
                   [-100, -100, -100], [-0.39, 1.25, 2014], [eps, eps, eps], [eps, eps, -eps], [-eps, -eps, -eps]],
                  dtype=np.float64)
    cam = np.array([[0.1, 0.1, 0, 0, 0, 0, 0, 0, 0], [100, 100, 0, 0, 0, 0, 0, 0, 0], [500, 500, 0, 0, 0, 0, 0, 640, 480],
                    [100, 100, eps, 0, 0, 0, 0, 0, 0], [500, 500, -eps, -eps, 0, 0, 0, 0, 0],
                    [860, 860, 0.001, 0, 0, 0, 0, 100, 200]], dtype=np.float64)
    dcam = np.array([[0.1, 0.1, 0, 0, 0, 0, 0, 0, 0], [100, 100, 0.01, 0, 0, 0, 0, 0, 0],
                     [500, 500, -0.03, 0, 0, 0, 0, 0, 0], [500, 500, -0.1, 0.02, 0, 0, 0, 0, 0]], dtype=np.float64)
    dimg = np.array([[0.10, 0.10], [0.21, 0.19], [1.10, 0.50]], dtype=np.float64)
    return pose, pt, cam, dcam, dimg


def dump_mat_test_grid():
    """What the REFERENCE's own w2c / w2i / distort (mat/cam.h:355-419, 49-72; compiled verbatim in oracle/_ref) return
    on the exact input grid of its own tests: 11 poses x 18 points (x 6 cameras), 4 cameras x 3 image points."""
    import ctypes as C
    lib = oracle.ref_lib()
    dp = C.POINTER(C.c_double)
    p = lambda a: a.ctypes.data_as(dp)  # noqa: E731
    pose, pt, cam, dcam, dimg = mat_test_fixtures()
    w2c = np.zeros((11, 18, 3))
    proj = np.zeros((11, 18, 6, 2))
    ok = np.zeros((11, 18, 6), dtype=np.int32)
    proj_nv = np.zeros((11, 18, 6, 2))            # w2i(validate = false), as solveRSpnp.cpp:65 calls it
    for i in range(11):
        for j in range(18):
            out = np.zeros(3)
            lib.rsba_ref_w2c(p(pose[i]), p(pt[j]), p(out))
            w2c[i, j] = out
            for k in range(6):
                o2 = np.zeros(2)
                ok[i, j, k] = lib.rsba_ref_w2i(p(cam[k]), p(pose[i]), p(pt[j]), p(o2), 1)
                proj[i, j, k] = o2
                if abs(out[2]) > 0.0:             # (z == 0 divides by zero: not a value to pin)
                    o3 = np.zeros(2)
                    lib.rsba_ref_w2i(p(cam[k]), p(pose[i]), p(pt[j]), p(o3), 0)
                    proj_nv[i, j, k] = o3
    dist = np.zeros((4, 3, 2))
    for k in range(4):
        for j in range(3):
            o2 = np.zeros(2)
            lib.rsba_ref_distort(p(dcam[k]), p(dimg[j]), p(o2))
            dist[k, j] = o2
    os.makedirs(os.path.join(HERE, "mat_test"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "mat_test", "grid.npz"), pose=pose, pt=pt, cam=cam, dcam=dcam, dimg=dimg,
                        w2c=w2c, proj=proj, ok=ok, proj_nv=proj_nv, distort=dist)
    print("mat_test/grid: w2i valid on", int(ok.sum()), "of", ok.size)
    # ... and the same grid through the reference's FUNCTOR under Jet autodiff, as ordinary golden scenes (one per
    # camera; global shutter: one pose per frame, stored twice; observation = origin, so residual = projection): the
    # port, the product's host-compiled K1 arithmetic and the GPU kernel are all compared with these
    from rsba_b200.scene import Scene
    fr, pi = np.meshgrid(np.arange(11), np.arange(18), indexing="ij")
    for k in range(6):
        sc = Scene(cam=cam[k], shutter=0, scanlines=np.array([0, 1280], dtype=np.int32), interpolate_rotation=True,
                   poses=np.concatenate([pose, pose], axis=1), points=pt, obs_xy=np.zeros((198, 2)),
                   obs_frame=fr.reshape(-1).astype(np.int32), obs_point=pi.reshape(-1).astype(np.int32),
                   const_frames=np.zeros(11, dtype=bool), name=f"mat_test_cam{k}")
        dump(f"mat_test_cam{k}", sc)


def dump_mat_test_slerp():
    """The 16-rotation table of the reference's TEST(SfM, SLERP) (mat_test.cc:76-94) as the two control poses of
    rolling-shutter frames -- every pair (i, j) the test visits is one frame -- seen through the reference functor under
    Jet autodiff at scan-line times spread over [0, 1] (and clamped outside)."""
    from rsba_b200.scene import Scene
    pi2, eps = np.pi / 2, np.finfo(np.float64).eps
    r = np.array([[0, 0.5, 0], [0, 1, 0], [1, 0, 0], [0, 0, 1], [0, 1, 1], [1, 1, 1], [0, -1, 0], [-1, 0, 0],
                  [-1, 0, -1], [0, pi2, 0], [pi2, 0, 0], [0, 1 - pi2, 0], [0, -1, pi2], [0, -1, 1 - pi2], [0, 0, eps],
                  [1, -1, eps]], dtype=np.float64)
    pairs = [(i, j) for i in range(16) for j in range(1, 16)]
    poses = np.array([np.concatenate([r[i], [0.0, 0.0, 0.0], r[j], [0.1, -0.05, 0.02]]) for i, j in pairs])
    pts = np.array([[10, 10, 10], [0, 0, 100], [1, 1, 1], [-0.39, 1.25, 2014], [100, 100, 100], [0.1, 0.1, 0.1],
                    [-10, 3, 8], [2, -30, -40]], dtype=np.float64)
    xs = np.array([-50.0, 0.0, 1.0, 320.0, 640.0, 1000.0, 1279.0, 1280.0, 1400.0])
    fr, pi = np.meshgrid(np.arange(len(pairs)), np.arange(len(pts)), indexing="ij")
    fr, pi = fr.reshape(-1), pi.reshape(-1)
    xy = np.stack([xs[(fr * 7 + pi * 3) % len(xs)], 100.0 + 37.0 * (pi % 5)], axis=1)
    sc = Scene(cam=np.array([860.0, 860.0, 1e-3, 0, 0, 0, 0, 640.0, 360.0]), shutter=1,
               scanlines=np.array([0, 1280], dtype=np.int32), interpolate_rotation=True, poses=poses, points=pts,
               obs_xy=xy, obs_frame=fr.astype(np.int32), obs_point=pi.astype(np.int32),
               const_frames=np.zeros(len(pairs), dtype=bool), name="mat_test_slerp")
    dump("mat_test_slerp", sc)


if __name__ == "__main__":
    oracle.build()
    assert oracle.ref_available(), "needs /root/reference to build oracle/_ref"
    c1 = small_scene()
    dump("c1_first600", c1, keep=np.arange(0, c1.num_obs, 8)[:600])
    for sh in (0, 1, 2):
        for ir in (0, 1):
            dump(f"edge_s{sh}_r{ir}", edge_scene(shutter=sh, interpolate_rotation=bool(ir)))
    dump_prior_functors()
    dump_mat_test_grid()
    dump_mat_test_slerp()
