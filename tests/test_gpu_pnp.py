"""Batched RS-PnP: rsba_cuda_pnp_batch against (a) rsba_cuda_solve on the
equivalent one-frame problem with constant points -- the same ceres::Solve that vision::solveRsPnP runs
per hypothesis (solveRSpnp.cpp:100-192), already parity-tested against the oracle -- (b) the numpy
restatement of the LM loop directly, and (c) the port's w2i(validate = false) for the inlier scoring
(solveRSpnp.cpp:226-263, 318-323)."""
import ctypes as C

import numpy as np
import pytest

from rsba_b200.scene import Scene, make_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import rsba_b200.api as api
    api.load_library()
    return api


def frame_problem(shutter=1, frame=5, seed=0):
    """2-D/3-D correspondences of one frame of a synthetic scene + a perturbed initial pose."""
    sc = make_scene(12, 1200, 8, name="pnp", shutter=shutter)
    sel = np.flatnonzero(sc.obs_frame == frame)
    pts = sc.points_true[sc.obs_point[sel]].astype(np.float32).astype(np.float64)   # through float, as the reference
    xy = sc.obs_xy[sel].astype(np.float32).astype(np.float64)
    rng = np.random.default_rng(seed)
    pose0 = sc.poses_true[frame] + rng.normal(0, 2e-3, 12)
    return sc, pts, xy, pose0


def one_frame_scene(sc, pts, xy, pose, idx):
    k = len(idx)
    return Scene(cam=sc.cam, shutter=sc.shutter, scanlines=sc.scanlines, interpolate_rotation=True,
                 poses=pose.reshape(1, 12).copy(), points=pts[idx].copy(), obs_xy=xy[idx].copy(),
                 obs_frame=np.zeros(k, np.int32), obs_point=np.arange(k, dtype=np.int32),
                 const_frames=np.zeros(1, bool), name="pnp1")


@pytest.mark.parametrize("shutter", [1, 2, 0])
def test_pnp_batch_matches_single_solves(api, oracle_built, shutter):
    from oracle import lm_oracle as lo
    sc, pts, xy, pose0 = frame_problem(shutter)
    n = pts.shape[0]
    assert n > 100
    rng = np.random.default_rng(7)
    H, k = 48, 6
    idx = np.stack([rng.choice(n, k, replace=False) for _ in range(H)]).astype(np.int32)
    idx[0] = np.arange(k)                                   # deterministic first sample
    poses = np.tile(pose0, (H, 1))
    with api.Problem(0) as pb:
        got = pb.pnp_batch(sc.cam, sc.shutter, sc.scanlines, pts, xy, idx, poses, inlier_threshold=3.0)
    assert got["usable"].all()
    # (a) the same hypothesis through rsba_cuda_solve: one frame, constant points
    for hidx in (0, 5, 17, 40):
        one = one_frame_scene(sc, pts, xy, pose0, idx[hidx])
        with api.Problem(0) as pb:
            pb.set_camera(one.cam, one.shutter, one.scanlines, True)
            pb.set_scene(one.obs_xy, one.obs_frame, one.obs_point, 1, k, None, np.ones(k, np.uint8))
            pb.set_parameters(one.poses, one.points)
            s = pb.solve(api.default_options(max_num_iterations=10))
            po, _ = pb.get_parameters()
        assert s.iterations == got["iterations"][hidx]
        assert abs(s.final_cost - got["cost"][hidx]) <= 1e-7 * max(s.final_cost, 1e-12) + 1e-18
        assert np.linalg.norm(po[0] - got["poses"][hidx]) <= 1e-7 * np.linalg.norm(po[0])
    # (b) the numpy LM loop on hypothesis 0
    one = one_frame_scene(sc, pts, xy, pose0, idx[0])
    ev = lambda po, pt, jac: oracle_built.evaluate(one, po, pt, jac=jac, impl="port")  # noqa: E731
    po, _, summ = lo.solve(one, ev, lo.Options(max_num_iterations=10), point_const=np.ones(k, np.uint8))
    assert abs(summ.final_cost - got["cost"][0]) <= 1e-6 * max(summ.final_cost, 1e-12) + 1e-18
    assert np.linalg.norm(po[0] - got["poses"][0]) <= 1e-6 * np.linalg.norm(po[0])
    # (c) inlier scoring with the port's primitives: scan line from the observation itself, no z test
    lib = oracle_built.port_lib()
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.rsba_oracle_interpolate_rs.argtypes = [dp, dp, C.c_int, ip, dp, dp, C.c_int]
    lib.rsba_oracle_w2i.argtypes = [dp, dp, dp, dp, C.c_int]
    cam = np.ascontiguousarray(sc.cam, dtype=np.float64)
    scan = np.ascontiguousarray(sc.scanlines, dtype=np.int32)
    for hidx in (0, 17):
        p = got["poses"][hidx]
        p0, p1 = np.ascontiguousarray(p[:6]), np.ascontiguousarray(p[6:])
        cnt = 0
        pose, proj = np.zeros(6), np.zeros(2)
        for i in range(n):
            o = np.ascontiguousarray(xy[i])
            X = np.ascontiguousarray(pts[i])
            lib.rsba_oracle_interpolate_rs(p0.ctypes.data_as(dp), p1.ctypes.data_as(dp), int(sc.shutter),
                                           scan.ctypes.data_as(ip), o.ctypes.data_as(dp), pose.ctypes.data_as(dp), 1)
            lib.rsba_oracle_w2i(cam.ctypes.data_as(dp), pose.ctypes.data_as(dp), X.ctypes.data_as(dp),
                                proj.ctypes.data_as(dp), 0)
            cnt += np.linalg.norm(o - proj) < 3.0
        assert cnt == got["inliers"][hidx]
    assert got["inliers"].max() > 0.5 * n                   # clean data: good hypotheses explain most points


def test_pnp_batch_argument_checks_and_outliers(api):
    sc, pts, xy, pose0 = frame_problem(1)
    n = pts.shape[0]
    xy_bad = xy.copy()
    xy_bad[::3] += 60.0                                     # a third of the matches are wrong
    rng = np.random.default_rng(1)
    idx = np.stack([rng.choice(n, 6, replace=False) for _ in range(256)]).astype(np.int32)
    with api.Problem(0) as pb:
        got = pb.pnp_batch(sc.cam, sc.shutter, sc.scanlines, pts, xy_bad, idx, np.tile(pose0, (256, 1)),
                           inlier_threshold=3.0)
        best = int(np.argmax(got["inliers"]))
        assert got["inliers"][best] > 0.55 * n              # RANSAC finds an all-inlier sample
        assert got["inliers"].min() < got["inliers"][best]
        with pytest.raises(api.RsbaError):
            pb.pnp_batch(sc.cam, sc.shutter, sc.scanlines, pts, xy, np.full((2, 6), n, np.int32), np.tile(pose0, (2, 1)))
        with pytest.raises(api.RsbaError):
            pb.pnp_batch(sc.cam, sc.shutter, sc.scanlines, pts, xy, np.zeros((2, 40), np.int32), np.tile(pose0, (2, 1)))
