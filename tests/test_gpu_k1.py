"""GPU parity tests for K1 (residual + Jacobian) and K1r (cost only), through the C ABI.

Bar (BASELINE.json north_star): residuals and Jacobians within 1e-6 relative of the
reference CPU path on identical inputs; valid bits exact.  The checker is the oracle
(port; oracle/_ref too when its prebuilt .so travelled) and the committed golden vectors.
"""
import os

import numpy as np
import pytest

from conftest import rel_block_err
from helpers import edge_scene, shuffled, small_scene
from test_oracle_cpu import GOLDEN, scene_from_golden

pytestmark = pytest.mark.gpu

TOL = 1e-6  # relative, north_star


@pytest.fixture(scope="module")
def api():
    import rsba_b200.api as api
    api.load_library()
    return api


def gpu_eval(api, sc, **kw):
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        return pb.evaluate(check=False, **kw)


def check(sc, got, want, tol=TOL):
    cost, r, J, v = got
    r0, J0, v0 = want
    assert np.array_equal(v, v0), "valid bits differ"
    ok = v0 == 1
    assert (np.abs(r - r0) / np.maximum(1.0, np.abs(r0)))[ok].max(initial=0) <= tol
    if J is not None:
        assert rel_block_err(J[ok], J0[ok]).max(initial=0) <= tol
        assert not J[~ok].any()
    assert not r[~ok].any()
    want_cost = 0.5 * np.sum(r0[ok] ** 2)
    assert abs(cost - want_cost) <= 1e-9 * max(1.0, abs(want_cost))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_k1_matches_reference_golden(api, path):
    sc, g = scene_from_golden(path)
    check(sc, gpu_eval(api, sc), (g["residuals"], g["jacobian"], g["valid"]))


def test_k1_c1_vs_oracle(api, oracle_built):
    sc = small_scene()
    want = oracle_built.evaluate(sc, impl="port")
    check(sc, gpu_eval(api, sc), want)
    if oracle_built.ref_available():
        check(sc, gpu_eval(api, sc), oracle_built.evaluate(sc, impl="ref"))


@pytest.mark.parametrize("shutter,interp", [(0, True), (1, True), (1, False), (2, True)])
def test_k1_edge_cases_vs_oracle(api, oracle_built, shutter, interp):
    sc = edge_scene(shutter, interp)
    check(sc, gpu_eval(api, sc), oracle_built.evaluate(sc, impl="port"))


def test_k1_unsorted_input_is_reported_in_caller_order(api, oracle_built):
    sc = shuffled(small_scene())
    check(sc, gpu_eval(api, sc), oracle_built.evaluate(sc, impl="port"))


def test_k1_ragged_and_empty(api, oracle_built):
    from rsba_b200.scene import Scene
    sc = small_scene()
    for n in (1, 31, 32, 33, 127, 129, 1000):
        sub = Scene(**{**sc.__dict__, "obs_xy": sc.obs_xy[:n], "obs_frame": sc.obs_frame[:n],
                       "obs_point": sc.obs_point[:n]})
        check(sub, gpu_eval(api, sub), oracle_built.evaluate(sub, impl="port"))
    with api.Problem(0) as pb:                       # empty problem: cost 0, nothing written
        pb.set_camera(sc.cam, sc.shutter, sc.scanlines, True)
        pb.set_scene(np.zeros((0, 2)), np.zeros(0, np.int32), np.zeros(0, np.int32), sc.num_frames, sc.num_points)
        pb.set_parameters(sc.poses, sc.points)
        cost, r, J, v = pb.evaluate()
        assert cost == 0.0 and r.shape == (0, 2)


def test_k1_wide_frame_span_uses_global_pose_path(api, oracle_built):
    """More than 8 frames inside one 128-observation CTA: poses come through L1, not smem."""
    from rsba_b200.scene import make_scene
    sc = make_scene(64, 64, 40, name="thin")         # 40 obs per frame
    check(sc, gpu_eval(api, sc), oracle_built.evaluate(sc, impl="port"))


def test_k1r_cost_matches_k1(api, oracle_built):
    sc = small_scene()
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        c_full, bad_full = pb.evaluate_device(with_jacobian=True)
        c_res, bad_res = pb.evaluate_device(with_jacobian=False)
    r0, _, _ = oracle_built.evaluate(sc, impl="port", jac=False)
    assert bad_full == 0 and bad_res == 0
    assert c_full == c_res
    assert abs(c_res - 0.5 * np.sum(r0 ** 2)) <= 1e-9 * c_res


def test_k1_pointer_api_matches_bulk(api, oracle_built):
    """AddResidualBlock-style construction (CeresHandler.h:250-255): block identity is the
    pointer; insertion order frame-major; results identical to the bulk path."""
    sc = small_scene()
    n = 1500
    poses, points = sc.poses.copy(), sc.points.copy()
    with api.Problem(0) as pb:
        pb.set_camera(sc.cam, sc.shutter, sc.scanlines, sc.interpolate_rotation)
        used_pts = {}
        for i in range(n):
            f, p = int(sc.obs_frame[i]), int(sc.obs_point[i])
            pb.add_rs_residual(sc.obs_xy[i], poses[f, :6], poses[f, 6:], points[p])
            used_pts[p] = 1
        pb.set_block_constant(poses[0, :6])
        pb.set_block_constant(poses[0, 6:])
        cost, r, J, v = pb.evaluate(num_obs=n)
    from rsba_b200.scene import Scene
    sub = Scene(**{**sc.__dict__, "obs_xy": sc.obs_xy[:n], "obs_frame": sc.obs_frame[:n], "obs_point": sc.obs_point[:n]})
    check(sub, (cost, r, J, v), oracle_built.evaluate(sub, impl="port"))


def test_k1_full_size_properties(api, oracle_built):
    """BASELINE config C2 (100 frames / 20k points / 500k obs): spot-check 4 096 random
    observations against the oracle, the F3 structure J_pose1*(1-tau) == J_pose0*tau on all of
    them, and cost == 1/2 sum r^2 of the returned residuals."""
    from rsba_b200.scene import Scene, make_config
    sc = make_config("C2")
    cost, r, J, v = gpu_eval(api, sc)
    assert v.all()
    assert abs(cost - 0.5 * np.sum(r * r)) <= 1e-10 * cost
    tau = np.clip(sc.obs_xy[:, 0] / 1280.0, 0, 1)[:, None]
    assert np.allclose(J[:, 12:24] * (1 - tau), J[:, :12] * tau, rtol=1e-9, atol=1e-9)
    idx = np.sort(np.random.default_rng(5).choice(sc.num_obs, 4096, replace=False))
    sub = Scene(**{**sc.__dict__, "obs_xy": sc.obs_xy[idx], "obs_frame": sc.obs_frame[idx], "obs_point": sc.obs_point[idx]})
    r0, J0, v0 = oracle_built.evaluate(sub, impl="port")
    assert (np.abs(r[idx] - r0) / np.maximum(1.0, np.abs(r0))).max() <= TOL
    assert rel_block_err(J[idx], J0).max() <= TOL
    if oracle_built.ref_available():
        # every one of the 500 000 observations against the reference's own functor under Jet<15>
        rr, Jr, vr = oracle_built.evaluate(sc, impl="ref")
        assert np.array_equal(v, vr)
        assert (np.abs(r - rr) / np.maximum(1.0, np.abs(rr))).max() <= TOL
        assert rel_block_err(J, Jr).max() <= TOL


@pytest.mark.parametrize("a", [0.5, 2.0])
def test_k1_huber_loss_matches_ceres_corrector(api, oracle_built, a):
    """ceres::HuberLoss(a) on every residual block (CeresHandler.h:85-90): cost = 1/2 sum rho(|r|^2),
    residuals and Jacobians rescaled by sqrt(rho') -- what problem.Evaluate returns."""
    sc = small_scene()
    xy = sc.obs_xy.copy()
    xy[::9] += 25.0                                   # gross outliers
    from rsba_b200.scene import Scene
    sc = Scene(**{**sc.__dict__, "obs_xy": xy})
    r0, J0, v0 = oracle_built.evaluate(sc, impl="port")
    rw, Jw, cost_w = oracle_built.apply_huber(r0, J0, a)
    assert (np.sum(r0 * r0, axis=1) > a * a).sum() > 100
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        pb.set_loss(a)
        cost, r, J, v = pb.evaluate()
        c_res, _ = pb.evaluate_device(with_jacobian=False)
        pb.set_loss(0.0)
        cost_plain, _, _, _ = pb.evaluate()
    assert np.array_equal(v, v0)
    assert abs(cost - cost_w) <= 1e-12 * cost_w and c_res == cost
    assert (np.abs(r - rw) / np.maximum(1.0, np.abs(rw))).max() <= TOL
    assert rel_block_err(J, Jw).max() <= TOL
    assert abs(cost_plain - 0.5 * np.sum(r0 * r0)) <= 1e-12 * cost_plain and cost < cost_plain


@pytest.mark.parametrize("shutter", [1, 2, 0])
def test_validate_sweep_matches_reference_predicate(api, oracle_built, shutter):
    """validate() swept over every observation (struct/VideoSfM.cc:159-169, VideoSfMHandler.cc:377-410):
    scan line from x or y by shutter direction, distance gate, squared-error threshold."""
    sc = edge_scene(shutter, True)
    want_ok, want_err = oracle_built.validate_sweep(sc, sqrd_threshold=4e5, min_distance=3.5)
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        ok, err = pb.validate(sqrd_threshold=4e5, min_distance_to_camera=3.5)
    assert np.array_equal(ok, want_ok)
    assert 0 < ok.sum() < ok.size
    good = want_err >= 0
    assert np.array_equal(err < 0, ~good)
    assert (np.abs(err[good] - want_err[good]) <= 1e-9 * np.maximum(1.0, want_err[good])).all()
    sc2 = small_scene()                                # a clean scene validates at the default threshold
    with api.Problem(0) as pb:
        pb.load_scene(sc2, poses=sc2.poses_true, points=sc2.points_true)
        ok2, err2 = pb.validate()
    assert ok2.all() and err2.max() < 16.0


@pytest.mark.parametrize("shutter", [1, 2, 0])
def test_iterative_reprojection_matches_reference_loop(api, oracle_built, shutter):
    """reproject() (struct/VideoSfM.cc:139-155): fixed point on the scan line from the principal point, limit
    and failure rules of the reference, for every (frame, point) pair of the scene plus pairs the frame never
    observed (some of which lie behind the camera)."""
    sc = edge_scene(shutter, True)
    rng = np.random.default_rng(11)
    frame = np.concatenate([sc.obs_frame, rng.integers(0, sc.num_frames, 200)]).astype(np.int32)
    point = np.concatenate([sc.obs_point, rng.integers(0, sc.num_points, 200)]).astype(np.int32)
    want_xy, want_ok = oracle_built.reproject_sweep(sc, frame, point)
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        xy, ok = pb.reproject(frame, point)
        _, ok0 = pb.reproject(frame, point, sqrd_threshold=0.0)
    assert np.array_equal(ok, want_ok) and not ok0.any()
    good = want_ok == 1
    assert good.sum() > 100
    assert (np.abs(xy[good] - want_xy[good]) <= 1e-9 * np.maximum(1280.0, np.abs(want_xy[good]))).all()
    # at the true parameters a re-projection lands on the (noisy, sigma = 0.5 px) observation it came from
    sc2 = small_scene()
    with api.Problem(0) as pb:
        pb.load_scene(sc2, poses=sc2.poses_true, points=sc2.points_true)
        xy2, ok2 = pb.reproject(sc2.obs_frame, sc2.obs_point)
    d = np.linalg.norm(xy2 - sc2.obs_xy, axis=1)
    assert ok2.all() and np.median(d) < 1.0 and d.max() < 4.0


@pytest.mark.parametrize("shutter,interp", [(1, True), (2, False), (0, True)])
def test_k1_uncalibrated_intrinsics_jacobian(api, oracle_built, shutter, interp):
    """<2; 9, 6, 6, 3> (RsBundleAdjustment::CreateWithCam, VideoSfmBaRs.h:38-49,68-80): residual, pose/point
    Jacobian unchanged, plus d residual / d (fx fy k1 k2 p1 p2 k3 cx cy) against the reference's own
    ReprojectionError::operator()(camera, ...) under Jet<24> and the numpy closed form."""
    sc = edge_scene(shutter, interp)
    with api.Problem(0) as pb:
        pb.set_intrinsics_free(True)
        pb.load_scene(sc)
        cost, r, J, v = pb.evaluate(check=False)
        Jc = pb.intrinsics_jacobian()
        assert np.array_equal(pb.get_camera(), sc.cam)
    r0, J0, v0 = oracle_built.evaluate(sc, impl="port")
    check(sc, (cost, r, J, v), (r0, J0, v0))
    want = oracle_built.intrinsics_jacobian(sc)
    ok = v0 == 1
    assert rel_block_err(Jc[ok], want[ok]).max() <= TOL and not Jc[~ok].any()
    if oracle_built.ref_available():
        _, _, Jref, vref = oracle_built.evaluate_cam_ref(sc)
        assert np.array_equal(v, vref) and rel_block_err(Jc[ok], Jref[ok]).max() <= TOL


@pytest.mark.gpu
def test_residual_only_evaluate_skips_the_jacobian_kernel_and_agrees():
    """rsba_cuda_evaluate without a Jacobian (problem.Evaluate(&cost, &residuals), CeresHandler.h:386) runs the cost-only
    kernel -- no 240 bytes per observation are allocated or written: the device Jacobian buffer stays NULL -- and gives
    the residuals, valid flags and cost of the full evaluation bit for bit, Huber loss included."""
    import rsba_b200.api as api
    from rsba_b200.scene import make_scene
    sc = make_scene(20, 1500, 10, name="res-only")
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        pb.set_loss(2.0)
        c0, r0, _, v0 = pb.evaluate(jacobian=False)
        assert pb.device_buffers()["jacobian"] is None          # nothing allocated for it
        c1, r1, J1, v1 = pb.evaluate()
        assert pb.device_buffers()["jacobian"] is not None
    assert c0 == c1 and np.array_equal(r0, r1) and np.array_equal(v0, v1) and np.abs(J1).max() > 0
