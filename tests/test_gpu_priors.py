"""GPU parity for the camera-only motion priors (RsConstVeloPrior /
RsConstAccelerationPrior, video_bundler_rs_inter.h:55-173, wired at CeresHandler.h:148-186) with a
constant interFrameRatio: cost, residuals, the LM step with the priors' J^T J in the camera blocks
and the frame-to-previous-frame couplings of the reduced system, and full solves."""
import numpy as np
import pytest

from rsba_b200.scene import make_scene

pytestmark = pytest.mark.gpu
TOL = 1e-6


@pytest.fixture(scope="module")
def api():
    import rsba_b200.api as api
    api.load_library()
    return api


@pytest.fixture(scope="module")
def lo(oracle_built):
    from oracle import lm_oracle
    return lm_oracle


def relerr(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def chain_priors(sc, kind, scale, ratio, first=1):
    return [(kind, scale, ratio, k, k - 1) for k in range(first, sc.num_frames)]


def load_with_priors(api, pb, sc, priors, huber=0.0):
    pb.load_scene(sc)
    pb.set_motion_priors([p[0] for p in priors], [p[1] for p in priors], [p[2] for p in priors],
                         [p[3] for p in priors], [p[4] for p in priors])
    if huber:
        pb.set_loss(huber)


@pytest.mark.parametrize("kind,scale,ratio", [(1, 5.0, 0.9), (2, 20.0, 1.3), (1, 3.0, 0.0)])
def test_prior_cost_and_residuals(api, oracle_built, kind, scale, ratio):
    sc = make_scene(20, 600, 8, name="priors")
    priors = chain_priors(sc, kind, scale, ratio)
    r0, _, _ = oracle_built.evaluate(sc, impl="port", jac=False)
    _, rx, cost_x = oracle_built.motion_prior_rows(sc, priors)
    with api.Problem(0) as pb:
        load_with_priors(api, pb, sc, priors)
        cost, r, J, v = pb.evaluate()
        got = pb.prior_residuals()
        c2, _ = pb.evaluate_device(with_jacobian=False)
    want = 0.5 * np.sum(r0 * r0) + cost_x
    assert abs(cost - want) <= 1e-12 * want and c2 == cost
    assert np.abs(got.reshape(-1) - rx).max() <= 1e-12 * max(1.0, np.abs(rx).max())
    if oracle_built.ref_available():          # and against the reference functor itself
        ok, r_ref, _, _ = oracle_built.motion_prior_eval_ref(kind, scale, ratio, sc.poses[5], sc.poses[4])
        assert ok and np.abs(got[4] - r_ref).max() <= 1e-12 * max(1.0, np.abs(r_ref).max())


@pytest.mark.parametrize("kind,scale,ratio", [(1, 5.0, 0.9), (2, 20.0, 1.3)])
def test_lm_step_with_priors(api, oracle_built, lo, kind, scale, ratio):
    sc = make_scene(40, 1500, 8, name="priors-band")      # 5 Cholesky tiles: couplings cross tile borders
    priors = chain_priors(sc, kind, scale, ratio)
    r, J, v = oracle_built.evaluate(sc, impl="port")
    Jx, rx, _ = oracle_built.motion_prior_rows(sc, priors)
    want = lo.lm_step(sc, r, J, 1e3, extra=(Jx, rx))
    plain = lo.lm_step(sc, r, J, 1e3)
    assert relerr(plain["delta_poses"], want["delta_poses"]) > 1e-3      # the priors matter
    for kw in (dict(), dict(reorder_tiles=0), dict(dense_cholesky=1)):
        with api.Problem(0) as pb:
            load_with_priors(api, pb, sc, priors)
            got = pb.linearize_and_step(1e3, api.default_options(**kw))
        assert relerr(got["S"], want["S"]) <= TOL
        assert relerr(got["rhs"], want["rhs"]) <= TOL
        assert relerr(got["delta_poses"], want["delta_poses"]) <= TOL
        assert relerr(got["delta_points"], want["delta_points"]) <= TOL
        assert abs(got["model_cost_change"] - want["model_cost_change"]) <= TOL * abs(want["model_cost_change"])


def test_lm_step_priors_with_constant_blocks_and_huber(api, oracle_built, lo):
    sc = make_scene(20, 600, 8, name="priors-masks")
    priors = chain_priors(sc, 1, 8.0, 0.8, first=2)        # frame 1 has no prior; frame 0 constant
    mask = np.zeros(sc.num_frames, dtype=np.uint16)
    mask[0] = 0xFFF
    mask[1] = 0xFFF                                         # CeresHandler.h:182-186 fixes the predecessor too
    mask[7] = 0b000111 | (0b000111 << 6)
    a = 1.5
    r, J, v = oracle_built.evaluate(sc, impl="port")
    rw, Jw, _ = oracle_built.apply_huber(r, J, a)
    Jx, rx, _ = oracle_built.motion_prior_rows(sc, priors, huber=a)
    want = lo.lm_step(sc, rw, Jw, 1e2, pose_mask=mask, extra=(Jx, rx))
    with api.Problem(0) as pb:
        pb.set_camera(sc.cam, sc.shutter, sc.scanlines, sc.interpolate_rotation)
        pb.set_scene(sc.obs_xy, sc.obs_frame, sc.obs_point, sc.num_frames, sc.num_points, mask)
        pb.set_parameters(sc.poses, sc.points)
        pb.set_motion_priors([p[0] for p in priors], [p[1] for p in priors], [p[2] for p in priors],
                             [p[3] for p in priors], [p[4] for p in priors])
        pb.set_loss(a)
        got = pb.linearize_and_step(1e2)
    for k in ("S", "rhs", "delta_poses", "delta_points"):
        assert relerr(got[k], want[k]) <= TOL, k
    assert not got["delta_poses"][0].any() and not got["delta_poses"][1].any()


def test_solve_with_priors_bulk_and_pointer_api(api, oracle_built):
    sc = make_scene(16, 400, 8, name="priors-solve")
    priors = chain_priors(sc, 2, 10.0, 1.2)
    with api.Problem(0) as pb:
        load_with_priors(api, pb, sc, priors)
        s = pb.solve(api.default_options(max_num_iterations=12))
        po, pt = pb.get_parameters()
    assert s.usable == 1 and s.final_cost < s.initial_cost
    r1, _, _ = oracle_built.evaluate(sc, po, pt, jac=False, impl="port")
    _, _, cx = oracle_built.motion_prior_rows(sc, priors, poses=po)
    want = 0.5 * np.sum(r1 * r1) + cx
    assert abs(s.final_cost - want) <= 1e-9 * want
    # pointer API: AddResidualBlock order of CeresHandler::Add -- prior first, then the observations
    poses, points = sc.poses.copy(), sc.points.copy()
    with api.Problem(0) as pb:
        pb.set_camera(sc.cam, sc.shutter, sc.scanlines, sc.interpolate_rotation)
        for f in range(sc.num_frames):
            if f > 0:
                pb.add_motion_prior(2, 10.0, 1.2, poses[f, :6], poses[f, 6:], poses[f - 1, :6], poses[f - 1, 6:])
            for i in np.flatnonzero(sc.obs_frame == f):
                pb.add_rs_residual(sc.obs_xy[i], poses[f, :6], poses[f, 6:], points[int(sc.obs_point[i])])
        pb.set_block_constant(poses[0, :6])
        pb.set_block_constant(poses[0, 6:])
        s2 = pb.solve(api.default_options(max_num_iterations=12))
    assert abs(s2.final_cost - s.final_cost) <= 1e-9 * s.final_cost
    assert relerr(poses, po) <= 1e-7


def test_prior_argument_checks(api):
    sc = make_scene(6, 100, 5, name="priors-args")
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        with pytest.raises(api.RsbaError):
            pb.set_motion_priors([1], [1.0], [-0.5], [1], [0])      # functor would return false
        with pytest.raises(api.RsbaError):
            pb.set_motion_priors([2], [1.0], [0.0], [1], [0])
        with pytest.raises(api.RsbaError):
            pb.set_motion_priors([1, 1], [1.0, 1.0], [1.0, 1.0], [2, 2], [1, 0])   # two priors on one frame
        pb.set_motion_priors([], [], [], [], [])
        cost, _ = pb.evaluate_device(with_jacobian=False)
        assert cost > 0


# ---------------------------------------------------------------------------------------------
# free interFrameRatio: the reference's default (interFrameRatio == 1 leaves `&opt.ceres.interFrameRatio`
# a variable, lower-bounded block shared by every prior, CeresHandler.h:156-180)
def free_ratio_step(oracle, lo, sc, priors, ratio, radius, free_cam=False, huber=0.0):
    r, J, v = oracle.evaluate(sc, impl="port")
    Jc = oracle.intrinsics_jacobian(sc) if free_cam else None
    if huber:
        if free_cam:   # the Corrector rescales every block of the row, the intrinsics columns too
            Jc = oracle.apply_huber(r, Jc, huber)[1]
        r, J, _ = oracle.apply_huber(r, J, huber)
    sc2, mask = lo.with_pseudo_frame(sc, ratio=ratio, free_cam=free_cam)
    Jx, rx, cx = oracle.motion_prior_rows(sc2, priors, huber=huber, free_ratio=ratio)
    cols = None
    if free_cam:
        _, _, cols = lo.with_intrinsics_block(sc, J, Jc)
    want = lo.lm_step(sc2, r, J, radius, pose_mask=mask, extra=(Jx, rx), cam_cols=cols)
    want["prior_cost"] = cx
    return want


@pytest.mark.parametrize("kind,scale,ratio", [(1, 5.0, 1.0), (2, 20.0, 1.0), (1, 4.0, 0.6), (2, 9.0, 1.7)])
def test_free_ratio_column_and_cost(api, oracle_built, kind, scale, ratio):
    sc = make_scene(20, 600, 8, name="ratio")
    priors = chain_priors(sc, kind, scale, 123.0)      # the per-prior value is ignored with a free ratio
    with api.Problem(0) as pb:
        load_with_priors(api, pb, sc, priors)
        pb.set_inter_frame_ratio_free(True, ratio)
        cost, r, J, v = pb.evaluate()
        got_r, got_j = pb.prior_residuals(), pb.prior_ratio_jacobian()
        assert pb.inter_frame_ratio() == ratio
    r0, _, _ = oracle_built.evaluate(sc, impl="port", jac=False)
    fixed = chain_priors(sc, kind, scale, ratio)
    _, rx, cx = oracle_built.motion_prior_rows(sc, fixed)
    want = 0.5 * np.sum(r0 * r0) + cx
    assert abs(cost - want) <= 1e-12 * want
    assert np.abs(got_r.reshape(-1) - rx).max() <= 1e-12 * max(1.0, np.abs(rx).max())
    for i, (_, _, _, fk, fp) in enumerate(fixed):
        col = oracle_built.motion_prior_ratio_column(kind, scale, ratio, sc.poses[fk], sc.poses[fp])
        assert np.abs(got_j[i] - col).max() <= 1e-12 * max(1.0, np.abs(col).max())
    if oracle_built.ref_available():          # and against the reference functor's own Jet column
        ok, _, _, col_ref = oracle_built.motion_prior_eval_ref(kind, scale, ratio, sc.poses[5], sc.poses[4])
        assert ok and np.abs(got_j[4] - col_ref).max() <= 1e-12 * max(1.0, np.abs(col_ref).max())


@pytest.mark.parametrize("frames,kind,scale,ratio", [(13, 1, 5.0, 1.0), (40, 2, 20.0, 1.0), (40, 1, 6.0, 0.8)])
def test_lm_step_free_ratio(api, oracle_built, lo, frames, kind, scale, ratio):
    """frames = 13: the pseudo-frame shares its sub-tile and tile with real frames; 40: five tiles + border."""
    sc = make_scene(frames, 40 * frames, 8, name=f"ratio{frames}")
    priors = chain_priors(sc, kind, scale, ratio)
    want = free_ratio_step(oracle_built, lo, sc, priors, ratio, 1e3)
    fixed = lo.lm_step(sc, *oracle_built.evaluate(sc, impl="port")[:2], 1e3,
                       extra=oracle_built.motion_prior_rows(sc, priors)[:2])
    assert abs(want["delta_poses"][frames, 9]) > 1e-6                       # the ratio moves ...
    assert relerr(want["delta_poses"][:frames], fixed["delta_poses"]) > 1e-6  # ... and that matters
    for kw in (dict(), dict(reorder_tiles=0), dict(dense_cholesky=1)):
        with api.Problem(0) as pb:
            load_with_priors(api, pb, sc, priors)
            pb.set_inter_frame_ratio_free(True, ratio)
            got = pb.linearize_and_step(1e3, api.default_options(**kw))
        for k in ("S", "rhs", "delta_poses", "delta_points"):
            assert relerr(got[k], want[k]) <= TOL, (kw, k)
        assert abs(got["model_cost_change"] - want["model_cost_change"]) <= TOL * abs(want["model_cost_change"])
        d = got["delta_poses"][frames]
        assert abs(d[9] - want["delta_poses"][frames, 9]) <= TOL * abs(want["delta_poses"][frames, 9])
        assert not d[:9].any() and not d[10:].any()


def test_lm_step_free_ratio_with_free_intrinsics_and_huber(api, oracle_built, lo):
    """Both tenants of the pseudo-frame at once: intrinsics (0..8) and the ratio (9)."""
    sc = make_scene(21, 800, 8, name="ratio-uncal")
    priors = chain_priors(sc, 2, 12.0, 1.0)
    want = free_ratio_step(oracle_built, lo, sc, priors, 1.1, 5e2, free_cam=True, huber=1.5)
    with api.Problem(0) as pb:
        pb.set_intrinsics_free(True)
        load_with_priors(api, pb, sc, priors, huber=1.5)
        pb.set_inter_frame_ratio_free(True, 1.1)
        got = pb.linearize_and_step(5e2)
    for k in ("S", "rhs", "delta_poses", "delta_points"):
        assert relerr(got[k], want[k]) <= TOL, k
    assert got["delta_poses"][21, :10].all() and not got["delta_poses"][21, 10:].any()


def test_solve_free_ratio_bulk_and_pointer_api(api, oracle_built):
    sc = make_scene(16, 400, 8, name="ratio-solve")
    priors = chain_priors(sc, 1, 10.0, 1.0)
    with api.Problem(0) as pb:
        load_with_priors(api, pb, sc, priors)
        s_fixed = pb.solve(api.default_options(max_num_iterations=12))
    with api.Problem(0) as pb:
        load_with_priors(api, pb, sc, priors)
        pb.set_inter_frame_ratio_free(True, 1.0)
        s = pb.solve(api.default_options(max_num_iterations=12))
        po, pt = pb.get_parameters()
        ratio = pb.inter_frame_ratio()
    assert s.usable == 1 and s.final_cost < s.initial_cost
    assert s.final_cost < s_fixed.final_cost and abs(ratio - 1.0) > 1e-4 and ratio > 0      # one more degree of freedom
    # Ceres 1.9 follows the projected step of a bounded problem with an Armijo line search that starts at step size 1;
    # the loop here has none: on this scene the search would have accepted every step as it is (a no-op)
    assert s.num_armijo_violations == 0 and s_fixed.num_armijo_violations == 0
    assert s.num_parameters_reduced == s_fixed.num_parameters_reduced + 1
    r1, _, _ = oracle_built.evaluate(sc, po, pt, jac=False, impl="port")
    _, _, cx = oracle_built.motion_prior_rows(sc, chain_priors(sc, 1, 10.0, ratio), poses=po)
    want = 0.5 * np.sum(r1 * r1) + cx
    assert abs(s.final_cost - want) <= 1e-9 * want
    # pointer API: the ratio is the caller's scalar block, written back in place
    poses, points, rblock = sc.poses.copy(), sc.points.copy(), np.array([1.0])
    with api.Problem(0) as pb:
        pb.set_camera(sc.cam, sc.shutter, sc.scanlines, sc.interpolate_rotation)
        pb.set_inter_frame_ratio_block(rblock)
        for f in range(sc.num_frames):
            if f > 0:
                pb.add_motion_prior(1, 10.0, 1.0, poses[f, :6], poses[f, 6:], poses[f - 1, :6], poses[f - 1, 6:])
            for i in np.flatnonzero(sc.obs_frame == f):
                pb.add_rs_residual(sc.obs_xy[i], poses[f, :6], poses[f, 6:], points[int(sc.obs_point[i])])
        pb.set_block_constant(poses[0, :6])
        pb.set_block_constant(poses[0, 6:])
        s2 = pb.solve(api.default_options(max_num_iterations=12))
    assert abs(s2.final_cost - s.final_cost) <= 1e-9 * s.final_cost
    assert relerr(poses, po) <= 1e-7 and abs(rblock[0] - ratio) <= 1e-9


def test_free_ratio_stays_on_its_lower_bound(api):
    """SetParameterLowerBound (CeresHandler.h:161,172): the trial point is projected onto the bound, so the
    functor never sees a ratio it would reject (video_bundler_rs_inter.h:92,157) and the solve stays usable."""
    sc = make_scene(12, 300, 8, name="ratio-bound")
    for kind in (1, 2):
        with api.Problem(0) as pb:
            load_with_priors(api, pb, sc, chain_priors(sc, kind, 50.0, 1.0))
            pb.set_inter_frame_ratio_free(True, 1e-3)
            s = pb.solve(api.default_options(max_num_iterations=15))
            ratio = pb.inter_frame_ratio()
        assert s.usable == 1 and np.isfinite(s.final_cost) and s.final_cost <= s.initial_cost
        assert ratio >= (0.0 if kind == 1 else np.finfo(np.float64).eps)
    with api.Problem(0) as pb:
        load_with_priors(api, pb, sc, chain_priors(sc, 2, 50.0, 1.0))
        with pytest.raises(api.RsbaError):
            pb.set_inter_frame_ratio_free(True, 0.0)        # below the acceleration prior's bound
