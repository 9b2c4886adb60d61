"""GPU parity tests for K2 (normal equations + Schur complement), K3 (tile Cholesky), K4
(back-substitution / step) and the full LM loop, through the C ABI, against the numpy
restatement of Ceres 1.9.0's LM step (oracle/lm_oracle.py; parity unpinned by the reference,
SURVEY 8c).  Bar: reduced system, step and model cost change within 1e-6 relative."""
import numpy as np
import pytest

from helpers import small_scene
from rsba_b200.scene import Scene, make_scene

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


def oracle_step(oracle, lo, sc, radius, pose_mask=None, point_const=None, **kw):
    r, J, v = oracle.evaluate(sc, impl="port")
    assert v.all()
    return lo.lm_step(sc, r, J, radius, lo.Options(**kw), None, pose_mask, point_const)


def gpu_step(api, sc, radius, pose_mask=None, point_const=None, **kw):
    with api.Problem(0) as pb:
        pb.set_camera(sc.cam, sc.shutter, sc.scanlines, sc.interpolate_rotation)
        mask = np.where(sc.const_frames, 0xFFF, 0).astype(np.uint16) if pose_mask is None else pose_mask
        pb.set_scene(sc.obs_xy, sc.obs_frame, sc.obs_point, sc.num_frames, sc.num_points, mask, point_const)
        pb.set_parameters(sc.poses, sc.points)
        return pb.linearize_and_step(radius, api.default_options(**kw))


def compare(got, want, tol=TOL):
    assert relerr(got["S"], want["S"]) <= tol
    assert relerr(got["rhs"], want["rhs"]) <= tol
    assert relerr(got["delta_poses"], want["delta_poses"]) <= tol
    assert relerr(got["delta_points"], want["delta_points"]) <= tol
    assert abs(got["model_cost_change"] - want["model_cost_change"]) <= tol * abs(want["model_cost_change"])


@pytest.mark.parametrize("radius", [1e4, 1.0, 1e-3])
def test_lm_step_c1(api, oracle_built, lo, radius):
    sc = small_scene()
    compare(gpu_step(api, sc, radius), oracle_step(oracle_built, lo, sc, radius))


def test_lm_step_dense_cholesky_same_result(api, oracle_built, lo):
    sc = small_scene()
    want = oracle_step(oracle_built, lo, sc, 1e4)
    compare(gpu_step(api, sc, 1e4, dense_cholesky=1), want)


def test_lm_step_no_jacobi_scaling(api, oracle_built, lo):
    sc = small_scene()
    compare(gpu_step(api, sc, 1e2, jacobi_scaling=0), oracle_step(oracle_built, lo, sc, 1e2, jacobi_scaling=False))


def test_lm_step_multi_tile_band(api, oracle_built, lo):
    """40 frames = 5 Cholesky tiles, band narrower than the matrix: exercises tile skipping,
    fill, trsm/update GEMMs and the multi-tile triangular solves."""
    sc = make_scene(40, 1500, 8, name="band")
    want = oracle_step(oracle_built, lo, sc, 1e3)
    compare(gpu_step(api, sc, 1e3), want)
    compare(gpu_step(api, sc, 1e3, dense_cholesky=1), want)


def test_lm_step_nested_dissection_vs_natural_order(api, oracle_built, lo):
    """160 frames = 20 Cholesky tiles: the nested-dissection ordering (elimination levels,
    conflict-free update groups, permuted right-hand side) must give the step of the natural
    order and of the numpy restatement."""
    sc = make_scene(160, 4000, 10, name="nd")
    want = oracle_step(oracle_built, lo, sc, 1e3)
    for kw in (dict(reorder_tiles=1), dict(reorder_tiles=0), dict(reorder_tiles=1, dense_cholesky=1)):
        compare(gpu_step(api, sc, 1e3, **kw), want)


def test_lm_step_duplicate_observations_in_one_frame(api, oracle_built, lo):
    """A point observed twice in the same frame (two residual blocks on the same parameter
    blocks): the frame's Schur panel is the sum of both."""
    sc = make_scene(12, 300, 6, name="dup")
    n = sc.num_obs
    extra = np.arange(0, n, 7)
    rng = np.random.default_rng(11)
    sc2 = Scene(**{**sc.__dict__,
                   "obs_xy": np.concatenate([sc.obs_xy, sc.obs_xy[extra] + rng.normal(0, 0.5, (extra.size, 2))]),
                   "obs_frame": np.concatenate([sc.obs_frame, sc.obs_frame[extra]]),
                   "obs_point": np.concatenate([sc.obs_point, sc.obs_point[extra]])})
    order = np.argsort(sc2.obs_frame, kind="stable")
    sc2 = Scene(**{**sc2.__dict__, "obs_xy": sc2.obs_xy[order], "obs_frame": sc2.obs_frame[order],
                   "obs_point": sc2.obs_point[order]})
    compare(gpu_step(api, sc2, 1e3), oracle_step(oracle_built, lo, sc2, 1e3))


def test_lm_step_constant_blocks_and_subsets(api, oracle_built, lo):
    """SetParameterBlockConstant on points / pose blocks and SubsetParameterization-style
    constant components (CeresHandler.h:288-300, 342-381)."""
    sc = make_scene(20, 600, 8, name="masks")
    mask = np.zeros(sc.num_frames, dtype=np.uint16)
    mask[0] = 0xFFF                   # first frame fixed (fixFirstNCameras)
    mask[1] = 0x03F                   # only pose0 of frame 1 fixed
    mask[sc.num_frames - 1] = 0b111000 << 6   # fixScale: last pose's position components
    mask[5] = 0b000111 | (0b000111 << 6)      # fixRotation on both poses of frame 5
    pconst = np.zeros(sc.num_points, dtype=np.uint8)
    pconst[::7] = 1                   # const3d-style frozen tracks
    want = oracle_step(oracle_built, lo, sc, 1e3, pose_mask=mask, point_const=pconst)
    got = gpu_step(api, sc, 1e3, pose_mask=mask, point_const=pconst)
    compare(got, want)
    assert not got["delta_points"][::7].any()
    assert not got["delta_poses"][0].any() and not got["delta_poses"][1, :6].any()


def test_lm_step_global_shutter_and_no_rotation_interp(api, oracle_built, lo):
    for kw in (dict(shutter=0), dict(interpolate_rotation=False)):
        sc = make_scene(12, 400, 8, name="modes", **kw)
        compare(gpu_step(api, sc, 1e3), oracle_step(oracle_built, lo, sc, 1e3))


def test_solve_c1_matches_oracle_loop(api, oracle_built, lo):
    sc = small_scene()
    ev = lambda po, pt, jac: oracle_built.evaluate(sc, po, pt, jac=jac, impl="port")  # noqa: E731
    po, pt, want = lo.solve(sc, ev, lo.Options(max_num_iterations=8))
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        s = pb.solve(api.default_options(max_num_iterations=8))
        gpo, gpt = pb.get_parameters()
    assert s.usable == 1
    assert s.iterations == want.iterations
    assert s.num_successful_steps == want.num_successful_steps
    assert abs(s.initial_cost - want.initial_cost) <= 1e-9 * want.initial_cost
    assert abs(s.final_cost - want.final_cost) <= TOL * want.final_cost
    assert relerr(gpo, po) <= 1e-5 and relerr(gpt, pt) <= 1e-5
    assert not (gpo[0] != sc.poses[0]).any()          # constant frame untouched


def test_solve_converges_and_reports(api, oracle_built):
    sc = make_scene(30, 2000, 8, name="solve30")
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        s = pb.solve(api.default_options(max_num_iterations=25))
        po, pt = pb.get_parameters()
    assert s.usable == 1 and s.final_cost < 0.05 * s.initial_cost
    r, _, v = oracle_built.evaluate(sc, po, pt, jac=False, impl="port")
    assert v.all()
    assert abs(0.5 * np.sum(r * r) - s.final_cost) <= 1e-9 * s.final_cost
    assert s.num_jacobian_evaluations == s.num_successful_steps + 1
    assert s.num_residual_evaluations == s.iterations or s.termination == 0


def test_solve_pointer_api_writes_back_in_place(api, oracle_built):
    sc = make_scene(8, 200, 6, name="ptr")
    poses, points = sc.poses.copy(), sc.points.copy()
    with api.Problem(0) as pb:
        pb.set_camera(sc.cam, sc.shutter, sc.scanlines, sc.interpolate_rotation)
        for i in range(sc.num_obs):
            f, p = int(sc.obs_frame[i]), int(sc.obs_point[i])
            pb.add_rs_residual(sc.obs_xy[i], poses[f, :6], poses[f, 6:], points[p])
        pb.set_block_constant(poses[0, :6])
        pb.set_block_constant(poses[0, 6:])
        s = pb.solve(api.default_options(max_num_iterations=10))
    with api.Problem(0) as pb2:
        pb2.load_scene(sc)
        s2 = pb2.solve(api.default_options(max_num_iterations=10))
        po2, pt2 = pb2.get_parameters()
    # the pointer API numbers points by first appearance, so reductions run in another order
    assert abs(s.final_cost - s2.final_cost) <= 1e-9 * s2.final_cost
    assert relerr(poses, po2) <= 1e-7 and relerr(points, pt2) <= 1e-7
    assert not poses[0].any()


def test_solve_fails_loudly_when_a_point_is_behind_the_camera(api):
    sc = small_scene()
    pts = sc.points.copy()
    pts[int(sc.obs_point[0])] = [0.0, 0.0, -5.0]
    with api.Problem(0) as pb:
        pb.load_scene(sc, points=pts)
        s = pb.solve(api.default_options(max_num_iterations=3), check=False)
    assert s.rc == api.ERR_EVALUATION_FAILED and s.usable == 0 and s.termination == 2


def test_lm_step_and_solve_with_huber_loss(api, oracle_built, lo):
    """HuberLoss through Ceres' Corrector: the LM step is the step of the rescaled residuals and
    Jacobians; the accepted cost is 1/2 sum rho."""
    sc = small_scene()
    xy = sc.obs_xy.copy()
    xy[::9] += 25.0
    sc = Scene(**{**sc.__dict__, "obs_xy": xy})
    a = 2.0
    r0, J0, v0 = oracle_built.evaluate(sc, impl="port")
    rw, Jw, _ = oracle_built.apply_huber(r0, J0, a)
    want = lo.lm_step(sc, rw, Jw, 1e3)
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        pb.set_loss(a)
        compare(pb.linearize_and_step(1e3), want)
        s = pb.solve(api.default_options(max_num_iterations=15))
        po, pt = pb.get_parameters()
    r1, _, v1 = oracle_built.evaluate(sc, po, pt, jac=False, impl="port")
    _, _, cost1 = oracle_built.apply_huber(r1, None, a)
    assert s.usable == 1 and s.final_cost < s.initial_cost
    assert abs(s.final_cost - cost1) <= 1e-9 * cost1
    with api.Problem(0) as pb:                          # options.huber_loss is the same switch
        pb.load_scene(sc)
        s2 = pb.solve(api.default_options(max_num_iterations=15, huber_loss=a))
    assert s2.final_cost == s.final_cost
    # robustified solve ends closer to the truth than the plain one on the contaminated scene
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        pb.solve(api.default_options(max_num_iterations=15))
        _, pt_plain = pb.get_parameters()
    e_rob = np.linalg.norm(pt - sc.points_true)
    e_plain = np.linalg.norm(pt_plain - sc.points_true)
    assert e_rob < e_plain


def gauge_free_problem(api, sc):
    pb = api.Problem(0)
    pb.set_camera(sc.cam, sc.shutter, sc.scanlines, sc.interpolate_rotation)
    pb.set_scene(sc.obs_xy, sc.obs_frame, sc.obs_point, sc.num_frames, sc.num_points,
                 const_pose_mask=np.zeros(sc.num_frames, dtype=np.uint16))
    pb.set_parameters(sc.poses, sc.points)
    return pb


def test_failed_linear_solve_is_an_invalid_step_not_the_end(api, oracle_built, lo):
    """Ceres 1.9 (trust_region_minimizer.cc; reached through CeresHandler.h:419): a reduced camera matrix that
    is not positive definite is an INVALID step -- radius /= decrease_factor, try again -- and only
    max_num_consecutive_invalid_steps of them in a row end the solve.  Scene: no constant frame (the 7-dof
    gauge freedom makes S singular) and a trust region so large that the damping is below rounding."""
    sc = small_scene()
    pm = np.zeros(sc.num_frames, dtype=np.int64)
    ev = lambda po, pt, jac: oracle_built.evaluate(sc, po, pt, jac=jac, impl="port")  # noqa: E731
    kw = dict(max_num_iterations=30, initial_trust_region_radius=1e20, max_trust_region_radius=1e32)
    _, _, want = lo.solve(sc, ev, lo.Options(max_num_consecutive_invalid_steps=40, **kw), pose_mask=pm)
    assert want.usable and any(t.get("reason") == "linear solver" for t in want.trace)
    with gauge_free_problem(api, sc) as pb:
        s = pb.solve(api.default_options(max_num_consecutive_invalid_steps=40, **kw))
        assert s.usable == 1 and s.termination in (0, 1)
        assert s.num_unsuccessful_steps >= 1
        assert s.iterations == s.num_successful_steps + s.num_unsuccessful_steps
        # same rule, same neighbourhood: where the factorisation first succeeds depends on rounding, so the two
        # trajectories are not identical -- both must have done most of the descent
        assert s.final_cost <= 1.5 * want.final_cost and want.final_cost <= 1.5 * s.final_cost
        assert s.final_cost < 0.05 * s.initial_cost


def test_successive_invalid_steps_fail_and_leave_the_last_accepted_iterate(api):
    sc = small_scene()
    with gauge_free_problem(api, sc) as pb:
        opt = api.default_options(max_num_iterations=30, initial_trust_region_radius=1e20, max_trust_region_radius=1e32,
                                  max_num_consecutive_invalid_steps=2)
        s = pb.solve(opt, check=False)
        assert s.usable == 0 and s.termination == 2
        assert s.num_successful_steps == 0 and s.iterations == 2
        assert b"invalid steps" in s.message or b"positive definite" in s.message
        po, pt = pb.get_parameters()
        assert np.array_equal(po, sc.poses) and np.array_equal(pt, sc.points)
