"""GPU parity for the "good initial guess" priors: GoodPosePrior <6; 6, 6>
(CeresHandler.h:55-73) as CeresHandler::Add wires it between f.priorPoses[i] and f.poses[i]
(CeresHandler.h:188-204).  The reference never fixes the prior block, so it is a free parameter block;
the device eliminates it in closed form.  The numpy restatement keeps the prior blocks as extra
pseudo-frames and solves the full system -- the LM step does not depend on the elimination order."""
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


def make_priors(sc, rot, pos, frames, seed=5, constant=()):
    """One prior per control pose of the listed frames: prior value = pose + small offset."""
    rng = np.random.default_rng(seed)
    out = []
    for f in frames:
        for which in (0, 1):
            val = sc.poses[f, 6 * which:6 * which + 6] + rng.normal(0, 1.0, 6) * np.array([2e-3] * 3 + [3e-2] * 3)
            out.append((f, which, rot, pos, val, (f, which) in constant))
    return out


def load(pb, sc, priors):
    pb.load_scene(sc)
    pb.set_pose_priors([p[0] for p in priors], [p[1] for p in priors], [p[2] for p in priors], [p[3] for p in priors],
                       np.array([p[4] for p in priors]), [int(p[5]) for p in priors])


def oracle_step(oracle, lo, sc, priors, radius, pose_mask=None):
    r, J, v = oracle.evaluate(sc, impl="port")
    sc2, mask2, Jx, rx, cx, ok = oracle.pose_prior_rows(sc, priors)
    assert ok
    if pose_mask is not None:
        mask2[:sc.num_frames] = pose_mask
    full = lo.lm_step(sc2, r, J, radius, pose_mask=mask2, extra=(Jx, rx))
    F, n = sc.num_frames, len(priors)
    nc = 12 * F
    free = np.concatenate([12 * (F + i) + np.arange(6) for i in range(n) if not priors[i][5]]).astype(int) \
        if any(not p[5] for p in priors) else np.zeros(0, dtype=int)
    S, rhs = full["S"], full["rhs"]
    Scc, Scp, Spp = S[:nc, :nc], S[:nc][:, free], S[free][:, free]
    X = np.linalg.solve(Spp, Scp.T) if free.size else np.zeros((0, nc))
    want = dict(S=Scc - Scp @ X, rhs=rhs[:nc] - X.T @ rhs[free], delta_poses=full["delta_poses"][:F],
                delta_points=full["delta_points"], model_cost_change=full["model_cost_change"],
                delta_priors=full["delta_poses"][F:, :6], prior_cost=cx)
    return want


@pytest.mark.parametrize("rot,pos", [(5.0, 2.0), (40.0, 0.0), (0.0, 15.0)])
def test_pose_prior_cost(api, oracle_built, rot, pos):
    sc = make_scene(12, 300, 8, name="pp-cost")
    priors = make_priors(sc, rot, pos, range(1, 12))
    r0, _, _ = oracle_built.evaluate(sc, impl="port", jac=False)
    _, _, _, _, cx, ok = oracle_built.pose_prior_rows(sc, priors)
    with api.Problem(0) as pb:
        load(pb, sc, priors)
        cost, _ = pb.evaluate_device(with_jacobian=False)
        cost2, _, _, _ = pb.evaluate()
    want = 0.5 * np.sum(r0 * r0) + cx
    assert ok and abs(cost - want) <= 1e-12 * want and cost2 == cost


@pytest.mark.parametrize("frames,rot,pos", [(12, 5.0, 2.0), (40, 30.0, 8.0)])
def test_lm_step_with_free_prior_blocks(api, oracle_built, lo, frames, rot, pos):
    sc = make_scene(frames, 40 * frames, 8, name=f"pp{frames}")
    priors = make_priors(sc, rot, pos, range(1, frames, 2))
    want = oracle_step(oracle_built, lo, sc, priors, 1e3)
    # (a FREE prior block absorbs its residual: the poses barely feel it -- the reference's behaviour --
    #  but the prior blocks themselves take a large step)
    assert np.abs(want["delta_priors"]).max() > 1e-3
    for kw in (dict(), dict(dense_cholesky=1), dict(jacobi_scaling=0)):
        w = want if "jacobi_scaling" not in kw else None
        if w is None:
            r, J, v = oracle_built.evaluate(sc, impl="port")
            sc2, mask2, Jx, rx, _, _ = oracle_built.pose_prior_rows(sc, priors)
            full = lo.lm_step(sc2, r, J, 1e3, lo.Options(jacobi_scaling=False), pose_mask=mask2, extra=(Jx, rx))
            w = dict(delta_poses=full["delta_poses"][:frames], delta_points=full["delta_points"],
                     model_cost_change=full["model_cost_change"], delta_priors=full["delta_poses"][frames:, :6])
        with api.Problem(0) as pb:
            load(pb, sc, priors)
            got = pb.linearize_and_step(1e3, api.default_options(**kw))
            val, trial = pb.pose_priors()
        for k in ("S", "rhs", "delta_poses", "delta_points"):
            if k in w:
                assert relerr(got[k], w[k]) <= TOL, (kw, k)
        assert abs(got["model_cost_change"] - w["model_cost_change"]) <= TOL * abs(w["model_cost_change"])
        assert relerr(trial - val, w["delta_priors"]) <= TOL


def test_lm_step_constant_prior_blocks_and_masked_poses(api, oracle_built, lo):
    """A prior block the caller fixed is a plain quadratic penalty; a prior on a (partly) constant pose keeps
    only its own column."""
    sc = make_scene(16, 500, 8, name="pp-const")
    priors = make_priors(sc, 12.0, 4.0, [0, 2, 3, 7, 9], constant={(2, 0), (2, 1), (7, 1)})
    mask = np.zeros(sc.num_frames, dtype=np.uint16)
    mask[0] = 0xFFF
    mask[9] = 0b000111 | (0b111000 << 6)
    want = oracle_step(oracle_built, lo, sc, priors, 2e2, pose_mask=mask)
    plain = lo.lm_step(sc, *oracle_built.evaluate(sc, impl="port")[:2], 2e2, pose_mask=mask)
    assert relerr(plain["delta_poses"], want["delta_poses"]) > 1e-4          # the constant priors pull the poses
    with api.Problem(0) as pb:
        pb.set_camera(sc.cam, sc.shutter, sc.scanlines, sc.interpolate_rotation)
        pb.set_scene(sc.obs_xy, sc.obs_frame, sc.obs_point, sc.num_frames, sc.num_points, mask)
        pb.set_parameters(sc.poses, sc.points)
        pb.set_pose_priors([p[0] for p in priors], [p[1] for p in priors], [p[2] for p in priors],
                           [p[3] for p in priors], np.array([p[4] for p in priors]), [int(p[5]) for p in priors])
        got = pb.linearize_and_step(2e2)
        val, trial = pb.pose_priors()
    for k in ("S", "rhs", "delta_poses", "delta_points"):
        assert relerr(got[k], want[k]) <= TOL, k
    assert relerr(trial - val, want["delta_priors"]) <= TOL
    cst = np.array([p[5] for p in priors])
    assert not (trial - val)[cst].any() and (trial - val)[~cst].any()


def test_solve_with_pose_priors_bulk_and_pointer_api(api, oracle_built):
    sc = make_scene(14, 400, 8, name="pp-solve")
    priors = make_priors(sc, 20.0, 6.0, range(1, 14))
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        s_plain = pb.solve(api.default_options(max_num_iterations=10))
    with api.Problem(0) as pb:
        load(pb, sc, priors)
        s = pb.solve(api.default_options(max_num_iterations=10))
        po, pt = pb.get_parameters()
        val, _ = pb.pose_priors()
    assert s.usable == 1 and s.final_cost < s.initial_cost
    assert s.num_parameters_reduced == s_plain.num_parameters_reduced + 6 * len(priors)
    moved = [(f, w, r_, p_, val[i], c) for i, (f, w, r_, p_, _, c) in enumerate(priors)]
    r1, _, _ = oracle_built.evaluate(sc, po, pt, jac=False, impl="port")
    _, _, _, _, cx, ok = oracle_built.pose_prior_rows(sc, moved, poses=po)
    want = 0.5 * np.sum(r1 * r1) + cx
    assert ok and abs(s.final_cost - want) <= 1e-9 * want
    assert np.abs(val - np.array([p[4] for p in priors])).max() > 1e-6        # the free prior blocks moved
    # pointer API, in the order of CeresHandler::Add: priors of the frame, then its observations
    poses, points = sc.poses.copy(), sc.points.copy()
    blocks = np.array([p[4] for p in priors]).copy()
    with api.Problem(0) as pb:
        pb.set_camera(sc.cam, sc.shutter, sc.scanlines, sc.interpolate_rotation)
        k = 0
        for f in range(sc.num_frames):
            if f >= 1:
                for which in (0, 1):
                    pb.add_pose_prior(20.0, 6.0, blocks[k], poses[f, 6 * which:6 * which + 6])
                    k += 1
            for i in np.flatnonzero(sc.obs_frame == f):
                pb.add_rs_residual(sc.obs_xy[i], poses[f, :6], poses[f, 6:], points[int(sc.obs_point[i])])
        pb.set_block_constant(poses[0, :6])
        pb.set_block_constant(poses[0, 6:])
        s2 = pb.solve(api.default_options(max_num_iterations=10))
    assert abs(s2.final_cost - s.final_cost) <= 1e-9 * s.final_cost
    assert relerr(poses, po) <= 1e-7 and relerr(blocks, val) <= 1e-7


def test_pose_prior_rotation_limit_fails_the_evaluation(api):
    """The functor returns false when rotation * (prior - pose)[0] >= 1 (CeresHandler.h:66): fatal at the
    initial point, like a point behind a camera."""
    sc = make_scene(8, 200, 6, name="pp-limit")
    val = sc.poses[3, :6].copy()
    val[0] += 0.2
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        pb.set_pose_priors([3], [0], [10.0], [1.0], val[None, :])
        with pytest.raises(api.RsbaError) as e:
            pb.solve(api.default_options(max_num_iterations=3))
        assert e.value.code == api.ERR_EVALUATION_FAILED
