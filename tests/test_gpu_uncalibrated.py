"""GPU parity of the uncalibrated variant: RsBundleAdjustment::CreateWithCam <2; 9, 6, 6, 3>
(VideoSfmBaRs.h:38-49,68-80; CeresHandler.h:256-264) -- the shared intrinsics as a parameter block.  The LM step
(reduced system with its dense border, step of frames / intrinsics / points, model cost change) against the
numpy restatement with the intrinsics ordered as a last pseudo-frame, and full solves."""
import numpy as np
import pytest

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


def oracle_step(oracle, lo, sc, radius):
    r, J, v = oracle.evaluate(sc, impl="port")
    Jc = oracle.intrinsics_jacobian(sc)
    sc2, mask, cols = lo.with_intrinsics_block(sc, J, Jc)
    return lo.lm_step(sc2, r, J, radius, pose_mask=mask, cam_cols=cols)


@pytest.mark.parametrize("frames,points,per", [(10, 500, 10), (13, 400, 6), (40, 1500, 8)])
def test_lm_step_uncalibrated(api, oracle_built, lo, frames, points, per):
    """frames = 13: the pseudo-frame shares its sub-tile and tile with real frames; 40: five tiles + border."""
    sc = make_scene(frames, points, per, name=f"uncal{frames}")
    want = oracle_step(oracle_built, lo, sc, 1e3)
    for kw in (dict(), dict(reorder_tiles=0), dict(dense_cholesky=1)):
        with api.Problem(0) as pb:
            pb.set_intrinsics_free(True)
            pb.load_scene(sc)
            got = pb.linearize_and_step(1e3, api.default_options(**kw))
        for k in ("S", "rhs", "delta_poses", "delta_points"):
            assert relerr(got[k], want[k]) <= TOL, (kw, k)
        assert abs(got["model_cost_change"] - want["model_cost_change"]) <= TOL * abs(want["model_cost_change"])
        assert np.abs(got["delta_poses"][frames, :9]).max() > 0 and not got["delta_poses"][frames, 9:].any()


def test_solve_uncalibrated_from_perturbed_intrinsics(api, oracle_built):
    sc = make_scene(30, 2000, 8, name="uncal-solve")
    cam0 = sc.cam.copy()
    cam0[0] *= 1.02
    cam0[1] *= 0.985
    cam0[7] += 6.0
    cam0[8] -= 4.0
    bad = Scene(**{**sc.__dict__, "cam": cam0})
    with api.Problem(0) as pb:
        pb.load_scene(bad)
        s_cal = pb.solve(api.default_options(max_num_iterations=25))
    with api.Problem(0) as pb:
        pb.set_intrinsics_free(True)
        pb.load_scene(bad)
        s = pb.solve(api.default_options(max_num_iterations=25))
        cam1 = pb.get_camera()
        po, pt = pb.get_parameters()
        cost_again, bad_obs = pb.evaluate_device(with_jacobian=False)
    assert s.usable == 1 and s.final_cost < s_cal.final_cost                # nine more degrees of freedom
    assert abs(cost_again - s.final_cost) <= 1e-12 * s.final_cost and bad_obs == 0
    assert s.num_parameters_reduced == s_cal.num_parameters_reduced + 9
    # the optimised intrinsics reproduce the final cost through the oracle
    final = Scene(**{**sc.__dict__, "cam": cam1})
    r, _, v = oracle_built.evaluate(final, po, pt, jac=False, impl="port")
    assert v.all() and abs(0.5 * np.sum(r * r) - s.final_cost) <= 1e-9 * s.final_cost
    assert np.abs(cam1 - cam0).max() > 1e-3                                  # the intrinsics did move


def test_solve_uncalibrated_matches_numpy_loop(api, oracle_built, lo):
    sc = make_scene(10, 500, 10, name="uncal-loop")
    # numpy LM loop over [frames | intrinsics pseudo-frame | points]
    def ev(po, pt, jac):
        cam = po[-1, :9]
        cur = Scene(**{**sc.__dict__, "cam": cam})
        r, J, v = oracle_built.evaluate(cur, po[:-1], pt, jac=jac, impl="port")
        ev.Jc = oracle_built.intrinsics_jacobian(cur, po[:-1], pt) if jac else None
        return r, J, v
    r, J, v = oracle_built.evaluate(sc, impl="port")
    sc2, mask, _ = lo.with_intrinsics_block(sc, J, oracle_built.intrinsics_jacobian(sc))
    # lm_oracle.solve has no hook for the dense columns: run the loop by hand with lm_step (same rules)
    poses, points = sc2.poses.copy(), sc.points.copy()
    opts = lo.Options(max_num_iterations=6)
    radius, decrease, scale = opts.initial_trust_region_radius, 2.0, None
    r, J, v = ev(poses, points, True)
    cost = 0.5 * np.sum(r * r)
    accepted_steps = 0
    for _ in range(opts.max_num_iterations):
        cols = np.zeros((2 * sc.num_obs, 12))
        cols[0::2, :9], cols[1::2, :9] = ev.Jc[:, :9], ev.Jc[:, 9:]
        cur = Scene(**{**sc2.__dict__, "poses": poses})
        st = lo.lm_step(cur, r, J, radius, opts, scale, mask, None, want_S=False, cam_cols=cols)
        scale = st["scale"]
        tp, tq = poses + st["delta_poses"], points + st["delta_points"]
        rt, _, vt = ev(tp, tq, False)
        new_cost = 0.5 * np.sum(rt * rt)
        rho = (cost - new_cost) / st["model_cost_change"]
        if vt.all() and rho > opts.min_relative_decrease:
            poses, points = tp, tq
            radius = min(opts.max_trust_region_radius, radius / max(1 / 3, 1 - (2 * rho - 1) ** 3))
            decrease = 2.0
            r, J, v = ev(poses, points, True)
            cost = 0.5 * np.sum(r * r)
            accepted_steps += 1
        else:
            radius /= decrease
            decrease *= 2
    with api.Problem(0) as pb:
        pb.set_intrinsics_free(True)
        pb.load_scene(sc)
        s = pb.solve(api.default_options(max_num_iterations=6, function_tolerance=0.0, parameter_tolerance=0.0,
                                         gradient_tolerance=0.0))
        cam1 = pb.get_camera()
        po, pt = pb.get_parameters()
    assert s.num_successful_steps == accepted_steps
    assert abs(s.final_cost - cost) <= 1e-6 * cost
    assert relerr(cam1, poses[-1, :9]) <= 1e-6 and relerr(po, poses[:-1]) <= 1e-5 and relerr(pt, points) <= 1e-5
