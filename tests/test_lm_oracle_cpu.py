"""CPU suite: independent checks of the Levenberg-Marquardt restatement the GPU solver is compared with
(oracle/lm_oracle.py, oracle/cpu_lm.cc).  Ceres 1.9.0 -- what the reference reaches through ceres::Solve
(CeresHandler.h:394-426) -- is not available here, so the restatement is pinned from two other sides:

* one LM step against a formulation that shares nothing with it: the damped step is the least-squares solution of the
  AUGMENTED system [J'; D] y = [-r; 0] by LAPACK's dense QR/SVD (no normal equations, no Schur complement);
* the whole loop against third-party solvers: MINPACK's Levenberg-Marquardt and scipy's trust-region reflective
  (scipy.optimize.least_squares) on the same residuals must reach the same minimum.
"""
import numpy as np
import scipy.optimize
import scipy.sparse as sp

import oracle
from oracle import lm_oracle
from helpers import damped_normal_equation_residual, small_scene
from rsba_b200.scene import make_scene


def _free_columns(scene):
    act_c, act_p = lm_oracle.param_masks(scene)
    return np.concatenate([act_c, act_p])


def test_lm_step_is_the_least_squares_solution_of_the_augmented_system():
    sc = small_scene()
    r, J, valid = oracle.evaluate(sc)
    assert valid.all()
    for radius in (1e4, 1e1, 1e-2):
        st = lm_oracle.lm_step(sc, r, J, radius, want_S=False)
        active = _free_columns(sc)
        Js = lm_oracle.sparse_jacobian(sc, J, active)
        Jp = (Js @ sp.diags(st["scale"])).toarray()[:, active]
        D = np.sqrt(st["D2"][active])
        A = np.vstack([Jp, np.diag(D)])
        b = np.concatenate([-r.reshape(-1), np.zeros(D.size)])
        y, *_ = np.linalg.lstsq(A, b, rcond=None)
        want = np.zeros(active.size)
        want[active] = y * st["scale"][active]
        got = np.concatenate([st["delta_poses"].reshape(-1), st["delta_points"].reshape(-1)])
        assert np.linalg.norm(got - want) <= 1e-8 * np.linalg.norm(want), radius
        # the model cost change is what the linearised residual predicts
        m = Jp @ y
        assert abs(st["model_cost_change"] - (-(m @ (r.reshape(-1) + 0.5 * m)))) <= 1e-9 * abs(st["model_cost_change"])


def test_cpu_lm_restatement_takes_the_same_step():
    # the C++/OpenMP restatement (bench.py's CPU leg) and the numpy one are written independently
    from oracle import cpu_lm
    sc = small_scene()
    r, J, _ = oracle.evaluate(sc)
    a = lm_oracle.lm_step(sc, r, J, 1e4, want_S=False)
    b = cpu_lm.CpuLm(sc).step(J, r, 1e4, compute_scale=True)
    for k in ("delta_poses", "delta_points"):
        assert np.linalg.norm(a[k] - b[k]) <= 1e-9 * np.linalg.norm(a[k]), k


def test_lm_loop_reaches_the_minimum_third_party_solvers_find():
    # small enough for dense third-party solvers: MINPACK's Levenberg-Marquardt (method "lm") and scipy's
    # trust-region reflective; two constant frames remove the gauge freedom so that the minimum is a point
    sc = make_scene(6, 60, 6, name="tiny")
    sc.const_frames[:2] = True
    F, P = sc.num_frames, sc.num_points
    active = _free_columns(sc)
    x0 = np.concatenate([sc.poses.reshape(-1), sc.points.reshape(-1)])

    def unpack(x_free):
        x = x0.copy()
        x[active] = x_free
        return x[:12 * F].reshape(F, 12), x[12 * F:].reshape(P, 3)

    def fun(x_free):
        poses, points = unpack(x_free)
        r, _, valid = oracle.evaluate(sc, poses, points, jac=False)
        assert valid.all()
        return r.reshape(-1)

    def jac(x_free):
        poses, points = unpack(x_free)
        _, J, _ = oracle.evaluate(sc, poses, points)
        return lm_oracle.sparse_jacobian(sc, J, active).toarray()[:, active]

    opts = lm_oracle.Options(max_num_iterations=300, function_tolerance=1e-15, parameter_tolerance=1e-14)
    poses, points, summ = lm_oracle.solve(sc, lambda po, pt, jac: oracle.evaluate(sc, po, pt, jac=jac), opts)
    assert summ.usable and "CONVERGENCE" in summ.termination and summ.num_successful_steps >= 3
    r, J, _ = oracle.evaluate(sc, poses, points)
    for method in ("lm", "trf"):
        ref = scipy.optimize.least_squares(fun, x0[active], jac=jac, method=method, x_scale="jac", ftol=1e-15,
                                           xtol=1e-15, gtol=1e-15, max_nfev=2000)
        assert ref.status > 0, ref.message
        # same minimum: cost (scipy's cost is 0.5 sum r^2 as well), residual vector, camera parameters
        assert abs(summ.final_cost - ref.cost) <= 1e-10 * ref.cost, (method, summ.final_cost, ref.cost)
        assert np.linalg.norm(r.reshape(-1) - ref.fun) <= 1e-6 * np.linalg.norm(ref.fun), method
        ref_poses, _ = unpack(ref.x)
        assert np.abs(poses - ref_poses).max() <= 1e-6, method
    # ... and stationary
    g = jac(np.concatenate([poses.reshape(-1), points.reshape(-1)])[active]).T @ r.reshape(-1)
    g0 = jac(x0[active]).T @ fun(x0[active])
    assert np.max(np.abs(g)) <= 1e-8 * np.max(np.abs(g0))
    # the reference's options (function tolerance 1e-6, CeresHandler.h:394-419) stop within 1e-5 of that minimum
    _, _, dflt = lm_oracle.solve(sc, lambda po, pt, jac: oracle.evaluate(sc, po, pt, jac=jac), lm_oracle.Options())
    assert "CONVERGENCE" in dflt.termination and 0.0 <= dflt.final_cost - summ.final_cost <= 1e-5 * summ.final_cost


def test_matrix_free_normal_equation_check_used_at_full_size():
    # the property the GPU test at C3 relies on (tests/test_gpu_zz_full_size.py), here at C2 with the C++ restatement
    # as the solver: it holds to rounding for a correct step and sees a single point's step off by 0.1 %
    from oracle import cpu_lm
    from rsba_b200.scene import make_config
    sc = make_config("C2")
    r, J, valid = oracle.evaluate(sc)
    assert valid.all()
    st = cpu_lm.CpuLm(sc).step(J, r, 1e4, compute_scale=True)
    rel, mcc, comp = damped_normal_equation_residual(sc, r, J, st["delta_poses"], st["delta_points"], 1e4)
    assert rel <= 1e-12 and comp <= 1e-12 and abs(mcc - st["model_cost_change"]) <= 1e-12 * mcc
    bad = st["delta_points"].copy()
    bad[12345] *= 1.001
    rel, _, comp = damped_normal_equation_residual(sc, r, J, st["delta_poses"], bad, 1e4)
    assert rel >= 2e-6 and comp >= 1e-4
    bad = st["delta_poses"].copy()
    bad[57, 4] *= 1.001
    assert damped_normal_equation_residual(sc, r, J, bad, st["delta_points"], 1e4)[2] >= 2e-6


def test_invalid_steps_shrink_the_radius_and_only_a_run_of_them_fails():
    """Ceres 1.9's rule for a failed linear solve / a step with model_cost_change <= 0 (trust_region_minimizer.cc,
    restated; constants in include/rsba_ceres_constants.h): the GPU loop follows the same rule
    (tests/test_gpu_lm.py::test_failed_linear_solve_is_an_invalid_step_not_the_end)."""
    from helpers import small_scene
    sc = small_scene()
    pm = np.zeros(sc.num_frames, dtype=np.int64)          # no constant frame: gauge freedom, S singular undamped
    ev = lambda po, pt, jac: oracle.evaluate(sc, po, pt, jac=jac, impl="port")  # noqa: E731
    kw = dict(max_num_iterations=30, initial_trust_region_radius=1e20, max_trust_region_radius=1e32)
    _, _, ok = lm_oracle.solve(sc, ev, lm_oracle.Options(max_num_consecutive_invalid_steps=40, **kw), pose_mask=pm)
    fails = [t for t in ok.trace if t.get("reason") == "linear solver"]
    assert ok.usable and len(fails) >= 2 and ok.final_cost < 0.05 * ok.initial_cost
    # the radius halves, quarters, ... exactly like after rejected steps
    radii = [t["radius"] for t in fails]
    assert np.allclose(radii[:2], [1e20 / 2, 1e20 / 8])
    po, pt, bad = lm_oracle.solve(sc, ev, lm_oracle.Options(max_num_consecutive_invalid_steps=2, **kw), pose_mask=pm)
    assert not bad.usable and "invalid" in bad.termination and bad.iterations == 2
    assert np.array_equal(po, sc.poses) and np.array_equal(pt, sc.points)
