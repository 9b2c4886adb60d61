"""GPU parity at BASELINE.json's FULL size (C3: 1 000 frames / 200 k points / 5 M observations), where the CPU
restatement is too slow to be the checker for every element: size-independent properties instead.  One LM step of the
device pipeline (K1 -> K2 normal equations + Schur complement -> K3 tile Cholesky -> K4 back-substitution) must solve
the damped normal equations of ITS OWN linearisation,

    (J'^T J' + D^2) y = -J'^T r,   J' = J s,  s = 1 / (1 + |column|),  D^2 = clamp(diag J'^T J') / radius,  delta = s y

(what ceres::Solve's LevenbergMarquardtStrategy + SchurEliminator compute, CeresHandler.h:403,419), checked on the host
from the returned (r, J, delta) alone with matrix-free products -- no Schur complement, no factorisation, no oracle.
The residual is taken globally and component by component (relative to the magnitude of the terms summed into the
component), so one pose component or one point off by 0.1 % fails it (tests/test_lm_oracle_cpu.py).  (This file sorts last:
generating the 5 M-observation scene takes ~20 s on a fresh box.)"""
import numpy as np
import pytest

from helpers import damped_normal_equation_residual

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import rsba_b200.api as api
    api.load_library()
    return api


@pytest.fixture(scope="module")
def c3():
    from rsba_b200.scene import make_config
    return make_config("C3")


def test_lm_step_at_c3_solves_its_own_damped_normal_equations(api, c3):
    sc = c3
    radius = 1e4
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        cost, r, J, valid = pb.evaluate()
        got = pb.linearize_and_step(radius, want_S=False)
    assert valid.all()
    assert abs(cost - 0.5 * np.sum(r * r)) <= 1e-10 * cost
    assert np.all(got["delta_poses"][np.asarray(sc.const_frames, dtype=bool)] == 0.0)
    rel, mcc, comp = damped_normal_equation_residual(sc, r, J, got["delta_poses"], got["delta_points"], radius)
    assert rel <= 1e-6 and comp <= 1e-6, (rel, comp)
    assert mcc > 0.0 and abs(got["model_cost_change"] - mcc) <= 1e-6 * mcc, (got["model_cost_change"], mcc)


def test_solve_at_c3_reports_the_cost_of_the_parameters_it_returns(api, c3, oracle_built):
    """The whole loop at full size: every step of the first iterations is a descent step, the reported final cost is
    the cost of the returned parameters as the CPU restatement of the functor evaluates it on all 5 M observations,
    and a second run is bit-identical (fixed summation order everywhere)."""
    sc = c3
    runs = []
    for _ in range(2):
        with api.Problem(0) as pb:
            pb.load_scene(sc)
            s = pb.solve(api.default_options(max_num_iterations=6, function_tolerance=0.0, parameter_tolerance=0.0,
                                             gradient_tolerance=0.0))
            runs.append((s, *pb.get_parameters()))
    s, poses, points = runs[0]
    assert s.usable == 1 and s.iterations == 6 and s.num_successful_steps >= 4
    assert s.num_jacobian_evaluations == s.num_successful_steps + 1 and s.num_residual_evaluations == 6
    assert s.final_cost < 0.5 * s.initial_cost
    r, _, valid = oracle_built.evaluate(sc, poses, points, jac=False, impl="port")
    assert valid.all()
    assert abs(0.5 * np.sum(r * r) - s.final_cost) <= 1e-9 * s.final_cost
    assert np.array_equal(poses[np.asarray(sc.const_frames, dtype=bool)], sc.poses[np.asarray(sc.const_frames, dtype=bool)])
    assert np.array_equal(poses, runs[1][1]) and np.array_equal(points, runs[1][2]) and s.final_cost == runs[1][0].final_cost
