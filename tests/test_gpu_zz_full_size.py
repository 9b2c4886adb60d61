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


def test_lm_step_at_c3_solves_its_own_damped_normal_equations(api):
    from rsba_b200.scene import make_config
    sc = make_config("C3")
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
