"""Camera-only motion priors: the closed form used on the device
(oracle.motion_prior_coefficients) against the reference's own functors compiled verbatim
(RsConstVeloPrior / RsConstAccelerationPrior, video_bundler_rs_inter.h:55-173) under Jet autodiff."""
import numpy as np
import pytest


@pytest.mark.parametrize("kind", [1, 2])
@pytest.mark.parametrize("ratio", [1.0, 0.7, 2.5, 1e-3])
def test_closed_form_matches_reference_functor(oracle_built, kind, ratio):
    if not oracle_built.ref_available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(int(ratio * 1000) + kind)
    for _ in range(5):
        fk, fp = rng.normal(0, 0.3, 12), rng.normal(0, 0.3, 12)
        scale = float(rng.uniform(0.1, 30.0))
        ok, r_ref, J_ref, _ = oracle_built.motion_prior_eval_ref(kind, scale, ratio, fk, fp)
        r, J = oracle_built.motion_prior_eval(kind, scale, ratio, fk, fp)
        assert ok
        assert np.abs(r - r_ref).max() <= 1e-12 * max(1.0, np.abs(r_ref).max())
        assert np.abs(J - J_ref).max() <= 1e-12 * np.abs(J_ref).max()


def test_velocity_prior_at_zero_ratio_uses_the_previous_velocity(oracle_built):
    """interFrameRatio <= eps switches the second half to (end1 - pose1) (video_bundler_rs_inter.h:78-82)."""
    if not oracle_built.ref_available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    fk, fp = rng.normal(0, 0.3, 12), rng.normal(0, 0.3, 12)
    ok, r_ref, J_ref, _ = oracle_built.motion_prior_eval_ref(1, 2.0, 0.0, fk, fp)
    r, J = oracle_built.motion_prior_eval(1, 2.0, 0.0, fk, fp)
    assert ok and np.allclose(r, r_ref, rtol=0, atol=1e-13) and np.allclose(J, J_ref, rtol=0, atol=1e-13)
    # the functors' validity: velocity needs ratio >= 0, acceleration ratio >= eps
    assert not oracle_built.motion_prior_eval_ref(1, 2.0, -0.1, fk, fp)[0]
    assert not oracle_built.motion_prior_eval_ref(2, 2.0, 0.0, fk, fp)[0]


@pytest.mark.parametrize("kind", [1, 2])
@pytest.mark.parametrize("ratio", [1.0, 0.7, 2.5, 1e-3])
def test_ratio_column_matches_reference_functor(oracle_built, kind, ratio):
    """Free interFrameRatio (the reference's default, CeresHandler.h:156-180): d residual / d ratio of the
    closed form against the reference functor's own Jet column for the <1> block."""
    if not oracle_built.ref_available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(int(ratio * 1000) + 7 * kind)
    for _ in range(5):
        fk, fp = rng.normal(0, 0.3, 12), rng.normal(0, 0.3, 12)
        scale = float(rng.uniform(0.1, 30.0))
        ok, _, _, col_ref = oracle_built.motion_prior_eval_ref(kind, scale, ratio, fk, fp)
        col = oracle_built.motion_prior_ratio_column(kind, scale, ratio, fk, fp)
        assert ok and np.abs(col - col_ref).max() <= 1e-12 * max(1.0, np.abs(col_ref).max())
    # velocity prior at ratio <= eps: the second half no longer depends on the ratio
    ok, _, _, col_ref = oracle_built.motion_prior_eval_ref(1, 2.0, 0.0, fk, fp)
    col = oracle_built.motion_prior_ratio_column(1, 2.0, 0.0, fk, fp)
    assert ok and np.allclose(col, col_ref, rtol=0, atol=1e-13) and not col[6:].any()


def test_closed_form_matches_committed_reference_vectors(oracle_built):
    """The same pin without oracle/_ref: tests/golden/priors/prior_functors.npz holds the reference functors'
    own outputs (tests/golden/make_golden.py, generated where /root/reference exists)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "priors", "prior_functors.npz"))
    assert g["kind"].size >= 30
    for i in range(g["kind"].size):
        kind, ratio, scale = int(g["kind"][i]), float(g["ratio"][i]), float(g["scale"][i])
        r, J = oracle_built.motion_prior_eval(kind, scale, ratio, g["fk"][i], g["fp"][i])
        col = oracle_built.motion_prior_ratio_column(kind, scale, ratio, g["fk"][i], g["fp"][i])
        assert bool(g["ok"][i])
        assert np.abs(r - g["r"][i]).max() <= 1e-12 * max(1.0, np.abs(g["r"][i]).max())
        assert np.abs(J - g["J"][i]).max() <= 1e-12 * np.abs(g["J"][i]).max()
        assert np.abs(col - g["col"][i]).max() <= 1e-12 * max(1.0, np.abs(g["col"][i]).max())
