"""Edge cases of the LM path through the C ABI: unobserved parameter blocks, tiny problems, a scene
that spans exactly one / several Cholesky tiles and sub-tiles, repeated solves on one handle, scene
replacement, radius-driven rejections -- each against the numpy restatement where a reference exists."""
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


def step_pair(api, oracle, lo, sc, radius, **kw):
    r, J, v = oracle.evaluate(sc, impl="port")
    want = lo.lm_step(sc, r, J, radius)
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        got = pb.linearize_and_step(radius, api.default_options(**kw))
    return got, want


@pytest.mark.parametrize("frames", [1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 17])
def test_frame_counts_around_tile_and_subtile_borders(api, oracle_built, lo, frames):
    """4 frames = one SYRK sub-tile, 8 = one Cholesky tile: padding rows, partial sub-tiles."""
    sc = make_scene(frames, 60 * frames + 40, min(4, frames), name=f"f{frames}")
    if frames == 1:
        sc.const_frames[:] = False                 # a single free frame: the gauge is only held by the LM damping
    got, want = step_pair(api, oracle_built, lo, sc, 1e2)
    for k in ("S", "rhs", "delta_poses", "delta_points"):
        assert relerr(got[k], want[k]) <= TOL, (frames, k)


def test_unobserved_frame_and_point_do_not_move(api, oracle_built, lo):
    sc = make_scene(10, 300, 5, name="holes")
    keep = (sc.obs_frame != 6) & (sc.obs_point != 17)
    sc = Scene(**{**sc.__dict__, "obs_xy": sc.obs_xy[keep], "obs_frame": sc.obs_frame[keep],
                  "obs_point": sc.obs_point[keep]})
    got, want = step_pair(api, oracle_built, lo, sc, 1e3)
    for k in ("S", "rhs", "delta_poses", "delta_points"):
        assert relerr(got[k], want[k]) <= TOL, k
    assert not got["delta_poses"][6].any() and not got["delta_points"][17].any()
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        s = pb.solve(api.default_options(max_num_iterations=6))
        po, pt = pb.get_parameters()
    assert s.usable == 1 and s.final_cost < s.initial_cost
    assert np.array_equal(po[6], sc.poses[6]) and np.array_equal(pt[17], sc.points[17])


def test_repeated_solves_and_scene_replacement_on_one_handle(api, oracle_built):
    a = make_scene(12, 400, 6, name="first")
    b = make_scene(20, 900, 7, name="second")
    with api.Problem(0) as pb:
        pb.load_scene(a)
        s1 = pb.solve(api.default_options(max_num_iterations=5))
        pb.set_parameters(a.poses, a.points)                    # same scene again: identical trajectory
        s2 = pb.solve(api.default_options(max_num_iterations=5))
        assert s1.final_cost == s2.final_cost and s1.iterations == s2.iterations
        s3 = pb.solve(api.default_options(max_num_iterations=5))   # continue from the optimum
        assert s3.initial_cost == s2.final_cost and s3.final_cost <= s3.initial_cost
        pb.load_scene(b)                                         # a different structure on the same handle
        s4 = pb.solve(api.default_options(max_num_iterations=5))
        po, pt = pb.get_parameters()
    with api.Problem(0) as pb2:
        pb2.load_scene(b)
        s5 = pb2.solve(api.default_options(max_num_iterations=5))
        po2, pt2 = pb2.get_parameters()
    assert s4.final_cost == s5.final_cost and np.array_equal(po, po2) and np.array_equal(pt, pt2)


def test_bit_reproducible_runs(api):
    sc = make_scene(40, 1500, 8, name="repro")
    out = []
    for _ in range(2):
        with api.Problem(0) as pb:
            pb.load_scene(sc)
            s = pb.solve(api.default_options(max_num_iterations=6))
            out.append((s.final_cost,) + pb.get_parameters())
    assert out[0][0] == out[1][0]
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])


def test_small_trust_region_rejects_and_recovers(api, oracle_built, lo):
    """A far-off start makes the first steps fail the rho test: radius halves (decrease factor doubling)
    exactly as in the numpy loop."""
    sc = make_scene(10, 300, 6, name="far")
    rng = np.random.default_rng(2)
    pts = sc.points + rng.normal(0, 0.6, sc.points.shape)
    far = Scene(**{**sc.__dict__, "points": pts})
    ev = lambda po, pt, jac: oracle_built.evaluate(far, po, pt, jac=jac, impl="port")  # noqa: E731
    po, pt, want = lo.solve(far, ev, lo.Options(max_num_iterations=12))
    with api.Problem(0) as pb:
        pb.load_scene(far)
        s = pb.solve(api.default_options(max_num_iterations=12), check=False)
    assert s.iterations == want.iterations
    assert s.num_successful_steps == want.num_successful_steps
    assert s.num_unsuccessful_steps == want.num_unsuccessful_steps
    assert abs(s.final_cost - want.final_cost) <= 1e-6 * want.final_cost


def test_empty_problem_is_rejected_cleanly(api):
    sc = make_scene(4, 50, 3, name="empty")
    with api.Problem(0) as pb:
        pb.set_camera(sc.cam, sc.shutter, sc.scanlines, True)
        pb.set_scene(np.zeros((0, 2)), np.zeros(0, np.int32), np.zeros(0, np.int32), 4, 50)
        pb.set_parameters(sc.poses, sc.points)
        s = pb.solve(api.default_options(max_num_iterations=3), check=False)
        assert s.usable == 1 and s.final_cost == 0.0 and s.iterations == 0      # gradient tolerance at once


def test_sparse_pair_key_path_gives_the_same_structure(api, oracle_built, lo, monkeypatch):
    """Sequences beyond 16 384 frames index the sub-tile pairs through a sorted key list instead of a
    dense table; RSBA_CUDA_SPARSE_KEYS forces that path on a small scene."""
    sc = make_scene(40, 1500, 8, name="sparse-keys")
    got_dense, want = step_pair(api, oracle_built, lo, sc, 1e3)
    monkeypatch.setenv("RSBA_CUDA_SPARSE_KEYS", "1")
    got_sparse, _ = step_pair(api, oracle_built, lo, sc, 1e3)
    for k in ("S", "rhs", "delta_poses", "delta_points"):
        assert np.array_equal(got_sparse[k], got_dense[k]), k
        assert relerr(got_sparse[k], want[k]) <= TOL
