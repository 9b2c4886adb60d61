"""CPU suite, part 1: pins the oracle.

* the plain-C port (oracle/rsba_oracle.c) against the committed golden vectors, which were
  produced by the REFERENCE's own headers (tests/golden/make_golden.py -> oracle/_ref);
* where oracle/_ref is present (build container, or prebuilt on the GPU box) port vs ref live;
* known-answer restatements of the reference's own unit tests (src/rsba/test/mat_test.cc);
* the Jacobian against central finite differences (guards the Jet shim itself).
"""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from conftest import rel_block_err
from helpers import edge_scene, small_scene
from rsba_b200.scene import Scene

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def scene_from_golden(path) -> tuple[Scene, dict]:
    g = dict(np.load(path))
    sc = Scene(cam=g["cam"], shutter=int(g["shutter"]), scanlines=g["scanlines"],
               interpolate_rotation=bool(g["interpolate_rotation"]), poses=g["poses"], points=g["points"],
               obs_xy=g["obs_xy"], obs_frame=g["obs_frame"], obs_point=g["obs_point"],
               const_frames=np.zeros(g["poses"].shape[0], dtype=bool), name=os.path.basename(path))
    return sc, g


def test_golden_present():
    assert len(GOLDEN) >= 7


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_port_matches_golden(oracle_built, path):
    sc, g = scene_from_golden(path)
    res, J, valid = oracle_built.evaluate(sc, impl="port")
    assert np.array_equal(valid, g["valid"])
    ok = valid == 1
    assert np.max((np.abs(res - g["residuals"]) / np.maximum(1.0, np.abs(g["residuals"])))[ok]) <= 1e-9
    assert rel_block_err(J[ok], g["jacobian"][ok]).max() <= 1e-10
    assert not J[~ok].any() and not res[~ok].any()
    # cost-only instantiation agrees with the Jacobian instantiation
    res2, _, valid2 = oracle_built.evaluate(sc, impl="port", jac=False)
    assert np.array_equal(valid2, valid)
    assert np.max(np.abs(res2 - res) / np.maximum(1.0, np.abs(res))) <= 1e-9


def test_port_matches_ref_live(oracle_built):
    if not oracle_built.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for sc in [small_scene()] + [edge_scene(s, bool(r)) for s in (0, 1, 2) for r in (0, 1)]:
        r0, J0, v0 = oracle_built.evaluate(sc, impl="ref")
        r1, J1, v1 = oracle_built.evaluate(sc, impl="port")
        assert np.array_equal(v0, v1)
        ok = v0 == 1
        assert np.max((np.abs(r1 - r0) / np.maximum(1.0, np.abs(r0)))[ok]) <= 1e-9
        assert rel_block_err(J1[ok], J0[ok]).max() <= 1e-10


def test_jacobian_vs_finite_differences(oracle_built):
    sc = small_scene()
    n = 400
    sub = Scene(**{**sc.__dict__, "obs_xy": sc.obs_xy[:n], "obs_frame": sc.obs_frame[:n], "obs_point": sc.obs_point[:n]})
    _, J, valid = oracle_built.evaluate(sub, impl="port")
    assert valid.all()
    h = 1e-6
    Jfd = np.zeros_like(J)
    for k in range(12):
        for sgn in (+1, -1):
            p = sub.poses.copy()
            p[:, k] += sgn * h
            r, _, _ = oracle_built.evaluate(sub, poses=p, jac=False, impl="port")
            blk, col = divmod(k, 6)
            for row in range(2):
                Jfd[:, blk * 12 + row * 6 + col] += sgn * r[:, row] / (2 * h)
    for k in range(3):
        for sgn in (+1, -1):
            X = sub.points.copy()
            X[:, k] += sgn * h
            r, _, _ = oracle_built.evaluate(sub, points=X, jac=False, impl="port")
            for row in range(2):
                Jfd[:, 24 + row * 3 + k] += sgn * r[:, row] / (2 * h)
    assert rel_block_err(J, Jfd).max() <= 1e-6


def test_f3_structure(oracle_built):
    """J_pose1 / J_pose0 == tau / (1 - tau) because tau is parameter-independent."""
    sc = small_scene()
    _, J, _ = oracle_built.evaluate(sc, impl="port")
    tau = np.clip(sc.obs_xy[:, 0] / 1280.0, 0, 1)
    J0, J1 = J[:, :12], J[:, 12:24]
    assert np.allclose(J1 * (1 - tau)[:, None], J0 * tau[:, None], rtol=1e-9, atol=1e-9)


# ---------------------------------------------------------------------------------------
# Restatements of src/rsba/test/mat_test.cc on the port's primitives
# ---------------------------------------------------------------------------------------
def _prim(oracle):
    lib = oracle.port_lib()
    dp = C.POINTER(C.c_double)

    def call(fn, *arrays, extra=()):
        args = [a.ctypes.data_as(dp) if isinstance(a, np.ndarray) else a for a in arrays]
        return getattr(lib, fn)(*args, *extra)
    return lib, call


def _rodrigues_matrix(aa):
    th = np.linalg.norm(aa)
    if th < 1e-12:
        return np.eye(3)
    k = aa / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def test_rotation_known_answers(oracle_built):
    """TEST(SfM, Rotation), mat_test.cc:34-72: +-pi and +-pi/2 about each axis, to 1e-9,
    including the in-place call (pt == result)."""
    lib, _ = _prim(oracle_built)
    dp = C.POINTER(C.c_double)
    pts = np.array([[1.0, 2.0, 3.0], [-0.5, 0.25, 4.0]])
    for axis in range(3):
        for ang in (np.pi, -np.pi, np.pi / 2, -np.pi / 2, 0.3):
            aa = np.zeros(3)
            aa[axis] = ang
            for p in pts:
                out = np.zeros(3)
                lib.rsba_oracle_rotate(aa.ctypes.data_as(dp), p.ctypes.data_as(dp), out.ctypes.data_as(dp))
                assert np.allclose(out, _rodrigues_matrix(aa) @ p, atol=1e-9)
                q = p.copy()  # in place, as mat/cam.h:365 does
                lib.rsba_oracle_rotate(aa.ctypes.data_as(dp), q.ctypes.data_as(dp), q.ctypes.data_as(dp))
                assert np.allclose(q, out, atol=0)


def _quat(aa):
    th = np.linalg.norm(aa)
    if th < 1e-15:
        return np.array([1.0, 0, 0, 0])
    return np.concatenate([[np.cos(th / 2)], np.sin(th / 2) * aa / th])


def _quat_slerp(q0, q1, t):
    d = float(np.dot(q0, q1))
    if d < 0:
        q1, d = -q1, -d
    if d > 1 - 1e-12:
        q = q0 + t * (q1 - q0)
        return q / np.linalg.norm(q)
    om = np.arccos(d)
    return (np.sin((1 - t) * om) * q0 + np.sin(t * om) * q1) / np.sin(om)


def _quat_to_aa(q):
    q = q / np.linalg.norm(q)
    if q[0] < 0:
        q = -q
    s = np.linalg.norm(q[1:])
    if s < 1e-15:
        return np.zeros(3)
    return 2 * np.arctan2(s, q[0]) * q[1:] / s


def test_slerp_is_linear_and_matches_reference_test_properties(oracle_built):
    """TEST(SfM, SLERP), mat_test.cc:75-142: endpoints to 1e-6, extrapolation consistency
    (:116-122) and the loose <= 0.2 agreement with true quaternion slerp (:124-139)."""
    lib, _ = _prim(oracle_built)
    dp = C.POINTER(C.c_double)
    lib.rsba_oracle_slerp.argtypes = [dp, dp, C.c_double, dp]
    rng = np.random.default_rng(0)
    rots = np.concatenate([np.zeros((1, 3)), rng.uniform(-0.1, 0.1, size=(15, 3))])
    for ri in rots:
        for rj in rots:
            out = np.zeros(3)
            lib.rsba_oracle_slerp(ri.ctypes.data_as(dp), rj.ctypes.data_as(dp), 0.0, out.ctypes.data_as(dp))
            assert np.allclose(out, ri, atol=1e-6)
            lib.rsba_oracle_slerp(ri.ctypes.data_as(dp), rj.ctypes.data_as(dp), 1.0, out.ctypes.data_as(dp))
            assert np.allclose(out, rj, atol=1e-6)
            ext, back = np.zeros(3), np.zeros(3)
            lib.rsba_oracle_slerp(ri.ctypes.data_as(dp), rj.ctypes.data_as(dp), 2.0, ext.ctypes.data_as(dp))
            lib.rsba_oracle_slerp(ri.ctypes.data_as(dp), ext.ctypes.data_as(dp), 0.5, back.ctypes.data_as(dp))
            assert np.allclose(back, rj, atol=1e-6)
            for t in (0.25, 0.5, 0.75):
                lib.rsba_oracle_slerp(ri.ctypes.data_as(dp), rj.ctypes.data_as(dp), t, out.ctypes.data_as(dp))
                assert np.allclose(out, ri + (rj - ri) * t, atol=1e-15)          # F1: it IS a lerp
                true = _quat_to_aa(_quat_slerp(_quat(ri), _quat(rj), t))
                assert np.max(np.abs(out - true)) <= 0.2                          # mat_test.cc:130-136


def test_projection_roundtrip_and_validity(oracle_built):
    """TEST(SfM, reprojection) grid (mat_test.cc:170-229): w2c followed by the inverse rigid
    motion returns the point to 1e-6; w2i with validate rejects z < 1e-8 (mat/cam.h:410-412)."""
    lib, _ = _prim(oracle_built)
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(1)
    cams = [np.array([860.0, 860.0, k1, 0, 0, 0, 0, 640.0, 360.0]) for k1 in (0.0, 1e-3, -1e-3)]
    for _ in range(11):
        pose = np.concatenate([rng.uniform(-1, 1, 3), rng.uniform(-2, 2, 3)])
        R = _rodrigues_matrix(pose[:3])
        for _ in range(18):
            X = rng.uniform(-5, 5, 3)
            pc = np.zeros(3)
            lib.rsba_oracle_w2c(pose.ctypes.data_as(dp), X.ctypes.data_as(dp), pc.ctypes.data_as(dp))
            assert np.allclose(R.T @ pc + pose[3:], X, atol=1e-6)
            for cam in cams:
                proj = np.zeros(2)
                ok = lib.rsba_oracle_w2i(cam.ctypes.data_as(dp), pose.ctypes.data_as(dp), X.ctypes.data_as(dp),
                                         proj.ctypes.data_as(dp), 1)
                assert bool(ok) == (pc[2] >= 1e-8)


def test_distortion_known_answer(oracle_built):
    """distort (mat/cam.h:49-72) against the closed form written independently in numpy."""
    lib, _ = _prim(oracle_built)
    dp = C.POINTER(C.c_double)
    cam = np.array([800.0, 810.0, 0.1, -0.02, 1e-3, -2e-3, 5e-3, 0, 0])
    for u in ([0.1, -0.2], [0.0, 0.0], [-0.4, 0.35]):
        u = np.array(u)
        out = np.zeros(2)
        lib.rsba_oracle_distort(cam.ctypes.data_as(dp), u.ctypes.data_as(dp), out.ctypes.data_as(dp))
        r2 = u @ u
        d = 1 + cam[2] * r2 + cam[3] * r2 ** 2 + cam[6] * r2 ** 3
        ex = d * u[0] + 2 * cam[4] * u[0] * u[1] + cam[5] * (r2 + 2 * u[0] ** 2)
        ey = d * u[1] + cam[4] * (r2 + 2 * u[1] ** 2) + 2 * cam[5] * u[0] * u[1]
        assert np.allclose(out, [ex, ey], rtol=1e-13, atol=1e-15)


def test_port_on_the_reference_tests_own_grid(oracle_built):
    """The exact input tables of TEST(SfM, reprojection) (mat_test.cc:170-229: 11 poses x 18 points x 6 cameras) and
    TEST(SfM, Distortion) (:145-167): the port's w2c / w2i / distort against what the REFERENCE's own functions return
    on them (tests/golden/mat_test/grid.npz, written by make_golden.py from oracle/_ref), plus the properties the
    reference test asserts that involve only hot-path functions."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mat_test", "grid.npz"))
    lib, _ = _prim(oracle_built)
    dp = C.POINTER(C.c_double)
    p = lambda a: np.ascontiguousarray(a).ctypes.data_as(dp)  # noqa: E731
    pose, pt, cam = g["pose"], g["pt"], g["cam"]
    assert pose.shape == (11, 6) and pt.shape == (18, 3) and cam.shape == (6, 9)
    n_valid = 0
    for i in range(11):
        R = _rodrigues_matrix(pose[i, :3])
        for j in range(18):
            pc = np.zeros(3)
            lib.rsba_oracle_w2c(p(pose[i]), p(pt[j]), p(pc))
            assert np.allclose(pc, g["w2c"][i, j], rtol=1e-14, atol=1e-300)
            # CHECK_LE(dist3(pt[j], w3), 1e-6) with w3 = c2w(pose, w2c(pose, pt))  (mat_test.cc:207-212)
            assert np.linalg.norm(R.T @ pc + pose[i, 3:] - pt[j]) <= 1e-6 * max(1.0, np.linalg.norm(pt[j]))
            for k in range(6):
                proj = np.zeros(2)
                ok = lib.rsba_oracle_w2i(p(cam[k]), p(pose[i]), p(pt[j]), p(proj), 1)
                assert ok == g["ok"][i, j, k]
                assert bool(ok) == (pc[2] >= 1e-8)                    # mat/cam.h:410-412
                if ok:
                    n_valid += 1
                    assert np.allclose(proj, g["proj"][i, j, k], rtol=1e-13, atol=1e-300)
                if pc[2] != 0.0:                                      # validate = false (solveRSpnp.cpp:65)
                    lib.rsba_oracle_w2i(p(cam[k]), p(pose[i]), p(pt[j]), p(proj), 0)
                    assert np.allclose(proj, g["proj_nv"][i, j, k], rtol=1e-13, atol=1e-300)
    assert n_valid == int(g["ok"].sum()) == 672
    for k in range(4):
        for j in range(3):
            out = np.zeros(2)
            lib.rsba_oracle_distort(p(g["dcam"][k]), p(g["dimg"][j]), p(out))
            assert np.allclose(out, g["distort"][k, j], rtol=1e-14, atol=0)
            # CHECK_LE(dist2(img, undistort(distort(img))), 1e-6): the distortion is invertible there (fixed point)
            u = out.copy()
            k1, k2 = g["dcam"][k, 2], g["dcam"][k, 3]
            for _ in range(200):
                r2 = u @ u
                u = out / (1.0 + k1 * r2 + k2 * r2 * r2)
            assert np.linalg.norm(u - g["dimg"][j]) <= 1e-6
    if oracle_built.ref_available():                                  # the fixture IS what the reference computes here
        ref = oracle_built.ref_lib()
        for i in (0, 4, 9):
            for j in (0, 14, 16):
                pc = np.zeros(3)
                ref.rsba_ref_w2c(p(pose[i]), p(pt[j]), p(pc))
                assert np.array_equal(pc, g["w2c"][i, j])


def test_rotation_and_slerp_on_the_reference_tests_own_tables(oracle_built):
    """The exact vectors of TEST(SfM, Rotation) (mat_test.cc:34-72) and the 16-rotation table of TEST(SfM, SLERP)
    (:75-142) with the assertions the reference makes on them."""
    lib, _ = _prim(oracle_built)
    dp = C.POINTER(C.c_double)
    lib.rsba_oracle_slerp.argtypes = [dp, dp, C.c_double, dp]
    pi, pi2, eps = np.pi, np.pi / 2, np.finfo(np.float64).eps

    def rot(aa, pt, inplace=False):
        aa, pt = np.array(aa, dtype=np.float64), np.array(pt, dtype=np.float64)
        out = pt if inplace else np.zeros(3)
        lib.rsba_oracle_rotate(aa.ctypes.data_as(dp), pt.ctypes.data_as(dp), out.ctypes.data_as(dp))
        return out

    p, p2, p3, p4 = [0, 0, -10], [0, 0, 10], [10, 0, 0], [0, 10, 0]
    test = rot([0, pi, 0], p)
    assert np.linalg.norm(test - p2) <= 1e-9                                       # pi
    assert np.linalg.norm(rot([0, -pi, 0], test, inplace=True) - p) <= 1e-9        # invert3(r), in place (:51-53)
    assert np.linalg.norm(rot([0, -pi, 0], p) - p2) <= 1e-9                        # -pi
    assert np.linalg.norm(rot([0, -pi2, 0], p) - p3) <= 1e-9                       # -pi/2
    assert np.linalg.norm(rot([pi2, 0, 0], p) - p4) <= 1e-9                        # pi/2

    def slerp(a, b, t):
        a, b, out = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64), np.zeros(3)
        lib.rsba_oracle_slerp(a.ctypes.data_as(dp), b.ctypes.data_as(dp), float(t), out.ctypes.data_as(dp))
        return out

    r = np.array([[0, 0.5, 0], [0, 1, 0], [1, 0, 0], [0, 0, 1], [0, 1, 1], [1, 1, 1], [0, -1, 0], [-1, 0, 0],
                  [-1, 0, -1], [0, pi2, 0], [pi2, 0, 0], [0, 1 - pi2, 0], [0, -1, pi2], [0, -1, 1 - pi2], [0, 0, eps],
                  [1, -1, eps]], dtype=np.float64)
    for i in range(16):
        for j in range(1, 16):
            assert np.linalg.norm(slerp(r[i], r[j], 1.0) - r[j]) <= 1e-6
            assert np.linalg.norm(slerp(r[i], r[j], 0.0) - r[i]) <= 1e-6
            r05, r15, r20 = slerp(r[i], r[j], 0.5), slerp(r[i], r[j], 1.5), slerp(r[i], r[j], 2.0)
            assert np.linalg.norm(slerp(r[i], r20, 0.5) - r[j]) <= 1e-6
            assert np.linalg.norm(slerp(r05, r15, 0.5) - r[j]) <= 1e-6
            if i != j:   # the loose agreement with a true quaternion slerp (:124-139, CHECK_LE(..., 0.2))
                qi, qj = _quat(r[i]), _quat(r[j])
                frac = 2
                while frac < 500:
                    for n in range(1, frac):
                        tau = n / frac
                        true = _quat_to_aa(_quat_slerp(qi, qj, tau))
                        assert np.linalg.norm(slerp(r[i], r[j], tau) - true) <= 0.2, (i, j, tau)
                    frac = frac * 2 - 1
