"""CPU suite: the host functor structs of include/rsba_cuda_functors.hpp -- the reference's
``ReprojectionError::operator()(pose, point, residuals)`` (video_bundler_free.h:33-41) and
``RsBundleAdjustment::operator()(pose0, pose1, point, residuals)`` (VideoSfmBaRs.h:25-35) signatures, T = double,
plus ``Evaluate`` in the shape of the AutoDiffCostFunction the reference wraps them in -- against the reference's own
functors: the committed golden vectors (generated from oracle/_ref = the reference headers compiled verbatim, incl.
the input grid of the reference's mat_test.cc) and, where oracle/_ref is present, live.
The header is compiled by plain g++ -std=c++11 -Wall -Werror: no CUDA in it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import rel_block_err
from helpers import edge_scene, small_scene
from test_oracle_cpu import GOLDEN, scene_from_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOLS = os.path.join(ROOT, "tests", "tools")


@pytest.fixture(scope="module")
def flib():
    so = os.path.join(TOOLS, "libfunctor_check.so")
    src = os.path.join(TOOLS, "functor_check.cc")
    hdrs = [os.path.join(ROOT, "include", h) for h in ("rsba_cuda_functors.hpp", "rsba_reproj_math.h", "rsba_ceres_constants.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in [src] + hdrs):
        subprocess.run(["g++", "-std=c++11", "-O2", "-Wall", "-Werror", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                        src, "-o", so], check=True)
    lib = C.CDLL(so)
    lib.functor_eval_rs.restype = C.c_long
    lib.functor_eval_single.restype = C.c_long
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def run_rs(lib, sc, with_cam=False):
    n = sc.num_obs
    res, J, v = np.zeros((n, 2)), np.zeros((n, 30)), np.zeros(n, np.uint8)
    Jc = np.zeros((n, 18)) if with_cam else None
    a = [np.ascontiguousarray(x, dtype=t) for x, t in (
        (sc.obs_xy, np.float64), (sc.obs_frame, np.int32), (sc.obs_point, np.int32), (sc.poses, np.float64),
        (sc.points, np.float64), (sc.cam, np.float64))]
    scan = np.ascontiguousarray(sc.scanlines, dtype=np.int32)
    bad = lib.functor_eval_rs(C.c_long(n), *[_p(x) for x in a], int(sc.shutter), _p(scan), int(bool(sc.interpolate_rotation)),
                              _p(res), _p(J), _p(v), _p(Jc) if with_cam else None)
    assert bad == 0, "operator() and Evaluate() of one functor disagree"
    return res, J, v, Jc


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_rs_functor_matches_reference_golden(flib, path):
    sc, g = scene_from_golden(path)
    res, J, valid, _ = run_rs(flib, sc)
    assert np.array_equal(valid, g["valid"])                     # the functor's bool, bit for bit
    ok = valid == 1
    assert rel_block_err(res[ok], g["residuals"][ok]).max() <= 1e-6 or np.abs(res - g["residuals"])[ok].max() < 1e-9
    assert rel_block_err(J[ok], g["jacobian"][ok]).max() <= 1e-9  # vs the reference functor under Jet<15>


def test_rs_functor_matches_reference_live(flib, oracle_built):
    if not oracle_built.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for sc in [small_scene()] + [edge_scene(s, bool(r)) for s in (0, 1, 2) for r in (0, 1)]:
        r0, J0, v0 = oracle_built.evaluate(sc, impl="ref")
        r1, J1, v1, _ = run_rs(flib, sc)
        assert np.array_equal(v0, v1)
        ok = v0 == 1
        assert np.max((np.abs(r1 - r0) / np.maximum(1.0, np.abs(r0)))[ok]) <= 1e-12
        assert rel_block_err(J1[ok], J0[ok]).max() <= 1e-10


def test_rs_functor_with_intrinsics_block_matches_reference(flib, oracle_built):
    """operator()(camera, pose0, pose1, point, residuals) (VideoSfmBaRs.h:38-49) and the <2; 9, 6, 6, 3> Jacobian."""
    if not oracle_built.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    sc = edge_scene(1, True)
    r0, J0, Jc0, v0 = oracle_built.evaluate_cam_ref(sc)
    r1, J1, v1, Jc1 = run_rs(flib, sc, with_cam=True)
    assert np.array_equal(v0, v1)
    ok = v0 == 1
    assert np.max((np.abs(r1 - r0) / np.maximum(1.0, np.abs(r0)))[ok]) <= 1e-12
    assert rel_block_err(J1[ok], J0[ok]).max() <= 1e-10
    assert rel_block_err(Jc1[ok], Jc0[ok]).max() <= 1e-10


def test_single_pose_functor_is_the_global_shutter_case(flib, oracle_built):
    """ReprojectionError::operator()(pose, point, residuals): the same projection with one pose -- equals the RS
    functor of a GLOBAL-shutter session, whose pose is pose0 (mat/cam.h:320-323)."""
    sc = small_scene()
    gs = type(sc)(**{**sc.__dict__, "shutter": 0})
    r0, J0, v0 = oracle_built.evaluate(gs, impl="port")
    n = sc.num_obs
    res, Jp, Jx, v = np.zeros((n, 2)), np.zeros((n, 12)), np.zeros((n, 6)), np.zeros(n, np.uint8)
    a = [np.ascontiguousarray(x, dtype=t) for x, t in (
        (sc.obs_xy, np.float64), (sc.obs_frame, np.int32), (sc.obs_point, np.int32), (sc.poses, np.float64),
        (sc.points, np.float64), (sc.cam, np.float64))]
    assert flib.functor_eval_single(C.c_long(n), *[_p(x) for x in a], _p(res), _p(Jp), _p(Jx), _p(v)) == 0
    assert np.array_equal(v, v0)
    ok = v0 == 1
    assert np.max((np.abs(res - r0) / np.maximum(1.0, np.abs(r0)))[ok]) <= 1e-12
    assert rel_block_err(Jp[ok], J0[ok][:, :12]).max() <= 1e-10
    assert rel_block_err(Jx[ok], J0[ok][:, 24:]).max() <= 1e-10
