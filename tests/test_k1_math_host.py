"""CPU suite, part 3: the product's per-observation K1 arithmetic
(rsba_b200/csrc/reproj_math.cuh) compiled for the host by tests/tools and compared with the
oracle and the golden vectors.  A debugging aid for the GPU-less build container; the GPU
parity tests proper are in test_gpu_k1.py."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import rel_block_err
from helpers import edge_scene, small_scene
from test_oracle_cpu import GOLDEN, scene_from_golden

TOOLS = os.path.join(os.path.dirname(__file__), "tools")


@pytest.fixture(scope="module")
def hostlib():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    so = os.path.join(TOOLS, "libk1_host_check.so")
    src = os.path.join(TOOLS, "k1_host_check.cu")
    hdr = os.path.join(TOOLS, "..", "..", "rsba_b200", "csrc", "reproj_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-w", "-o", so, src], check=True)
    lib = C.CDLL(so)
    lib.k1_host_eval.restype = C.c_long
    return lib


def run_host(lib, sc):
    n = sc.num_obs
    res, J, v = np.zeros((n, 2)), np.zeros((n, 30)), np.zeros(n, np.uint8)
    arrs = [np.ascontiguousarray(a, dtype=t) for a, t in (
        (sc.obs_xy, np.float64), (sc.obs_frame, np.int32), (sc.obs_point, np.int32), (sc.poses, np.float64),
        (sc.points, np.float64), (sc.cam, np.float64))]
    scan = np.ascontiguousarray(sc.scanlines, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib.k1_host_eval(C.c_long(n), *[p(a) for a in arrs], int(sc.shutter), p(scan), int(bool(sc.interpolate_rotation)),
                     p(res), p(J), p(v))
    return res, J, v


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_k1_math_matches_golden(hostlib, path):
    sc, g = scene_from_golden(path)
    res, J, valid = run_host(hostlib, sc)
    assert np.array_equal(valid, g["valid"])
    ok = valid == 1
    assert rel_block_err(res[ok], g["residuals"][ok]).max() <= 1e-6 or np.abs(res - g["residuals"])[ok].max() < 1e-9
    assert rel_block_err(J[ok], g["jacobian"][ok]).max() <= 1e-9
    assert not J[~ok].any() and not res[~ok].any()


def test_k1_math_matches_oracle_c1(hostlib, oracle_built):
    sc = small_scene()
    r0, J0, v0 = oracle_built.evaluate(sc, impl="port")
    r1, J1, v1 = run_host(hostlib, sc)
    assert np.array_equal(v0, v1)
    assert np.abs(r1 - r0).max() <= 1e-9
    assert rel_block_err(J1, J0).max() <= 1e-9
