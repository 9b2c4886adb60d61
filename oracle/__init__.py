"""TEST INFRASTRUCTURE -- CPU checkers for the rolling-shutter BA hot path.

* ``port``  : ``liboracle.so``  -- plain-C restatement (``oracle/rsba_oracle.c``)
* ``ref``   : ``_ref/librsba_ref.so`` -- the reference's own headers compiled verbatim
  (``oracle/ref_driver.cc``; built only where ``/root/reference`` exists, travels prebuilt)
* ``lm_oracle`` : numpy restatement of one Ceres-1.9-style LM/Schur step

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package.  Nothing under ``rsba_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_bp = C.POINTER(C.c_ubyte)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


def build(quiet: bool = True) -> None:
    """Compile the C port (and oracle/_ref when the reference tree is present)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


_port = None
_ref = None


def port_lib():
    global _port
    if _port is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.rsba_oracle_eval.restype = C.c_long
        lib.rsba_oracle_eval.argtypes = [C.c_long, _dp, _ip, _ip, _dp, _dp, _dp, C.c_int, _ip, C.c_int,
                                         _dp, _dp, _bp, C.c_int]
        lib.rsba_oracle_w2i.restype = C.c_int
        _port = lib
    return _port


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "librsba_ref.so"))


def ref_lib():
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "librsba_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref/librsba_ref.so not built (needs /root/reference)")
        lib = C.CDLL(path)
        lib.rsba_ref_problem_create.restype = C.c_void_p
        lib.rsba_ref_problem_create.argtypes = [C.c_long, _dp, _dp, C.c_int, _ip, C.c_int]
        lib.rsba_ref_problem_destroy.argtypes = [C.c_void_p]
        lib.rsba_ref_problem_eval.restype = C.c_long
        lib.rsba_ref_problem_eval.argtypes = [C.c_void_p, C.c_long, _ip, _ip, _dp, _dp, _dp, _dp, _bp, C.c_int]
        lib.rsba_ref_w2i.restype = C.c_int
        lib.rsba_ref_norm3.restype = C.c_double
        lib.rsba_ref_slerp.argtypes = [_dp, _dp, C.c_double, _dp]
        lib.rsba_ref_velo_prior.argtypes = [C.c_double] + [_dp] * 7
        lib.rsba_ref_accel_prior.argtypes = [C.c_double] + [_dp] * 7
        _ref = lib
    return _ref


def _prep(scene, poses, points):
    poses = np.ascontiguousarray(scene.poses if poses is None else poses, dtype=np.float64)
    points = np.ascontiguousarray(scene.points if points is None else points, dtype=np.float64)
    return poses, points


def evaluate(scene, poses=None, points=None, jac=True, impl="port", nthreads=0):
    """Residuals [N,2], Jacobians [N,30] (or None), valid [N] for ``scene`` at the given
    parameters.  ``impl`` = "port" (C restatement) or "ref" (reference headers verbatim)."""
    poses, points = _prep(scene, poses, points)
    n = scene.num_obs
    res = np.zeros((n, 2))
    J = np.zeros((n, 30)) if jac else None
    valid = np.zeros(n, dtype=np.uint8)
    cam = np.ascontiguousarray(scene.cam, dtype=np.float64)
    scan = np.ascontiguousarray(scene.scanlines, dtype=np.int32)
    oxy = np.ascontiguousarray(scene.obs_xy, dtype=np.float64)
    fi = np.ascontiguousarray(scene.obs_frame, dtype=np.int32)
    pi = np.ascontiguousarray(scene.obs_point, dtype=np.int32)
    if impl == "port":
        port_lib().rsba_oracle_eval(n, _ptr(oxy, _dp), _ptr(fi, _ip), _ptr(pi, _ip), _ptr(poses, _dp),
                                    _ptr(points, _dp), _ptr(cam, _dp), int(scene.shutter), _ptr(scan, _ip),
                                    int(bool(scene.interpolate_rotation)), _ptr(res, _dp), _ptr(J, _dp),
                                    _ptr(valid, _bp), nthreads)
    elif impl == "ref":
        lib = ref_lib()
        h = lib.rsba_ref_problem_create(n, _ptr(oxy, _dp), _ptr(cam, _dp), int(scene.shutter),
                                        _ptr(scan, _ip), int(bool(scene.interpolate_rotation)))
        try:
            lib.rsba_ref_problem_eval(h, n, _ptr(fi, _ip), _ptr(pi, _ip), _ptr(poses, _dp), _ptr(points, _dp),
                                      _ptr(res, _dp), _ptr(J, _dp), _ptr(valid, _bp), nthreads)
        finally:
            lib.rsba_ref_problem_destroy(h)
    else:
        raise ValueError(impl)
    return res, J, valid


class RefProblem:
    """Reference cost functions built once (untimed), evaluated many times (timed) --
    the way ceres::Problem owns them.  Used by bench.py's CPU baseline."""

    def __init__(self, scene):
        self.scene = scene
        self.lib = ref_lib()
        self.cam = np.ascontiguousarray(scene.cam, dtype=np.float64)
        self.scan = np.ascontiguousarray(scene.scanlines, dtype=np.int32)
        self.oxy = np.ascontiguousarray(scene.obs_xy, dtype=np.float64)
        self.fi = np.ascontiguousarray(scene.obs_frame, dtype=np.int32)
        self.pi = np.ascontiguousarray(scene.obs_point, dtype=np.int32)
        self.h = self.lib.rsba_ref_problem_create(scene.num_obs, _ptr(self.oxy, _dp), _ptr(self.cam, _dp),
                                                  int(scene.shutter), _ptr(self.scan, _ip),
                                                  int(bool(scene.interpolate_rotation)))

    def eval(self, poses, points, res, J, valid, nthreads=0):
        return self.lib.rsba_ref_problem_eval(self.h, self.scene.num_obs, _ptr(self.fi, _ip), _ptr(self.pi, _ip),
                                              _ptr(poses, _dp), _ptr(points, _dp), _ptr(res, _dp), _ptr(J, _dp),
                                              _ptr(valid, _bp), nthreads)

    def close(self):
        if self.h:
            self.lib.rsba_ref_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()
