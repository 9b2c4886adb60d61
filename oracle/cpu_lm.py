"""TEST INFRASTRUCTURE / CPU BASELINE -- Python driver of ``oracle/cpu_lm.cc`` (multi-threaded C++
restatement of Ceres 1.9.0's Schur-based LM linear solve; PARITY UNPINNED, see that file) with
the reduced camera system factorised by LAPACK ``dpbsv`` (``scipy.linalg.solveh_banded``) in
place of CHOLMOD.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np
import scipy.linalg

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libcpu_lm.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", _HERE, "libcpu_lm.so"], check=True, stdout=subprocess.DEVNULL)
        l = C.CDLL(path)
        vp = C.c_void_p
        l.cpu_lm_create.restype = vp
        l.cpu_lm_create.argtypes = [C.c_long, C.c_int, C.c_int, vp, vp, vp, vp]
        l.cpu_lm_destroy.argtypes = [vp]
        l.cpu_lm_band.argtypes = [vp]
        l.cpu_lm_build.argtypes = [vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, vp, vp, _dp, C.c_int]
        l.cpu_lm_backsub.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_int]
        _lib = l
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class CpuLm:
    """Structure built once per scene (as ceres::Solve does), one ``step`` per linear solve."""

    def __init__(self, scene, pose_mask=None, point_const=None, nthreads=0):
        self.l = lib()
        self.F, self.P, self.N = scene.num_frames, scene.num_points, scene.num_obs
        self.fr = np.ascontiguousarray(scene.obs_frame, dtype=np.int32)
        self.pt = np.ascontiguousarray(scene.obs_point, dtype=np.int32)
        if pose_mask is None:
            pose_mask = np.where(np.asarray(scene.const_frames, dtype=bool), 0xFFF, 0)
        self.pm = np.ascontiguousarray(pose_mask, dtype=np.uint16)
        self.pc = None if point_const is None else np.ascontiguousarray(point_const, dtype=np.uint8)
        self.h = self.l.cpu_lm_create(self.N, self.F, self.P, _p(self.fr), _p(self.pt), _p(self.pm), _p(self.pc))
        self.kd = self.l.cpu_lm_band(self.h)
        self.n = 12 * self.F
        self.ab = np.zeros((self.kd + 1, self.n))
        self.rhs = np.zeros(self.n)
        self.nthreads = nthreads
        self.times = {}

    def close(self):
        if self.h:
            self.l.cpu_lm_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def step(self, J, r, radius, min_diag=1e-6, max_diag=1e32, jacobi_scaling=True, compute_scale=False,
             want_S=False):
        """One linear solve.  Returns dict(delta_poses, delta_points, model_cost_change, step_norm,
        gmax[, S, rhs])."""
        J = np.ascontiguousarray(J, dtype=np.float64)
        r = np.ascontiguousarray(r, dtype=np.float64)
        gmax = C.c_double(0.0)
        t0 = time.perf_counter()
        self.l.cpu_lm_build(self.h, _p(J), _p(r), float(radius), float(min_diag), float(max_diag),
                            int(bool(jacobi_scaling)), int(bool(compute_scale)), _p(self.ab), _p(self.rhs),
                            C.byref(gmax), self.nthreads)
        t1 = time.perf_counter()
        out = {}
        if want_S:
            S = np.zeros((self.n, self.n))
            for d in range(self.kd + 1):
                idx = np.arange(self.n - d)
                S[idx + d, idx] = self.ab[d, :self.n - d]
                S[idx, idx + d] = self.ab[d, :self.n - d]
            out["S"] = S
            out["rhs"] = -self.rhs.copy()
        y = scipy.linalg.solveh_banded(self.ab, self.rhs, lower=True, overwrite_ab=True, check_finite=False)
        t2 = time.perf_counter()
        dc = np.zeros((self.F, 12))
        dpt = np.zeros((self.P, 3))
        sc = np.zeros(2)
        self.l.cpu_lm_backsub(self.h, _p(J), _p(r), _p(y), _p(dc), _p(dpt), _p(sc), self.nthreads)
        t3 = time.perf_counter()
        self.times = {"schur_ms": (t1 - t0) * 1e3, "cholesky_ms": (t2 - t1) * 1e3, "update_ms": (t3 - t2) * 1e3}
        out.update(delta_poses=dc, delta_points=dpt, model_cost_change=float(sc[0]), step_norm=float(sc[1]),
                   gmax=gmax.value)
        return out
