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


def apply_huber(res, J, a):
    """TEST INFRASTRUCTURE.  ceres::HuberLoss(a) as Ceres 1.9.0 applies it to a residual block
    (loss_function.cc HuberLoss::Evaluate + corrector.cc; third-party, restated): with
    s = |r|^2, rho(s) = s (s <= a^2) else 2 a sqrt(s) - a^2; rho'' <= 0, hence residual and
    Jacobian are rescaled by sqrt(rho') and the block's cost is 1/2 rho(s).
    Returns (corrected residuals, corrected Jacobian or None, total cost)."""
    s = np.sum(res * res, axis=1)
    out = s > a * a
    sr = np.sqrt(np.where(out, s, 1.0))
    w = np.where(out, np.sqrt(a / sr), 1.0)
    rho = np.where(out, 2.0 * a * sr - a * a, s)
    return res * w[:, None], None if J is None else J * w[:, None], 0.5 * float(np.sum(rho))


def validate_sweep(scene, poses=None, points=None, sqrd_threshold=16.0, min_distance=0.0):
    """TEST INFRASTRUCTURE.  validate(sess, f, opt, pt, obs) (struct/VideoSfM.cc:159-169) for every
    observation, restated on the port's primitives: getPose -> interpolate_rs with the REAL
    observation (x or y by shutter, struct/VideoSfM.cc:108-111), norm3(c - X) >= minDistanceToCamera,
    reprojection_error + threshold (mat/cam.h:425-457).  Python loop: small scenes only.
    Returns (ok [N] uint8, squared error [N], -1 where w2i failed)."""
    poses, points = _prep(scene, poses, points)
    poses, points = poses.reshape(-1, 12), points.reshape(-1, 3)
    lib = port_lib()
    lib.rsba_oracle_interpolate_rs.argtypes = [_dp, _dp, C.c_int, _ip, _dp, _dp, C.c_int]
    lib.rsba_oracle_w2i.argtypes = [_dp, _dp, _dp, _dp, C.c_int]
    cam = np.ascontiguousarray(scene.cam, dtype=np.float64)
    scan = np.ascontiguousarray(scene.scanlines, dtype=np.int32)
    n = scene.num_obs
    ok = np.zeros(n, dtype=np.uint8)
    err = np.full(n, -1.0)
    pose = np.zeros(6)
    proj = np.zeros(2)
    for i in range(n):
        f, p = int(scene.obs_frame[i]), int(scene.obs_point[i])
        obs = np.ascontiguousarray(scene.obs_xy[i], dtype=np.float64)
        p0 = np.ascontiguousarray(poses[f, :6])
        p1 = np.ascontiguousarray(poses[f, 6:])
        X = np.ascontiguousarray(points[p])
        lib.rsba_oracle_interpolate_rs(_ptr(p0, _dp), _ptr(p1, _dp), int(scene.shutter), _ptr(scan, _ip),
                                       _ptr(obs, _dp), _ptr(pose, _dp), int(bool(scene.interpolate_rotation)))
        good = lib.rsba_oracle_w2i(_ptr(cam, _dp), _ptr(pose, _dp), _ptr(X, _dp), _ptr(proj, _dp), 1)
        if good:
            err[i] = float(np.sum((proj - obs) ** 2))
            ok[i] = 1 if (np.linalg.norm(pose[3:] - X) >= min_distance and err[i] < sqrd_threshold) else 0
    return ok, err


def reproject_sweep(scene, frame, point, poses=None, points=None, sqrd_threshold=16.0):
    """TEST INFRASTRUCTURE.  reproject(sess, f, opt, pt, obs) (struct/VideoSfM.cc:139-155) for a list of
    (frame, point) pairs, restated on the port's primitives: start at the principal point, repeat
    getPose(proj) -> w2i until the projection moves by <= 1e-3 px (limit 50, pre-decremented), then
    ::vision::validate of the projection against itself.  Python loop: small batches only.
    Returns (proj [n, 2], ok [n] uint8)."""
    poses, points = _prep(scene, poses, points)
    poses, points = poses.reshape(-1, 12), points.reshape(-1, 3)
    lib = port_lib()
    lib.rsba_oracle_interpolate_rs.argtypes = [_dp, _dp, C.c_int, _ip, _dp, _dp, C.c_int]
    lib.rsba_oracle_w2i.argtypes = [_dp, _dp, _dp, _dp, C.c_int]
    cam = np.ascontiguousarray(scene.cam, dtype=np.float64)
    scan = np.ascontiguousarray(scene.scanlines, dtype=np.int32)
    n = len(frame)
    out, ok = np.zeros((n, 2)), np.zeros(n, dtype=np.uint8)
    pose = np.zeros(6)
    for i in range(n):
        p0 = np.ascontiguousarray(poses[int(frame[i]), :6])
        p1 = np.ascontiguousarray(poses[int(frame[i]), 6:])
        X = np.ascontiguousarray(points[int(point[i])])
        proj = np.array([cam[7], cam[8]])
        limit, good = 50, False
        while True:
            limit -= 1
            if limit < 1:
                break
            proj0 = proj.copy()
            lib.rsba_oracle_interpolate_rs(_ptr(p0, _dp), _ptr(p1, _dp), int(scene.shutter), _ptr(scan, _ip),
                                           _ptr(proj0, _dp), _ptr(pose, _dp), int(bool(scene.interpolate_rotation)))
            if not lib.rsba_oracle_w2i(_ptr(cam, _dp), _ptr(pose, _dp), _ptr(X, _dp), _ptr(proj, _dp), 1):
                break
            d = proj0 - proj
            if not (d @ d > 1e-6):
                good = True
                break
        out[i] = proj
        ok[i] = 1 if (good and 0.0 < sqrd_threshold) else 0
    return out, ok


# --------------------------------------------------------------------------------------------
# camera-only motion priors
PRIOR_VELOCITY, PRIOR_ACCELERATION = 1, 2
_EPS = np.finfo(np.float64).eps


def motion_prior_coefficients(kind, ratio):
    """TEST INFRASTRUCTURE.  RsConstVeloPrior / RsConstAccelerationPrior
    (video_bundler_rs_inter.h:55-108, 113-173) with the interFrameRatio block constant are linear in
    the four pose blocks: each 6-residual half is  sigma * (c_p0 pose0 + c_e0 end0 + c_p1 pose1 + c_e1 end1)
    with sigma = scale * (0.01, 0.01, 0.01, 1, 1, 1).  Returns coefficient rows [2][4] in the block
    order (pose0, end0, pose1, end1) = (frame k first, frame k last, frame k-1 first, frame k-1 last)."""
    r = float(ratio)
    if kind == PRIOR_VELOCITY:
        a = [1.0, 0.0, r, -(1.0 + r)]                       # pose0 - (end1 + r (end1 - pose1))
        if r > _EPS:
            b = [-(1.0 + 1.0 / r), 1.0, 0.0, 1.0 / r]      # end0 - (pose0 + (pose0 - end1) / r)
        else:
            b = [-1.0, 1.0, 1.0, -1.0]                     # end0 - (pose0 + (end1 - pose1))
    elif kind == PRIOR_ACCELERATION:
        a = [0.5, 0.0, 0.5 * r, -0.5 * (1.0 + r)]
        b = [-0.5 * (1.0 + 1.0 / r), 0.5, 0.0, 0.5 / r]
    else:
        raise ValueError(kind)
    return np.array([a, b])


def motion_prior_eval(kind, scale, ratio, frame_k, frame_km1):
    """Residuals [12] and Jacobian [12, 24] (columns: frame k's 12 parameters, then frame k-1's) of one
    prior from the closed form above."""
    c = motion_prior_coefficients(kind, ratio)
    sigma = scale * np.array([0.01, 0.01, 0.01, 1.0, 1.0, 1.0])
    blocks = [frame_k[:6], frame_k[6:], frame_km1[:6], frame_km1[6:]]
    res = np.zeros(12)
    J = np.zeros((12, 24))
    for h in range(2):
        acc = sum(c[h, b] * blocks[b] for b in range(4))
        res[6 * h:6 * h + 6] = sigma * acc
        for b in range(4):
            J[6 * h + np.arange(6), 6 * b + np.arange(6)] = sigma * c[h, b]
    return res, J


def motion_prior_ratio_column(kind, scale, ratio, frame_k, frame_km1):
    """TEST INFRASTRUCTURE.  d residual / d interFrameRatio [12] of one prior when the ratio block is a free
    parameter (the reference's default, CeresHandler.h:156-180): the residuals are linear in the coefficients
    c(ratio) of motion_prior_coefficients, so the column is the same form with dc/dratio."""
    r = float(ratio)
    h = 1.0 if kind == PRIOR_VELOCITY else 0.5
    a = [0.0, 0.0, h, -h]
    if kind == PRIOR_ACCELERATION or r > _EPS:
        b = [h / (r * r), 0.0, 0.0, -h / (r * r)]
    else:
        b = [0.0, 0.0, 0.0, 0.0]
    dc = np.array([a, b])
    sigma = scale * np.array([0.01, 0.01, 0.01, 1.0, 1.0, 1.0])
    blocks = [frame_k[:6], frame_k[6:], frame_km1[:6], frame_km1[6:]]
    col = np.zeros(12)
    for hh in range(2):
        col[6 * hh:6 * hh + 6] = sigma * sum(dc[hh, k] * blocks[k] for k in range(4))
    return col


def motion_prior_eval_ref(kind, scale, ratio, frame_k, frame_km1):
    """The reference's own functors under Jet autodiff (oracle/_ref).  Returns (valid, residuals [12],
    Jacobian [12, 24] w.r.t. the pose blocks, d residual / d interFrameRatio [12])."""
    lib = ref_lib()
    fn = lib.rsba_ref_velo_prior if kind == PRIOR_VELOCITY else lib.rsba_ref_accel_prior
    fn.restype = C.c_int
    ifr = np.array([float(ratio)])
    p0, e0 = np.ascontiguousarray(frame_k[:6]), np.ascontiguousarray(frame_k[6:])
    p1, e1 = np.ascontiguousarray(frame_km1[:6]), np.ascontiguousarray(frame_km1[6:])
    res = np.zeros(12)
    jac = np.zeros((12, 25))
    ok = fn(float(scale), _ptr(ifr, _dp), _ptr(p0, _dp), _ptr(e0, _dp), _ptr(p1, _dp), _ptr(e1, _dp),
            _ptr(res, _dp), _ptr(jac, _dp))
    return bool(ok), res, jac[:, 1:].copy(), jac[:, 0].copy()


def motion_prior_rows(scene, priors, poses=None, huber=0.0, free_ratio=None):
    """TEST INFRASTRUCTURE.  Residual rows of a list of priors (kind, scale, ratio, frame, prev_frame)
    over the full parameter vector [12 F + 3 P]: returns (J scipy CSR [12 n, 12F+3P], r [12 n], cost).
    With huber > 0 the 12-residual blocks are corrected like every other block (apply_huber).
    free_ratio = value: the shared interFrameRatio block is a parameter -- every prior uses that value and
    gets the column d r / d ratio at parameter 9 of a pseudo-frame behind the real frames (``scene`` must
    then already hold the pseudo-frame as its last frame, see lm_oracle.with_pseudo_frame)."""
    import scipy.sparse as sp
    poses = np.asarray(scene.poses if poses is None else poses, dtype=np.float64).reshape(-1, 12)
    F, P = scene.num_frames, scene.num_points
    rows, cols, vals, res = [], [], [], []
    cost = 0.0
    for i, (kind, scale, ratio, fk, fp) in enumerate(priors):
        if free_ratio is not None:
            ratio = float(free_ratio)
        r, J = motion_prior_eval(kind, scale, ratio, poses[fk], poses[fp])
        s = float(r @ r)
        w = 1.0
        if huber > 0 and s > huber * huber:
            w = np.sqrt(huber / np.sqrt(s))
            s = 2.0 * huber * np.sqrt(s) - huber * huber
        cost += 0.5 * s
        res.append(w * r)
        rr, cc = np.nonzero(J)
        rows.append(12 * i + rr)
        cols.append(np.where(cc < 12, 12 * fk + cc, 12 * fp + (cc - 12)))
        vals.append(w * J[rr, cc])
        if free_ratio is not None:
            col = motion_prior_ratio_column(kind, scale, ratio, poses[fk], poses[fp])
            rows.append(12 * i + np.arange(12))
            cols.append(np.full(12, 12 * (F - 1) + 9))
            vals.append(w * col)
    n = len(priors)
    Jx = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(12 * n, 12 * F + 3 * P))
    return Jx, np.concatenate(res), cost


def pose_prior_rows(scene, priors, poses=None):
    """TEST INFRASTRUCTURE.  GoodPosePrior <6; 6, 6> (CeresHandler.h:55-73, wired at :188-204):
    r = diag(rot x3, pos x3) (prior - pose); the functor is valid iff r[0] < 1.  ``priors`` is a list of
    (frame, which_pose, rotation, position, prior_values[6], constant).  The prior blocks are parameter
    blocks of their own (the reference never fixes them); for the dense restatement they are appended as
    pseudo-frames F, F+1, ... (parameters 0..5 = the prior block, 6..11 constant).  Returns
    (scene' with those frames, pose_mask', J scipy CSR [6 n, 12 (F + n) + 3 P], r [6 n], cost, valid)."""
    import copy
    import scipy.sparse as sp
    F, P, n = scene.num_frames, scene.num_points, len(priors)
    poses = np.asarray(scene.poses if poses is None else poses, dtype=np.float64).reshape(-1, 12)
    sc = copy.copy(scene)
    extra = np.zeros((n, 12))
    mask = np.where(np.asarray(scene.const_frames, dtype=bool), 0xFFF, 0).astype(np.int64)
    mask = np.concatenate([mask, np.full(n, 0xFC0, dtype=np.int64)])
    rows, cols, vals, res = [], [], [], []
    valid = True
    for i, (f, which, rot, pos, val, constant) in enumerate(priors):
        extra[i, :6] = val
        if constant:
            mask[F + i] = 0xFFF
        w = np.array([rot] * 3 + [pos] * 3, dtype=np.float64)
        r = w * (np.asarray(val, dtype=np.float64) - poses[f, 6 * which:6 * which + 6])
        valid = valid and bool(r[0] < 1.0)
        res.append(r)
        rows += [6 * i + np.arange(6)] * 2
        cols += [12 * (F + i) + np.arange(6), 12 * f + 6 * which + np.arange(6)]
        vals += [w, -w]
    sc.poses = np.vstack([poses, extra])
    sc.const_frames = np.concatenate([np.asarray(scene.const_frames, dtype=bool), np.zeros(n, dtype=bool)])
    Jx = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                       shape=(6 * n, 12 * (F + n) + 3 * P))
    rr = np.concatenate(res)
    return sc, mask, Jx, rr, 0.5 * float(rr @ rr), valid


def evaluate_cam_ref(scene, poses=None, points=None, cam=None, nthreads=0):
    """TEST INFRASTRUCTURE.  Uncalibrated variant <2; 9, 6, 6, 3> (VideoSfmBaRs.h:38-49) through the
    reference's own ReprojectionError::operator()(camera, pose, point, residuals) under Jet<24>
    (oracle/_ref).  Returns residuals [N,2], J [N,30], Jcam [N,18] (= [2][9]), valid [N]."""
    poses, points = _prep(scene, poses, points)
    lib = ref_lib()
    lib.rsba_ref_eval_cam.restype = C.c_long
    lib.rsba_ref_eval_cam.argtypes = [C.c_long, _dp, _ip, _ip, _dp, C.c_int, _ip, C.c_int, _dp, _dp, _dp, _dp, _dp, _bp, C.c_int]
    n = scene.num_obs
    res, J, Jc = np.zeros((n, 2)), np.zeros((n, 30)), np.zeros((n, 18))
    valid = np.zeros(n, dtype=np.uint8)
    cam = np.ascontiguousarray(scene.cam if cam is None else cam, dtype=np.float64)
    scan = np.ascontiguousarray(scene.scanlines, dtype=np.int32)
    oxy = np.ascontiguousarray(scene.obs_xy, dtype=np.float64)
    fi = np.ascontiguousarray(scene.obs_frame, dtype=np.int32)
    pi = np.ascontiguousarray(scene.obs_point, dtype=np.int32)
    lib.rsba_ref_eval_cam(n, _ptr(oxy, _dp), _ptr(fi, _ip), _ptr(pi, _ip), _ptr(cam, _dp), int(scene.shutter),
                          _ptr(scan, _ip), int(bool(scene.interpolate_rotation)), _ptr(poses, _dp), _ptr(points, _dp),
                          _ptr(res, _dp), _ptr(J, _dp), _ptr(Jc, _dp), _ptr(valid, _bp), nthreads)
    return res, J, Jc, valid


def intrinsics_jacobian(scene, poses=None, points=None, cam=None):
    """TEST INFRASTRUCTURE.  Closed form of d residual / d (fx fy k1 k2 p1 p2 k3 cx cy) from c2i + distort
    (mat/cam.h:372-395, 49-72), numpy; travels without oracle/_ref.  Returns Jcam [N,18] (zeros for invalid rows)."""
    poses, points = _prep(scene, poses, points)
    cam = np.asarray(scene.cam if cam is None else cam, dtype=np.float64)
    # camera-frame point through the port (value only): recover xp, yp from the residual-free projection
    fx, fy, k1, k2, t1, t2, k3, cx, cy = cam
    # the normalised coordinates are obtained by numerically inverting nothing: recompute them directly
    fr, pt = scene.obs_frame, scene.obs_point
    ox = scene.obs_xy[:, 0]
    poses = poses.reshape(-1, 12)
    points = points.reshape(-1, 3)
    if scene.shutter != 0:
        tau = np.clip((ox - scene.scanlines[0]) / float(scene.scanlines[1] - scene.scanlines[0]), 0.0, 1.0)
    else:
        tau = np.zeros_like(ox)
    p0, p1 = poses[fr, :6], poses[fr, 6:]
    rot = p0[:, :3] + (p1[:, :3] - p0[:, :3]) * tau[:, None] if (scene.shutter != 0 and scene.interpolate_rotation) else p0[:, :3]
    cen = p0[:, 3:] + (p1[:, 3:] - p0[:, 3:]) * tau[:, None]
    q = points[pt] - cen
    th = np.linalg.norm(rot, axis=1)
    small = th * th <= np.finfo(np.float64).eps
    w = rot / np.where(small, 1.0, th)[:, None]
    c, s = np.cos(th)[:, None], np.sin(th)[:, None]
    P = q * c + np.cross(w, q) * s + w * np.sum(w * q, axis=1, keepdims=True) * (1 - c)
    P[small] = (q + np.cross(rot, q))[small]
    ok = ~(P[:, 2] < 1e-8)
    z = np.where(ok, P[:, 2], 1.0)
    xp, yp = P[:, 0] / z, P[:, 1] / z
    r2 = xp * xp + yp * yp
    dist = 1 + r2 * (k1 + r2 * (k2 + r2 * k3))
    xy = xp * yp
    px = dist * xp + (2 * t1 * xy + t2 * (r2 + 2 * xp * xp))
    py = dist * yp + (t1 * (r2 + 2 * yp * yp) + 2 * t2 * xy)
    n = scene.num_obs
    Jc = np.zeros((n, 18))
    Jc[:, 0], Jc[:, 10] = px, py
    Jc[:, 2], Jc[:, 11] = fx * xp * r2, fy * yp * r2
    Jc[:, 3], Jc[:, 12] = fx * xp * r2 ** 2, fy * yp * r2 ** 2
    Jc[:, 4], Jc[:, 13] = fx * 2 * xy, fy * (r2 + 2 * yp * yp)
    Jc[:, 5], Jc[:, 14] = fx * (r2 + 2 * xp * xp), fy * 2 * xy
    Jc[:, 6], Jc[:, 15] = fx * xp * r2 ** 3, fy * yp * r2 ** 3
    Jc[:, 7], Jc[:, 17] = 1.0, 1.0
    Jc[~ok] = 0.0
    return Jc
