"""ctypes mirror of the C ABI (``include/rsba_cuda.h``) -- used by tests and ``bench.py``.

The product is ``rsba_b200/lib/librsba_cuda.so``; this module only binds it.  It never
imports ``oracle`` and has no CPU fallback: if the shared library is missing or no CUDA
device is present, construction raises.

``Problem`` mirrors the slice of ``ceres::Problem`` that ``CeresHandler`` uses
(``CeresHandler.h:245-255, 283-300, 335-382, 394-426``):

===============================  =========================================================
reference call                   here
===============================  =========================================================
``RsBundleAdjustment::Create``   ``Problem.add_rs_residual(obs, pose0, pose1, point)``
 + ``AddResidualBlock``
``SetParameterBlockConstant``    ``Problem.set_block_constant(block)``
``SubsetParameterization``       ``Problem.set_subset_constant(pose, [components])``
``problem.Evaluate``             ``Problem.evaluate()``
``ceres::Solve``                 ``Problem.solve(options) -> summary``
===============================  =========================================================
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librsba_cuda.so")

RSBA_OK = 0
ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_NO_DEVICE, ERR_STATE = -1, -2, -3, -4
ERR_EVALUATION_FAILED, ERR_LINEAR_SOLVER, ERR_NCCL, ERR_INTERNAL = -5, -6, -7, -8

# every symbol include/rsba_cuda.h declares (tests check the library exports all of them)
SYMBOLS = [
    "rsba_cuda_create", "rsba_cuda_destroy", "rsba_cuda_last_error", "rsba_cuda_set_stream",
    "rsba_cuda_default_options", "rsba_cuda_set_camera", "rsba_cuda_set_intrinsics_free", "rsba_cuda_add_rs_residual_with_intrinsics", "rsba_cuda_get_camera", "rsba_cuda_get_intrinsics_jacobian",
    "rsba_cuda_set_loss", "rsba_cuda_add_rs_residual",
    "rsba_cuda_add_frame_blocks", "rsba_cuda_add_motion_prior", "rsba_cuda_set_motion_priors", "rsba_cuda_get_prior_residuals",
    "rsba_cuda_set_inter_frame_ratio_block", "rsba_cuda_set_inter_frame_ratio_free", "rsba_cuda_get_inter_frame_ratio",
    "rsba_cuda_get_prior_ratio_jacobian",
    "rsba_cuda_add_pose_prior", "rsba_cuda_set_pose_priors", "rsba_cuda_get_pose_priors",
    "rsba_cuda_set_block_constant", "rsba_cuda_set_subset_constant", "rsba_cuda_set_scene",
    "rsba_cuda_set_parameters", "rsba_cuda_get_parameters", "rsba_cuda_evaluate",
    "rsba_cuda_validate", "rsba_cuda_reproject", "rsba_cuda_evaluate_device", "rsba_cuda_device_buffers", "rsba_cuda_observation_order",
    "rsba_cuda_solve", "rsba_cuda_linearize_and_step", "rsba_cuda_plan_reduced_system", "rsba_cuda_plan_task_graph", "rsba_cuda_reduced_solve", "rsba_cuda_measure_fp64_peak", "rsba_cuda_analyze_structure", "rsba_cuda_structure_array",
    "rsba_cuda_structure_free", "rsba_cuda_sort_observations", "rsba_cuda_pnp_batch", "rsba_cuda_nccl_unique_id",
    "rsba_cuda_comm_init", "rsba_cuda_point_owners", "rsba_cuda_launch_count", "rsba_cuda_stage_ms", "rsba_cuda_version",
    "rsba_cuda_device_count", "rsba_cuda_create_multi", "rsba_cuda_multi_size", "rsba_cuda_multi_handle", "rsba_cuda_multi_solve", "rsba_cuda_destroy_multi",
]


class SolveOptions(C.Structure):
    _fields_ = [
        ("max_num_iterations", C.c_int),
        ("initial_trust_region_radius", C.c_double),
        ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("jacobi_scaling", C.c_int),
        ("huber_loss", C.c_double),
        ("verbose", C.c_int),
        ("dense_cholesky", C.c_int),
        ("reorder_tiles", C.c_int),
        ("max_num_consecutive_invalid_steps", C.c_int),
    ]


class SolveSummary(C.Structure):
    _fields_ = [
        ("usable", C.c_int),
        ("termination", C.c_int),
        ("iterations", C.c_int),
        ("num_successful_steps", C.c_int),
        ("num_unsuccessful_steps", C.c_int),
        ("num_jacobian_evaluations", C.c_int),
        ("num_residual_evaluations", C.c_int),
        ("num_residual_blocks", C.c_long),
        ("num_parameters_reduced", C.c_long),
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("final_radius", C.c_double),
        ("final_gradient_max_norm", C.c_double),
        ("time_total_ms", C.c_double),
        ("time_jacobian_ms", C.c_double),
        ("time_residual_ms", C.c_double),
        ("time_schur_ms", C.c_double),
        ("time_cholesky_ms", C.c_double),
        ("time_update_ms", C.c_double),
        ("time_allreduce_ms", C.c_double),
        ("message", C.c_char * 128),
        ("tile_flops", C.c_double),
        ("reduced_levels", C.c_int),
        ("reduced_tiles", C.c_int),
        ("num_armijo_violations", C.c_int),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["message"] = self.message.decode(errors="replace")
        return d


class RsbaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rsba_cuda error {code}: {msg}")
        self.code = code


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def load_library():
    """dlopen the product library and declare the signatures.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
            "rsba_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.rsba_cuda_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.rsba_cuda_destroy.argtypes = [vp]
    lib.rsba_cuda_destroy.restype = None
    lib.rsba_cuda_last_error.restype = C.c_char_p
    lib.rsba_cuda_version.restype = C.c_char_p
    lib.rsba_cuda_set_stream.argtypes = [vp, vp]
    lib.rsba_cuda_default_options.argtypes = [C.POINTER(SolveOptions)]
    lib.rsba_cuda_default_options.restype = None
    lib.rsba_cuda_set_camera.argtypes = [vp, _dp, C.c_int, _ip, C.c_int]
    lib.rsba_cuda_set_intrinsics_free.argtypes = [vp, C.c_int]
    lib.rsba_cuda_add_rs_residual_with_intrinsics.argtypes = [vp, _dp, vp, vp, vp, vp]
    lib.rsba_cuda_get_camera.argtypes = [vp, vp]
    lib.rsba_cuda_get_intrinsics_jacobian.argtypes = [vp, vp]
    lib.rsba_cuda_set_loss.argtypes = [vp, C.c_double]
    lib.rsba_cuda_add_rs_residual.argtypes = [vp, _dp, vp, vp, vp]
    lib.rsba_cuda_add_frame_blocks.argtypes = [vp, vp, vp]
    lib.rsba_cuda_add_motion_prior.argtypes = [vp, C.c_int, C.c_double, C.c_double, vp, vp, vp, vp]
    lib.rsba_cuda_set_motion_priors.argtypes = [vp, C.c_int, _ip, _dp, _dp, _ip, _ip]
    lib.rsba_cuda_get_prior_residuals.argtypes = [vp, vp]
    lib.rsba_cuda_get_prior_residuals.restype = C.c_long
    lib.rsba_cuda_set_inter_frame_ratio_block.argtypes = [vp, vp]
    lib.rsba_cuda_set_inter_frame_ratio_free.argtypes = [vp, C.c_int, C.c_double]
    lib.rsba_cuda_get_inter_frame_ratio.argtypes = [vp, vp]
    lib.rsba_cuda_get_prior_ratio_jacobian.argtypes = [vp, vp]
    lib.rsba_cuda_get_prior_ratio_jacobian.restype = C.c_long
    lib.rsba_cuda_add_pose_prior.argtypes = [vp, C.c_double, C.c_double, vp, vp]
    lib.rsba_cuda_set_pose_priors.argtypes = [vp, C.c_int, _ip, _ip, _dp, _dp, _dp, vp]
    lib.rsba_cuda_get_pose_priors.argtypes = [vp, vp, vp]
    lib.rsba_cuda_get_pose_priors.restype = C.c_long
    lib.rsba_cuda_set_block_constant.argtypes = [vp, vp]
    lib.rsba_cuda_set_subset_constant.argtypes = [vp, vp, C.c_int, _ip]
    lib.rsba_cuda_set_scene.argtypes = [vp, C.c_long, _dp, _ip, _ip, C.c_int, C.c_int,
                                        C.POINTER(C.c_ushort), C.POINTER(C.c_ubyte)]
    lib.rsba_cuda_set_parameters.argtypes = [vp, vp, vp]
    lib.rsba_cuda_get_parameters.argtypes = [vp, vp, vp]
    lib.rsba_cuda_evaluate.argtypes = [vp, _dp, vp, vp, vp]
    lib.rsba_cuda_validate.argtypes = [vp, C.c_double, C.c_double, vp, vp]
    lib.rsba_cuda_reproject.argtypes = [vp, C.c_long, _ip, _ip, C.c_double, vp, vp]
    lib.rsba_cuda_evaluate_device.argtypes = [vp, C.c_int, _dp, C.POINTER(C.c_long)]
    lib.rsba_cuda_device_buffers.argtypes = [vp] + [C.POINTER(vp)] * 5
    lib.rsba_cuda_observation_order.argtypes = [vp, vp]
    lib.rsba_cuda_observation_order.restype = C.c_long
    lib.rsba_cuda_solve.argtypes = [vp, C.POINTER(SolveOptions), C.POINTER(SolveSummary)]
    lib.rsba_cuda_linearize_and_step.argtypes = [vp, C.POINTER(SolveOptions), C.c_double, vp, vp, vp, vp, _dp]
    lib.rsba_cuda_plan_reduced_system.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, vp] + [vp] * 9
    lib.rsba_cuda_plan_task_graph.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    lib.rsba_cuda_reduced_solve.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            vp, vp, vp, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_float), vp]
    lib.rsba_cuda_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.rsba_cuda_analyze_structure.argtypes = [C.c_long, vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp,
                                                C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.rsba_cuda_structure_array.argtypes = [vp, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    lib.rsba_cuda_structure_array.restype = C.c_long
    lib.rsba_cuda_structure_free.argtypes = [vp]
    lib.rsba_cuda_structure_free.restype = None
    lib.rsba_cuda_sort_observations.argtypes = [C.c_long, vp, vp, C.c_int, C.c_int, vp]
    lib.rsba_cuda_sort_observations.restype = C.c_long
    lib.rsba_cuda_pnp_batch.argtypes = [vp, _dp, C.c_int, _ip, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp,
                                        C.POINTER(SolveOptions), C.c_double, vp, vp, vp, vp]
    lib.rsba_cuda_nccl_unique_id.argtypes = [C.POINTER(C.c_ubyte)]
    lib.rsba_cuda_comm_init.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]
    lib.rsba_cuda_point_owners.argtypes = [C.c_int, C.c_int, C.c_long, vp, vp, C.c_int, vp]
    lib.rsba_cuda_launch_count.argtypes = [vp]
    lib.rsba_cuda_launch_count.restype = C.c_long
    lib.rsba_cuda_stage_ms.argtypes = [vp, C.c_int]
    lib.rsba_cuda_stage_ms.restype = C.c_double
    lib.rsba_cuda_device_count.argtypes = []
    lib.rsba_cuda_create_multi.argtypes = [C.POINTER(C.c_void_p), _ip, C.c_int]
    lib.rsba_cuda_multi_size.argtypes = [vp]
    lib.rsba_cuda_multi_handle.argtypes = [vp, C.c_int]
    lib.rsba_cuda_multi_handle.restype = C.c_void_p
    lib.rsba_cuda_multi_solve.argtypes = [vp, C.POINTER(SolveOptions), C.POINTER(SolveSummary)]
    lib.rsba_cuda_destroy_multi.argtypes = [vp]
    lib.rsba_cuda_destroy_multi.restype = None
    _lib = lib
    return lib


def default_options(**kw) -> SolveOptions:
    o = SolveOptions()
    load_library().rsba_cuda_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def _addr(a):
    """Address of a numpy array / torch tensor / int / None, as c_void_p."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


STAGES = ("jacobian", "residual", "schur", "cholesky", "update", "allreduce", "point_blocks", "frame_blocks",
          "phi_build", "schur_syrk", "schur_reduce", "factor", "tri_solve", "point_step", "finalize", "pnp")


class Problem:
    """One BA problem on one GPU (the analogue of ``ceres::Problem`` + ``ceres::Solve``)."""

    def __init__(self, device: int = 0, _borrowed=None):
        self.lib = load_library()
        h = C.c_void_p()
        self._h = None
        self._owned = _borrowed is None
        if self._owned:
            self._check(self.lib.rsba_cuda_create(C.byref(h), int(device)))
        else:
            h = C.c_void_p(_borrowed)     # a rank of a MultiProblem: destroyed with it
        self._h = h
        self._keep = []          # arrays whose memory the pointer API refers to
        self.num_obs = 0
        self.num_frames = 0
        self.num_points = 0

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc, allow=()):
        if rc != RSBA_OK and rc not in allow:
            raise RsbaError(rc, self.lib.rsba_cuda_last_error().decode(errors="replace"))
        return rc

    def close(self):
        if self._h is not None:
            if self._owned:
                self.lib.rsba_cuda_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.rsba_cuda_set_stream(self._h, C.c_void_p(cuda_stream)))

    # ------------------------------------------------------------------ construction
    def set_camera(self, cam, shutter, scanlines, interpolate_rotation=True):
        cam = np.ascontiguousarray(cam, dtype=np.float64)
        scan = np.ascontiguousarray(scanlines, dtype=np.int32)
        assert cam.size == 9 and scan.size == 2
        self._check(self.lib.rsba_cuda_set_camera(self._h, cam.ctypes.data_as(_dp), int(shutter),
                                                  scan.ctypes.data_as(_ip), int(bool(interpolate_rotation))))

    def set_intrinsics_free(self, free=True):
        """Uncalibrated variant <2; 9, 6, 6, 3> (VideoSfmBaRs.h:38-49): the shared intrinsics are optimised."""
        self._check(self.lib.rsba_cuda_set_intrinsics_free(self._h, int(bool(free))))
        self.free_intrinsics = bool(free)

    def get_camera(self):
        cam = np.zeros(9)
        self._check(self.lib.rsba_cuda_get_camera(self._h, _addr(cam)))
        return cam

    def intrinsics_jacobian(self, num_obs=None):
        n = self.num_obs if num_obs is None else num_obs
        Jc = np.zeros((n, 18))
        self._check(self.lib.rsba_cuda_get_intrinsics_jacobian(self._h, _addr(Jc)))
        return Jc

    def set_loss(self, huber_a: float):
        """``ceres::HuberLoss(huber_a)`` on every residual block (CeresHandler.h:85-90); 0 = none."""
        self._check(self.lib.rsba_cuda_set_loss(self._h, float(huber_a)))

    def add_rs_residual(self, observed, pose0: np.ndarray, pose1: np.ndarray, point: np.ndarray):
        """``pose0``/``pose1``/``point`` are float64 numpy views; identity = their address."""
        obs = np.ascontiguousarray(observed, dtype=np.float64)
        for a in (pose0, pose1, point):
            assert a.dtype == np.float64 and a.flags.c_contiguous
        self._keep.extend((pose0, pose1, point))
        self._check(self.lib.rsba_cuda_add_rs_residual(self._h, obs.ctypes.data_as(_dp), _addr(pose0),
                                                       _addr(pose1), _addr(point)))

    def add_frame_blocks(self, pose0: np.ndarray, pose1: np.ndarray):
        """ceres::Problem::AddParameterBlock for a frame's two control poses (a frame with pose priors only)."""
        self._keep += [pose0, pose1]
        self._check(self.lib.rsba_cuda_add_frame_blocks(self._h, _addr(pose0), _addr(pose1)))

    def add_motion_prior(self, kind, scale, ratio, pose0, end0, pose1, end1):
        """RsConstVeloPrior (kind 1) / RsConstAccelerationPrior (kind 2) between frame k = (pose0, end0)
        and frame k-1 = (pose1, end1), constant interFrameRatio (CeresHandler.h:148-186)."""
        self._keep.extend((pose0, end0, pose1, end1))
        self._check(self.lib.rsba_cuda_add_motion_prior(self._h, int(kind), float(scale), float(ratio), _addr(pose0),
                                                        _addr(end0), _addr(pose1), _addr(end1)))

    def set_motion_priors(self, kind, scale, ratio, frame, prev_frame):
        """Bulk form: arrays of equal length; ``frame`` / ``prev_frame`` index the pose array."""
        kind = np.ascontiguousarray(kind, dtype=np.int32)
        scale = np.ascontiguousarray(scale, dtype=np.float64)
        ratio = np.ascontiguousarray(ratio, dtype=np.float64)
        frame = np.ascontiguousarray(frame, dtype=np.int32)
        prev = np.ascontiguousarray(prev_frame, dtype=np.int32)
        self._check(self.lib.rsba_cuda_set_motion_priors(self._h, int(kind.size), kind.ctypes.data_as(_ip),
                                                         scale.ctypes.data_as(_dp), ratio.ctypes.data_as(_dp),
                                                         frame.ctypes.data_as(_ip), prev.ctypes.data_as(_ip)))

    def set_inter_frame_ratio_free(self, free=True, value=1.0):
        """The interFrameRatio block of the motion priors as a free, lower-bounded parameter (the reference's
        default, CeresHandler.h:156-180); bulk form."""
        self._check(self.lib.rsba_cuda_set_inter_frame_ratio_free(self._h, int(bool(free)), float(value)))
        self.free_ratio = bool(free)

    def set_inter_frame_ratio_block(self, ratio):
        """Pointer form: ``ratio`` is a float64 numpy array of one element (&opt.ceres.interFrameRatio)."""
        assert ratio.dtype == np.float64 and ratio.size == 1
        self._keep.append(ratio)
        self._check(self.lib.rsba_cuda_set_inter_frame_ratio_block(self._h, _addr(ratio)))
        self.free_ratio = True

    def inter_frame_ratio(self):
        v = C.c_double(0.0)
        self._check(self.lib.rsba_cuda_get_inter_frame_ratio(self._h, C.byref(v)))
        return v.value

    def prior_ratio_jacobian(self):
        n = self.lib.rsba_cuda_get_prior_ratio_jacobian(self._h, None)
        if n < 0:
            raise RsbaError(ERR_STATE, self.lib.rsba_cuda_last_error().decode())
        j = np.zeros((n, 12))
        if n > 0:
            self.lib.rsba_cuda_get_prior_ratio_jacobian(self._h, _addr(j))
        return j

    def add_pose_prior(self, rotation, position, prior_block, pose_block):
        """GoodPosePrior between a 6-wide prior block and a control-pose block (CeresHandler.h:188-204)."""
        assert prior_block.dtype == np.float64 and prior_block.size == 6 and prior_block.flags.c_contiguous
        self._keep.extend((prior_block, pose_block))
        self._check(self.lib.rsba_cuda_add_pose_prior(self._h, float(rotation), float(position), _addr(prior_block),
                                                      _addr(pose_block)))

    def set_pose_priors(self, frame, which_pose, rotation, position, prior_values, prior_constant=None):
        frame = np.ascontiguousarray(frame, dtype=np.int32)
        which = np.ascontiguousarray(which_pose, dtype=np.int32)
        rot = np.ascontiguousarray(rotation, dtype=np.float64)
        pos = np.ascontiguousarray(position, dtype=np.float64)
        val = np.ascontiguousarray(prior_values, dtype=np.float64).reshape(-1)
        assert val.size == 6 * frame.size
        cst = None if prior_constant is None else np.ascontiguousarray(prior_constant, dtype=np.uint8)
        self._check(self.lib.rsba_cuda_set_pose_priors(self._h, int(frame.size), frame.ctypes.data_as(_ip),
                                                       which.ctypes.data_as(_ip), rot.ctypes.data_as(_dp),
                                                       pos.ctypes.data_as(_dp), val.ctypes.data_as(_dp),
                                                       None if cst is None else _addr(cst)))

    def pose_priors(self):
        """(current values [n, 6], trial values of the last LM step [n, 6]) of the prior blocks."""
        n = self.lib.rsba_cuda_get_pose_priors(self._h, None, None)
        val, trial = np.zeros((max(n, 0), 6)), np.zeros((max(n, 0), 6))
        if n > 0:
            self.lib.rsba_cuda_get_pose_priors(self._h, _addr(val), _addr(trial))
        return val, trial

    def reproject(self, frame, point, sqrd_threshold=16.0):
        """Iterative rolling-shutter re-projection of (frame, point) pairs (struct/VideoSfM.cc:139-155).
        Returns (proj_xy [n, 2], ok [n])."""
        frame = np.ascontiguousarray(frame, dtype=np.int32)
        point = np.ascontiguousarray(point, dtype=np.int32)
        n = int(frame.size)
        xy, ok = np.zeros((n, 2)), np.zeros(n, dtype=np.uint8)
        self._check(self.lib.rsba_cuda_reproject(self._h, C.c_long(n), frame.ctypes.data_as(_ip), point.ctypes.data_as(_ip),
                                                 C.c_double(sqrd_threshold), _addr(xy), _addr(ok)))
        return xy, ok

    def prior_residuals(self):
        n = self.lib.rsba_cuda_get_prior_residuals(self._h, None)
        r = np.zeros((max(n, 0), 12))
        if n > 0:
            self.lib.rsba_cuda_get_prior_residuals(self._h, _addr(r))
        return r

    def add_rs_residual_with_intrinsics(self, observed, cam, pose0, pose1, point):
        """<2; 9, 6, 6, 3>: ``cam`` is the shared 9-wide intrinsics block (CeresHandler.h:256-264)."""
        obs = np.ascontiguousarray(observed, dtype=np.float64)
        self._keep.extend((cam, pose0, pose1, point))
        self._check(self.lib.rsba_cuda_add_rs_residual_with_intrinsics(self._h, obs.ctypes.data_as(_dp), _addr(cam),
                                                                       _addr(pose0), _addr(pose1), _addr(point)))
        self.free_intrinsics = True

    def set_block_constant(self, block: np.ndarray):
        self._check(self.lib.rsba_cuda_set_block_constant(self._h, _addr(block)))

    def set_subset_constant(self, pose_block: np.ndarray, components):
        comp = np.ascontiguousarray(components, dtype=np.int32)
        self._check(self.lib.rsba_cuda_set_subset_constant(self._h, _addr(pose_block), int(comp.size),
                                                           comp.ctypes.data_as(_ip)))

    def set_scene(self, obs_xy, obs_frame, obs_point, num_frames, num_points, const_pose_mask=None,
                  const_point=None):
        xy = np.ascontiguousarray(obs_xy, dtype=np.float64)
        fr = np.ascontiguousarray(obs_frame, dtype=np.int32)
        pt = np.ascontiguousarray(obs_point, dtype=np.int32)
        n = fr.shape[0]
        assert xy.size == 2 * n and pt.shape[0] == n
        pm = None if const_pose_mask is None else np.ascontiguousarray(const_pose_mask, dtype=np.uint16)
        pc = None if const_point is None else np.ascontiguousarray(const_point, dtype=np.uint8)
        self._check(self.lib.rsba_cuda_set_scene(
            self._h, n, xy.ctypes.data_as(_dp), fr.ctypes.data_as(_ip), pt.ctypes.data_as(_ip),
            int(num_frames), int(num_points),
            None if pm is None else pm.ctypes.data_as(C.POINTER(C.c_ushort)),
            None if pc is None else pc.ctypes.data_as(C.POINTER(C.c_ubyte))))
        self.num_obs, self.num_frames, self.num_points = n, int(num_frames), int(num_points)

    def load_scene(self, scene, poses=None, points=None):
        """Bulk construction from a :class:`rsba_b200.scene.Scene` (frame-constant flags become
        ``SetParameterBlockConstant`` on both pose blocks, CeresHandler.h:342-346)."""
        self.set_camera(scene.cam, scene.shutter, scene.scanlines, scene.interpolate_rotation)
        mask = np.where(np.asarray(scene.const_frames, dtype=bool), 0xFFF, 0).astype(np.uint16)
        self.set_scene(scene.obs_xy, scene.obs_frame, scene.obs_point, scene.num_frames, scene.num_points,
                       const_pose_mask=mask)
        self.set_parameters(scene.poses if poses is None else poses,
                            scene.points if points is None else points)

    def set_parameters(self, poses, points):
        """Host arrays (numpy, or pinned torch CPU tensors) -> device."""
        if isinstance(poses, np.ndarray):
            poses = np.ascontiguousarray(poses, dtype=np.float64)
        if isinstance(points, np.ndarray):
            points = np.ascontiguousarray(points, dtype=np.float64)
        self._check(self.lib.rsba_cuda_set_parameters(self._h, _addr(poses), _addr(points)))

    def get_parameters(self, poses=None, points=None):
        if poses is None:
            poses = np.empty((self.num_frames, 12))
        if points is None:
            points = np.empty((self.num_points, 3))
        self._check(self.lib.rsba_cuda_get_parameters(self._h, _addr(poses), _addr(points)))
        return poses, points

    # ------------------------------------------------------------------ evaluation
    def evaluate(self, residuals=True, jacobian=True, valid=True, num_obs=None, check=True):
        """Host outputs in the caller's observation order.
        Returns (cost, residuals [N,2], jacobian [N,30], valid [N])."""
        n = self.num_obs if num_obs is None else num_obs
        cost = C.c_double(0.0)
        r = np.zeros((n, 2)) if residuals else None
        J = np.zeros((n, 30)) if jacobian else None
        v = np.zeros(n, dtype=np.uint8) if valid else None
        rc = self.lib.rsba_cuda_evaluate(self._h, C.byref(cost), _addr(r), _addr(J), _addr(v))
        self._check(rc, allow=() if check else (ERR_EVALUATION_FAILED,))
        return cost.value, r, J, v

    def validate(self, sqrd_threshold=16.0, min_distance_to_camera=0.0, num_obs=None):
        """Track-validation sweep (struct/VideoSfM.cc:159-169): (ok [N] uint8, squared error [N])."""
        n = self.num_obs if num_obs is None else num_obs
        ok = np.zeros(n, dtype=np.uint8)
        err = np.zeros(n)
        self._check(self.lib.rsba_cuda_validate(self._h, float(sqrd_threshold), float(min_distance_to_camera),
                                                _addr(ok), _addr(err)))
        return ok, err

    def evaluate_device(self, with_jacobian=True, fetch=True):
        """HBM-resident evaluation; with ``fetch`` returns (cost, num_invalid) (synchronises)."""
        if fetch:
            cost, bad = C.c_double(0.0), C.c_long(0)
            self._check(self.lib.rsba_cuda_evaluate_device(self._h, int(with_jacobian), C.byref(cost), C.byref(bad)))
            return cost.value, bad.value
        self._check(self.lib.rsba_cuda_evaluate_device(self._h, int(with_jacobian), None, None))
        return None

    def device_buffers(self):
        ptrs = [C.c_void_p() for _ in range(5)]
        self._check(self.lib.rsba_cuda_device_buffers(self._h, *[C.byref(p) for p in ptrs]))
        return dict(zip(("residuals", "jacobian", "valid", "poses", "points"), [p.value for p in ptrs]))

    def observation_order(self):
        n = self.lib.rsba_cuda_observation_order(self._h, None)
        order = np.empty(n, dtype=np.int64)
        self.lib.rsba_cuda_observation_order(self._h, _addr(order))
        return order

    # ------------------------------------------------------------------ solve
    def solve(self, options: SolveOptions | None = None, check=True) -> SolveSummary:
        if options is None:
            options = default_options()
        s = SolveSummary()
        rc = self.lib.rsba_cuda_solve(self._h, C.byref(options), C.byref(s))
        if check:
            self._check(rc)
        s.rc = rc
        return s

    def linearize_and_step(self, radius, options: SolveOptions | None = None, want_S=True, fetch=True):
        """One linearisation + LM step at the current parameters (does not move them).
        Returns dict(S, rhs, delta_poses, delta_points, model_cost_change); with ``fetch=False``
        nothing is copied back (profiling)."""
        if options is None:
            options = default_options()
        if not fetch:
            mcc = C.c_double(0.0)
            self._check(self.lib.rsba_cuda_linearize_and_step(self._h, C.byref(options), float(radius), None,
                                                              None, None, None, C.byref(mcc)))
            return dict(model_cost_change=mcc.value)
        # + the pseudo-frame (free intrinsics = parameters 0..8, free interFrameRatio = parameter 9)
        nf = self.num_frames + (1 if getattr(self, "free_intrinsics", False) or getattr(self, "free_ratio", False) else 0)
        n = 12 * nf
        S = np.zeros((n, n)) if want_S else None
        rhs = np.zeros(n)
        dposes = np.zeros((nf, 12))
        dpoints = np.zeros((self.num_points, 3))
        mcc = C.c_double(0.0)
        self._check(self.lib.rsba_cuda_linearize_and_step(self._h, C.byref(options), float(radius), _addr(S),
                                                          _addr(rhs), _addr(dposes), _addr(dpoints), C.byref(mcc)))
        return dict(S=S, rhs=rhs, delta_poses=dposes, delta_points=dpoints, model_cost_change=mcc.value)

    # ------------------------------------------------------------------ batched RS-PnP
    def pnp_batch(self, cam, shutter, scanlines, points3d, obs_xy, sample_idx, poses, options=None,
                  inlier_threshold=8.0):
        """Batched solveRsPnP inner solve + inlier scoring (solveRSpnp.cpp:100-192, 265-335).
        ``poses`` [H, 12] initial -> returns dict(poses, cost, usable, iterations, inliers)."""
        if options is None:
            options = default_options(max_num_iterations=10)
        cam = np.ascontiguousarray(cam, dtype=np.float64)
        scan = np.ascontiguousarray(scanlines, dtype=np.int32)
        pts = np.ascontiguousarray(points3d, dtype=np.float64)
        xy = np.ascontiguousarray(obs_xy, dtype=np.float64)
        idx = np.ascontiguousarray(sample_idx, dtype=np.int32)
        out = np.ascontiguousarray(poses, dtype=np.float64).copy()
        H = idx.shape[0]
        cost, usable = np.zeros(H), np.zeros(H, dtype=np.int32)
        its, inl = np.zeros(H, dtype=np.int32), np.zeros(H, dtype=np.int32)
        self._check(self.lib.rsba_cuda_pnp_batch(self._h, cam.ctypes.data_as(_dp), int(shutter), scan.ctypes.data_as(_ip),
                                                 pts.shape[0], _addr(pts), _addr(xy), H, idx.shape[1], _addr(idx),
                                                 _addr(out), C.byref(options), float(inlier_threshold), _addr(cost),
                                                 _addr(usable), _addr(its), _addr(inl)))
        return dict(poses=out, cost=cost, usable=usable, iterations=its, inliers=inl)

    # ------------------------------------------------------------------ multi-GPU / introspection
    def comm_init(self, rank: int, world_size: int, unique_id: bytes):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._check(self.lib.rsba_cuda_comm_init(self._h, int(rank), int(world_size), buf))

    def launch_count(self) -> int:
        return int(self.lib.rsba_cuda_launch_count(self._h))

    def stage_ms(self, stage) -> float:
        idx = STAGES.index(stage) if isinstance(stage, str) else int(stage)
        return float(self.lib.rsba_cuda_stage_ms(self._h, idx))


class MultiProblem:
    """One host thread, N GPUs of one node (``rsba_cuda_create_multi``): ``ranks[r]`` are :class:`Problem` views of
    the per-device handles -- builder calls go to every one of them (``each``) -- and ``solve`` runs the ranks' LM
    loops on worker threads inside the library.  Results are read from ``ranks[0]``."""

    def __init__(self, devices):
        self.lib = load_library()
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        m = C.c_void_p()
        self._m = None
        rc = self.lib.rsba_cuda_create_multi(C.byref(m), dev.ctypes.data_as(_ip), int(dev.size))
        if rc != RSBA_OK:
            raise RsbaError(rc, self.lib.rsba_cuda_last_error().decode(errors="replace"))
        self._m = m
        self.ranks = [Problem(_borrowed=self.lib.rsba_cuda_multi_handle(m, r)) for r in range(self.lib.rsba_cuda_multi_size(m))]

    def each(self, fn):
        return [fn(pb) for pb in self.ranks]

    def load_scene(self, scene, poses=None, points=None):
        self.each(lambda pb: pb.load_scene(scene, poses, points))

    def solve(self, options: SolveOptions | None = None, check=True) -> SolveSummary:
        if options is None:
            options = default_options()
        s = SolveSummary()
        rc = self.lib.rsba_cuda_multi_solve(self._m, C.byref(options), C.byref(s))
        if check and rc != RSBA_OK:
            raise RsbaError(rc, self.lib.rsba_cuda_last_error().decode(errors="replace"))
        s.rc = rc
        return s

    def get_parameters(self):
        return self.ranks[0].get_parameters()

    def close(self):
        if self._m is not None:
            for pb in self.ranks:
                pb.close()
            self.lib.rsba_cuda_destroy_multi(self._m)
            self._m = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def plan_reduced_system(n_tiles, pair_a, pair_b, dense=False, reorder=True):
    """Host-only symbolic analysis of the reduced camera system (no GPU needed): ordering,
    fill, elimination levels, conflict-free update groups.  Returns a dict of numpy arrays."""
    lib = load_library()
    pa = np.ascontiguousarray(pair_a, dtype=np.int32)
    pb = np.ascontiguousarray(pair_b, dtype=np.int32)
    counts = np.zeros(6, dtype=np.int64)

    def call(*outs):
        rc = lib.rsba_cuda_plan_reduced_system(int(n_tiles), int(pa.size), _addr(pa), _addr(pb), int(dense),
                                               int(reorder), _addr(counts), *[_addr(o) for o in outs])
        if rc != RSBA_OK:
            raise RsbaError(rc, lib.rsba_cuda_last_error().decode(errors="replace"))

    call(*([None] * 9))
    L, nnz, ntrsm, nupd, ngroups = (int(c) for c in counts[:5])
    out = dict(tile_pos=np.zeros(n_tiles, np.int32), nz_tiles=np.zeros((nnz, 2), np.int32),
               panels=np.zeros(n_tiles, np.int32), panel_ptr=np.zeros(L + 1, np.int32),
               trsm=np.zeros((ntrsm, 2), np.int32), trsm_ptr=np.zeros(L + 1, np.int32),
               upd=np.zeros((nupd, 3), np.int32), group_ptr=np.zeros(ngroups + 1, np.int64),
               level_group_ptr=np.zeros(L + 1, np.int32))
    call(*out.values())
    out.update(n_levels=L, flops=float(counts[5]))
    return out


TASK_FACTOR, TASK_TRSM, TASK_UPDATE, TASK_BACK_FIN, TASK_BACK_TILE = range(5)


def plan_task_graph(n_tiles, pair_a, pair_b, dense=False, reorder=True, merge_levels=4):
    """Host-only: the numeric phase of the reduced-system solve as the static task list of the persistent
    kernel (rsba_b200/csrc/k3_dag.cu).  Returns dict(tasks[n, 8], sources[m, 2], need[n_nz, 4],
    n_factor_tasks)."""
    lib = load_library()
    pa = np.ascontiguousarray(pair_a, dtype=np.int32)
    pb = np.ascontiguousarray(pair_b, dtype=np.int32)
    counts = np.zeros(4, dtype=np.int64)

    def call(*outs):
        rc = lib.rsba_cuda_plan_task_graph(int(n_tiles), int(pa.size), _addr(pa), _addr(pb), int(dense), int(reorder),
                                           int(merge_levels), _addr(counts), *[_addr(o) for o in outs])
        if rc != RSBA_OK:
            raise RsbaError(rc, lib.rsba_cuda_last_error().decode(errors="replace"))

    call(None, None, None)
    tasks = np.zeros((int(counts[0]), 8), np.int32)
    sources = np.zeros((max(int(counts[1]), 1), 2), np.int32)
    need = np.zeros((max(int(counts[2]), 1), 4), np.int32)
    call(tasks, sources, need)
    return dict(tasks=tasks, sources=sources[:int(counts[1])], need=need[:int(counts[2])],
                n_factor_tasks=int(counts[3]))


def reduced_solve(A, rhs, n_tiles, pair_a=(), pair_b=(), dense=False, reorder=True, mode="dag", merge_levels=4,
                  want_L=False, device=0, repeats=1, want_trace=False):
    """K3 on its own, on the GPU: solves A x = rhs for an SPD matrix with the given tile pattern (tile = 96 rows).
    Returns dict(x, info, ms[, L, tile_pos]); L is in the permuted tile order that tile_pos describes."""
    lib = load_library()
    n = 96 * int(n_tiles)
    A = np.ascontiguousarray(A, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    assert A.shape == (n, n) and rhs.shape == (n,)
    pa = np.ascontiguousarray(pair_a, dtype=np.int32)
    pb = np.ascontiguousarray(pair_b, dtype=np.int32)
    x = np.zeros(n)
    L = np.zeros((n, n)) if want_L else None
    pos = np.zeros(int(n_tiles), dtype=np.int32)
    info, ms = C.c_int(0), C.c_float(0)
    trace = None
    if want_trace:
        n_tasks = len(plan_task_graph(n_tiles, pa, pb, dense, reorder, merge_levels)["tasks"])
        trace = np.zeros((n_tasks, 16), dtype=np.int64)
    rc = lib.rsba_cuda_reduced_solve(int(device), int(n_tiles), int(pa.size), _addr(pa), _addr(pb), int(dense),
                                     int(reorder), {"dag": 0, "levels": 1}[mode], int(merge_levels), int(repeats),
                                     _addr(A), _addr(rhs), _addr(x), _addr(L), _addr(pos), C.byref(info), C.byref(ms),
                                     _addr(trace))
    if rc != RSBA_OK:
        raise RsbaError(rc, lib.rsba_cuda_last_error().decode(errors="replace"))
    out = dict(x=x, info=info.value, ms=ms.value, tile_pos=pos)
    if want_L:
        out["L"] = L
    if want_trace:
        out["trace"] = trace
    return out


def measure_fp64_peak(device=0):
    """(DFMA, DMMA) TFLOP/s of the device, measured now (rsba_cuda_measure_fp64_peak)."""
    lib = load_library()
    a, b = C.c_double(0), C.c_double(0)
    rc = lib.rsba_cuda_measure_fp64_peak(int(device), C.byref(a), C.byref(b))
    if rc != RSBA_OK:
        raise RsbaError(rc, lib.rsba_cuda_last_error().decode(errors="replace"))
    return a.value, b.value


_STRUCTURE_ARRAYS = {
    "pt_ptr": (np.int32, 1), "pt_obs": (np.int32, 1), "chunk_frame": (np.int32, 1), "chunk_beg": (np.int32, 1),
    "chunk_cnt": (np.int32, 1), "frame_chunk_ptr": (np.int32, 1), "inc_point": (np.int32, 1), "inc_tile": (np.int32, 1),
    "slot_beg": (np.int32, 4), "slot_cnt": (np.uint8, 4), "pt_inc_ptr": (np.int32, 1), "cam_inc": (np.int32, 1),
    "inc_half": (np.uint8, 1), "obs_phi_off": (np.int32, 1), "dup_inc": (np.int32, 1), "pair_a": (np.int32, 1),
    "pair_b": (np.int32, 1), "pair_item_ptr": (np.int32, 1), "items": (np.int32, 4), "entries": (np.int32, 2),
    "fwd_slot": (np.int32, 1), "plan.tile_pos": (np.int32, 1), "plan.pos_tile": (np.int32, 1),
    "plan.nz_tiles": (np.int32, 2), "plan.tile_slot": (np.int32, 1), "plan.panels": (np.int32, 1),
    "plan.panel_ptr": (np.int32, 1), "plan.trsm": (np.int32, 2), "plan.trsm_ptr": (np.int32, 1),
    "plan.upd": (np.int32, 4), "plan.lrow_ptr": (np.int32, 1), "plan.lrow_cols": (np.int32, 1),
    "local_ids": (np.int64, 1), "point_owned": (np.uint8, 1),
    "point_groups": (np.int32, 2), "point_big": (np.int32, 1),
}


def analyze_structure(obs_frame, obs_point, n_frames, n_points, const_point=None, free_intrinsics=False,
                      free_ratio=False, prior_frame=(), prior_prev=(), dense=False, reorder=True, sparse_keys=False,
                      rank=0, world_size=1):
    """Host-only: the one-off structure analysis of rsba_cuda_solve (point CSR, frame chunks, Schur incidences,
    sub-tile pairs, work items and entry lists, tile plan) as a dict of numpy arrays.  No GPU needed.
    Observations must be sorted by frame.  ``world_size`` > 1: the share of ``rank`` (multi-GPU sharding)."""
    lib = load_library()
    fr = np.ascontiguousarray(obs_frame, dtype=np.int32)
    pt = np.ascontiguousarray(obs_point, dtype=np.int32)
    cp = None if const_point is None else np.ascontiguousarray(const_point, dtype=np.uint8)
    pf = np.ascontiguousarray(prior_frame, dtype=np.int32)
    pp = np.ascontiguousarray(prior_prev, dtype=np.int32)
    handle = C.c_void_p()
    rc = lib.rsba_cuda_analyze_structure(fr.size, _addr(fr), _addr(pt), int(n_frames), int(n_points), _addr(cp),
                                         int(free_intrinsics), int(free_ratio), int(pf.size), _addr(pf), _addr(pp),
                                         int(dense), int(reorder), int(sparse_keys), int(rank), int(world_size),
                                         C.byref(handle))
    if rc != RSBA_OK:
        raise RsbaError(rc, lib.rsba_cuda_last_error().decode(errors="replace"))
    try:
        out = {}
        for name, (dtype, width) in _STRUCTURE_ARRAYS.items():
            data, eb = C.c_void_p(), C.c_int()
            n = lib.rsba_cuda_structure_array(handle, name.encode(), C.byref(data), C.byref(eb))
            if n < 0:
                raise RsbaError(ERR_INVALID_ARGUMENT, name)
            itemsize = np.dtype(dtype).itemsize
            per = eb.value // itemsize if eb.value else 1
            if n == 0 or not data.value:
                arr = np.zeros((0, max(per, width)) if max(per, width) > 1 else 0, dtype)
            else:
                arr = np.ctypeslib.as_array(C.cast(data, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), (n * per,)).copy()
                if per > 1:
                    arr = arr.reshape(n, per)
                elif width > 1:
                    arr = arr.reshape(-1, width)
            out[name] = arr
        for name in ("T", "n_inc", "n_items"):
            out[name] = int(lib.rsba_cuda_structure_array(handle, name.encode(), None, None))
        return out
    finally:
        lib.rsba_cuda_structure_free(handle)


def sort_observations(obs_frame, obs_point, n_frames, n_points) -> np.ndarray:
    """Host-only: sorted position -> caller's observation index (stable sort by frame), as the library
    orders a scene internally."""
    lib = load_library()
    fr = np.ascontiguousarray(obs_frame, dtype=np.int32)
    pt = np.ascontiguousarray(obs_point, dtype=np.int32)
    order = np.zeros(fr.size, dtype=np.int64)
    if lib.rsba_cuda_sort_observations(fr.size, _addr(fr), _addr(pt), int(n_frames), int(n_points), _addr(order)) < 0:
        raise RsbaError(ERR_INVALID_ARGUMENT, lib.rsba_cuda_last_error().decode(errors="replace"))
    return order


def point_owners(scene, world_size: int) -> np.ndarray:
    """Host-only: owner rank of every point under the multi-GPU sharding rule."""
    lib = load_library()
    fr = np.ascontiguousarray(scene.obs_frame, dtype=np.int32)
    pt = np.ascontiguousarray(scene.obs_point, dtype=np.int32)
    owner = np.zeros(scene.num_points, dtype=np.int32)
    rc = lib.rsba_cuda_point_owners(scene.num_frames, scene.num_points, fr.size, _addr(fr), _addr(pt),
                                    int(world_size), _addr(owner))
    if rc != RSBA_OK:
        raise RsbaError(rc, lib.rsba_cuda_last_error().decode(errors="replace"))
    return owner


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = (C.c_ubyte * 128)()
    rc = lib.rsba_cuda_nccl_unique_id(buf)
    if rc != RSBA_OK:
        raise RsbaError(rc, lib.rsba_cuda_last_error().decode(errors="replace"))
    return bytes(buf)
