"""TEST INFRASTRUCTURE -- numpy/scipy restatement of one Levenberg-Marquardt step and of the
whole trust-region loop as Ceres Solver 1.9.0 runs them for rsba's problem
(``ceres::Solve`` with ``SPARSE_SCHUR``: ``CeresHandler.h:394-419``, ``VideoSfMHandler.cc:579-583``).

PARITY UNPINNED: Ceres is a third-party dependency that is absent from /root/reference
(pinned only by ``.travis.yml:33`` = ceres-solver-1.9.0) and from this container, and the
reference has no test that touches ``ceres::Solve``.  The algorithm below is
restated from Ceres 1.9.0's published implementation
(``trust_region_minimizer.cc``, ``levenberg_marquardt_strategy.cc``, ``schur_eliminator_impl.h``):

* Jacobi scaling: ``scale = 1 / (1 + sqrt(colnorm2(J)))`` computed at the FIRST iteration and
  kept; the minimizer works on ``J' = J diag(scale)``.
* LM diagonal: ``D^2 = clamp(colnorm2(J'), min_lm_diagonal, max_lm_diagonal) / radius``,
  refreshed only after an accepted step (``reuse_diagonal``).
* Linear solve: ``(J'^T J' + D^2) y = J'^T r`` by Schur elimination of the 3x3 point blocks and
  Cholesky of the reduced camera system; ``step' = -y``; ``delta = scale * step'``.
* ``model_cost_change = -m . (r + m / 2)`` with ``m = J' step'``.
* ``rho = (cost - new_cost) / model_cost_change``; accept iff ``rho > min_relative_decrease``;
  accept: ``radius /= max(1/3, 1 - (2 rho - 1)^3)`` (capped at max radius), ``decrease_factor = 2``;
  reject: ``radius /= decrease_factor; decrease_factor *= 2``.
* Invalid steps: a failed factorisation of the reduced camera matrix (not positive definite), a non-finite
  step or ``model_cost_change <= 0`` shrink the radius like a rejected step; ``max_num_consecutive_invalid_steps``
  (5) of them in a row end the solve with FAILURE.
* Termination: parameter tolerance ``|delta| <= 1e-8 (|x| + 1e-8)``, function tolerance
  ``|cost change| < 1e-6 cost``, gradient tolerance ``max|g| <= 1e-10``, max iterations.
* Constant parameter blocks / components are removed from the program (their columns do not
  exist); here their columns are zeroed, their diagonal set to one and their step is zero.

What is checked without Ceres (tests/test_lm_oracle_cpu.py): one step equals the least-squares solution of the
augmented system [J'; D] y = [-r; 0] by dense LAPACK, and the loop reaches the minimum MINPACK's Levenberg-Marquardt
and scipy's trust-region reflective find on the same residuals.  The iterate-by-iterate trajectory stays a restatement.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.linalg
import scipy.sparse as sp


@dataclass
class Options:
    max_num_iterations: int = 50
    initial_trust_region_radius: float = 1e4
    max_trust_region_radius: float = 1e16
    min_trust_region_radius: float = 1e-32
    min_relative_decrease: float = 1e-3
    min_lm_diagonal: float = 1e-6
    max_lm_diagonal: float = 1e32
    function_tolerance: float = 1e-6
    gradient_tolerance: float = 1e-10
    parameter_tolerance: float = 1e-8
    jacobi_scaling: bool = True
    max_num_consecutive_invalid_steps: int = 5


def param_masks(scene, pose_mask=None, point_const=None):
    """active[k] for the 12F + 3P scalar parameters (False = constant)."""
    F, P = scene.num_frames, scene.num_points
    if pose_mask is None:
        pose_mask = np.where(np.asarray(scene.const_frames, dtype=bool), 0xFFF, 0)
    pose_mask = np.asarray(pose_mask, dtype=np.int64)
    bits = (pose_mask[:, None] >> np.arange(12)[None, :]) & 1
    act_c = (bits == 0).reshape(-1)
    act_p = np.ones(3 * P, dtype=bool)
    if point_const is not None:
        act_p = np.repeat(np.asarray(point_const) == 0, 3)
    return act_c, act_p


def sparse_jacobian(scene, J, active):
    """scipy CSR [2N, 12F + 3P] from the per-observation blocks J[N,30]; constant columns zero."""
    N, F = scene.num_obs, scene.num_frames
    fr = scene.obs_frame.astype(np.int64)
    pt = scene.obs_point.astype(np.int64)
    rows = np.repeat(np.arange(2 * N).reshape(N, 2), 15, axis=1).reshape(N, 2, 15)
    cols_c = 12 * fr[:, None] + np.arange(12)[None, :]
    cols_p = 12 * F + 3 * pt[:, None] + np.arange(3)[None, :]
    cols = np.concatenate([cols_c, cols_p], axis=1)                       # [N,15]
    cols = np.broadcast_to(cols[:, None, :], (N, 2, 15))
    vals = np.empty((N, 2, 15))
    vals[:, :, 0:6] = J[:, 0:12].reshape(N, 2, 6)
    vals[:, :, 6:12] = J[:, 12:24].reshape(N, 2, 6)
    vals[:, :, 12:15] = J[:, 24:30].reshape(N, 2, 3)
    vals = vals * active[cols]
    return sp.csr_matrix((vals.reshape(-1), (rows.reshape(-1), cols.reshape(-1))),
                         shape=(2 * N, 12 * F + 3 * scene.num_points))


def jacobi_scale(Js, active, enabled=True):
    n = Js.shape[1]
    if not enabled:
        return np.ones(n)
    col2 = np.asarray(Js.multiply(Js).sum(axis=0)).reshape(-1)
    s = 1.0 / (1.0 + np.sqrt(col2))
    s[~active] = 1.0
    return s


def with_intrinsics_block(scene, J, Jcam):
    """Uncalibrated variant <2; 9, 6, 6, 3>: the shared intrinsics as a pseudo-frame behind the real frames
    (parameters 0..8 = fx fy k1 k2 p1 p2 k3 cx cy, 9..11 constant) -- the parameter ordering of the device
    path.  Returns (scene', pose_mask', extra_columns) for lm_step: scene' has F+1 frames, the extra
    columns are the dense [2N, 12] block of the pseudo-frame."""
    import copy
    F = scene.num_frames
    sc = copy.copy(scene)
    sc.poses = np.vstack([scene.poses, np.concatenate([scene.cam, np.zeros(3)])[None, :]])
    sc.const_frames = np.concatenate([np.asarray(scene.const_frames, dtype=bool), [False]])
    mask = np.where(sc.const_frames, 0xFFF, 0).astype(np.int64)
    mask[F] = 0xE00
    N = scene.num_obs
    cols = np.zeros((2 * N, 12))
    cols[0::2, :9] = Jcam[:, :9]
    cols[1::2, :9] = Jcam[:, 9:]
    return sc, mask, cols


def with_pseudo_frame(scene, ratio=None, free_cam=False):
    """The pseudo-frame of the device path behind the real frames: parameters 0..8 = the shared intrinsics when
    they are free (uncalibrated variant), parameter 9 = the motion priors' interFrameRatio when it is free
    (CeresHandler.h:156-180), everything else constant.  Returns (scene', pose_mask')."""
    import copy
    F = scene.num_frames
    sc = copy.copy(scene)
    row = np.zeros(12)
    if free_cam:
        row[:9] = scene.cam
    if ratio is not None:
        row[9] = ratio
    sc.poses = np.vstack([scene.poses, row[None, :]])
    sc.const_frames = np.concatenate([np.asarray(scene.const_frames, dtype=bool), [False]])
    mask = np.where(sc.const_frames, 0xFFF, 0).astype(np.int64)
    mask[F] = 0xFFF & ~(0x1FF if free_cam else 0) & ~(0x200 if ratio is not None else 0)
    return sc, mask


def lm_step(scene, r, J, radius, opts: Options = Options(), scale=None, pose_mask=None, point_const=None,
            want_S=True, extra=None, cam_cols=None):
    """One linear solve of the LM subproblem.  Returns a dict with the reduced system
    ``S delta_c' = rhs`` (scaled space, constant parameters as identity rows), the unscaled
    step and model_cost_change."""
    F, P = scene.num_frames, scene.num_points
    nc = 12 * F
    act_c, act_p = param_masks(scene, pose_mask, point_const)
    active = np.concatenate([act_c, act_p])
    Js = sparse_jacobian(scene, J, active)
    rr = r.reshape(-1)
    if cam_cols is not None:
        # dense columns of the intrinsics pseudo-frame (the LAST frame of `scene`, see with_intrinsics_block)
        c0 = 12 * (F - 1)
        Js = Js.tolil()
        Js[:, c0:c0 + 12] = cam_cols * active[c0:c0 + 12][None, :]
        Js = Js.tocsr()
    if extra is not None:
        # additional residual blocks (camera-only motion priors): rows over the same parameter vector
        Jx, rx = extra
        Js = sp.vstack([Js, sp.csr_matrix(Jx) @ sp.diags(active.astype(np.float64))]).tocsr()
        rr = np.concatenate([rr, np.asarray(rx).reshape(-1)])
    if scale is None:
        scale = jacobi_scale(Js, active, opts.jacobi_scaling)
    Jp = Js @ sp.diags(scale)
    H = (Jp.T @ Jp).tocsr()
    g = Jp.T @ rr
    diag = H.diagonal()
    D2 = np.clip(diag, opts.min_lm_diagonal, opts.max_lm_diagonal) / radius
    D2[~active] = 1.0                       # identity rows for constant parameters
    B = H[:nc, :nc].toarray() + np.diag(D2[:nc])
    E = H[:nc, nc:].tocsr()
    # 3x3 point blocks
    Cd = H[nc:, nc:].tocsr()
    C = np.zeros((P, 3, 3))
    Cd = Cd.tocoo()
    C[Cd.row // 3, Cd.row % 3, Cd.col % 3] = Cd.data
    C += np.einsum("pi,ij->pij", D2[nc:].reshape(P, 3), np.eye(3))
    Cinv = np.linalg.inv(C)
    Cinv_sp = _block_diag(Cinv)
    ECinv = (E @ Cinv_sp).tocsr()
    S = B - (ECinv @ E.T).toarray()
    rhs_y = g[:nc] - ECinv @ g[nc:]
    solver_ok = True
    try:    # Cholesky, like SparseSchurComplementSolver: a matrix that is not positive definite is a solver FAILURE
        y_c = scipy.linalg.cho_solve(scipy.linalg.cho_factor(S, lower=True), rhs_y)
    except (np.linalg.LinAlgError, ValueError):
        solver_ok = False
        y_c = np.full(nc, np.nan)
    y_p = np.einsum("pij,pj->pi", Cinv, (g[nc:] - E.T @ y_c).reshape(P, 3)).reshape(-1)
    step_s = -np.concatenate([y_c, y_p])
    step_s[~active] = 0.0
    m = Jp @ step_s
    mcc = -float(m @ (rr + 0.5 * m))
    delta = step_s * scale
    return dict(S=S if want_S else None, rhs=-rhs_y, delta_poses=delta[:nc].reshape(F, 12),
                delta_points=delta[nc:].reshape(P, 3), model_cost_change=mcc, scale=scale,
                gradient=(Js.T @ rr), D2=D2, step_scaled=step_s, solver_ok=solver_ok)


def _block_diag(blocks):
    P = blocks.shape[0]
    idx = np.arange(P)
    rows = (3 * idx[:, None, None] + np.arange(3)[None, :, None] + np.zeros((1, 1, 3), dtype=np.int64)).reshape(-1)
    cols = (3 * idx[:, None, None] + np.zeros((1, 3, 1), dtype=np.int64) + np.arange(3)[None, None, :]).reshape(-1)
    return sp.csr_matrix((blocks.reshape(-1), (rows, cols)), shape=(3 * P, 3 * P))


@dataclass
class Summary:
    iterations: int = 0
    num_successful_steps: int = 0
    num_unsuccessful_steps: int = 0
    initial_cost: float = 0.0
    final_cost: float = 0.0
    final_radius: float = 0.0
    termination: str = ""
    usable: bool = True
    trace: list = field(default_factory=list)


def solve(scene, evaluate, opts: Options = Options(), pose_mask=None, point_const=None):
    """The trust-region loop.  ``evaluate(poses, points, jac)`` -> (residuals [N,2], J [N,30] or
    None, valid [N]) is the functor evaluation (oracle.evaluate).  Returns (poses, points, Summary)."""
    F, P = scene.num_frames, scene.num_points
    poses, points = scene.poses.copy(), scene.points.copy()
    act_c, act_p = param_masks(scene, pose_mask, point_const)
    # x of the reduced program: parameter blocks that are not entirely constant
    blk_c = np.repeat(act_c.reshape(2 * F, 6).any(axis=1), 6)
    blk_p = np.repeat(act_p.reshape(P, 3).any(axis=1), 3)

    def xnorm(po, pt):
        return float(np.sqrt(np.sum(po.reshape(-1)[blk_c] ** 2) + np.sum(pt.reshape(-1)[blk_p] ** 2)))

    s = Summary()
    r, J, valid = evaluate(poses, points, True)
    if not valid.all():
        s.usable, s.termination = False, "FAILURE: initial evaluation failed"
        return poses, points, s
    cost = 0.5 * float(np.sum(r * r))
    s.initial_cost = cost
    radius, decrease = opts.initial_trust_region_radius, 2.0
    scale = None
    x_norm = xnorm(poses, points)
    step0 = lm_step(scene, r, J, radius, opts, None, pose_mask, point_const, want_S=False)
    scale = step0["scale"]
    gmax = float(np.max(np.abs(step0["gradient"][np.concatenate([act_c, act_p])]), initial=0.0))
    if gmax <= opts.gradient_tolerance:
        s.final_cost, s.final_radius, s.termination = cost, radius, "CONVERGENCE: gradient tolerance"
        return poses, points, s
    D2_kept = None
    reuse = False
    it = 0
    invalid_in_a_row = 0
    while True:
        if it >= opts.max_num_iterations:
            s.termination = "NO_CONVERGENCE: max iterations"
            break
        it += 1
        st = lm_step(scene, r, J, radius, opts, scale, pose_mask, point_const, want_S=False) if not reuse \
            else _lm_step_reuse(scene, r, J, radius, opts, scale, pose_mask, point_const, D2_kept)
        D2_kept = st["D2_unit"] if "D2_unit" in st else st["D2"] * radius
        mcc = st["model_cost_change"]
        if not st["solver_ok"] or not np.isfinite(mcc) or not (mcc > 0.0):       # invalid step
            invalid_in_a_row += 1
            s.num_unsuccessful_steps += 1
            if invalid_in_a_row >= opts.max_num_consecutive_invalid_steps:
                s.usable, s.termination = False, "FAILURE: successive invalid steps"
                break
            radius /= decrease
            decrease *= 2.0
            reuse = True
            s.trace.append(dict(it=it, cost=cost, accepted=False, radius=radius,
                                reason="model" if st["solver_ok"] else "linear solver"))
            if radius < opts.min_trust_region_radius:
                s.termination = "CONVERGENCE: radius too small"
                break
            continue
        invalid_in_a_row = 0
        new_poses, new_points = poses + st["delta_poses"], points + st["delta_points"]
        r_new, _, v_new = evaluate(new_poses, new_points, False)
        step_norm = float(np.sqrt(np.sum(st["delta_poses"] ** 2) + np.sum(st["delta_points"] ** 2)))
        ok = bool(v_new.all())
        new_cost = 0.5 * float(np.sum(r_new * r_new)) if ok else np.finfo(np.float64).max
        if ok:
            if step_norm <= opts.parameter_tolerance * (x_norm + opts.parameter_tolerance):
                s.termination = "CONVERGENCE: parameter tolerance"
                break
            if abs(cost - new_cost) < opts.function_tolerance * cost:
                s.termination = "CONVERGENCE: function tolerance"
                break
        rho = (cost - new_cost) / mcc
        accepted = ok and rho > opts.min_relative_decrease
        if accepted:
            s.num_successful_steps += 1
            poses, points = new_poses, new_points
            x_norm = xnorm(poses, points)
            radius = min(opts.max_trust_region_radius, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease, reuse = 2.0, False
            r, J, valid = evaluate(poses, points, True)
            cost = 0.5 * float(np.sum(r * r))
            act = np.concatenate([act_c, act_p])
            Js = sparse_jacobian(scene, J, act)
            gmax = float(np.max(np.abs((Js.T @ r.reshape(-1))[act]), initial=0.0))
            s.trace.append(dict(it=it, cost=cost, accepted=True, radius=radius, rho=rho, gmax=gmax,
                                step_norm=step_norm, mcc=mcc))
            if gmax <= opts.gradient_tolerance:
                s.termination = "CONVERGENCE: gradient tolerance"
                break
        else:
            s.num_unsuccessful_steps += 1
            radius /= decrease
            decrease *= 2.0
            reuse = True
            s.trace.append(dict(it=it, cost=cost, accepted=False, radius=radius, rho=rho))
            if radius < opts.min_trust_region_radius:
                s.termination = "CONVERGENCE: radius too small"
                break
    s.iterations = it
    s.final_cost, s.final_radius = cost, radius
    return poses, points, s


def _lm_step_reuse(scene, r, J, radius, opts, scale, pose_mask, point_const, D2_unit):
    """Rejected step: Ceres keeps the (clamped) diagonal and only rescales it by the new radius.
    With an unchanged Jacobian the recomputed diagonal is identical, so this is lm_step."""
    return lm_step(scene, r, J, radius, opts, scale, pose_mask, point_const, want_S=False)
