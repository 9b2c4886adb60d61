"""Shared test inputs: the small synthetic scenes plus a hand-built edge-case scene."""
from __future__ import annotations

import numpy as np

from rsba_b200.scene import Scene, make_scene, DEFAULT_CAM


def small_scene(**kw) -> Scene:
    """BASELINE config C1: 10 frames / 500 points / 5 000 observations."""
    return make_scene(10, 500, 10, name="C1", **kw)


def edge_scene(shutter=1, interpolate_rotation=True, seed=7) -> Scene:
    """Observations that hit every branch of the functor:
    frame 0 all-zero poses (small-angle Rodrigues branch), tau clamped below 0 and above 1,
    points behind the camera / at z ~ 1e-8 (functor returns false), large rotations near pi,
    full distortion model (k1 k2 p1 p2 k3 all non-zero)."""
    rng = np.random.default_rng(seed)
    F, P = 6, 40
    cam = np.array([860.0, 870.0, 1e-2, -3e-3, 4e-4, -2e-4, 1e-3, 640.0, 360.0])
    poses = np.zeros((F, 12))
    poses[1, :3] = [1e-9, -2e-9, 1e-9]            # theta^2 below DBL_EPSILON
    poses[1, 6:9] = [2e-9, 1e-9, -1e-9]
    poses[1, 3:6] = [0.1, 0.0, 0.0]
    poses[1, 9:12] = [0.12, 0.01, 0.0]
    poses[2, :3] = [3.0, 0.5, -0.4]               # near pi
    poses[2, 6:9] = [3.05, 0.45, -0.38]
    poses[2, 3:6] = [0.0, 0.2, -0.5]
    poses[2, 9:12] = [0.05, 0.2, -0.45]
    for f in range(3, F):
        poses[f, :3] = rng.normal(0, 0.3, 3)
        poses[f, 6:9] = poses[f, :3] + rng.normal(0, 0.02, 3)
        poses[f, 3:6] = rng.normal(0, 0.5, 3)
        poses[f, 9:12] = poses[f, 3:6] + rng.normal(0, 0.05, 3)
    points = np.stack([rng.uniform(-2, 2, P), rng.uniform(-1.5, 1.5, P), rng.uniform(3, 9, P)], axis=1)
    points[0] = [0.1, 0.1, -3.0]                  # behind frames 0/1
    points[1] = [0.3, -0.2, 1e-9]                 # z just below the 1e-8 gate for frame 0
    points[2] = [0.3, -0.2, 2e-8]                 # just above it
    fr, pt = np.meshgrid(np.arange(F), np.arange(P), indexing="ij")
    fr, pt = fr.reshape(-1).astype(np.int32), pt.reshape(-1).astype(np.int32)
    n = fr.size
    xy = np.stack([rng.uniform(-200, 1500, n), rng.uniform(0, 720, n)], axis=1)  # tau <0 and >1 occur
    xy[:8, 0] = [-5.0, 0.0, 1280.0, 1290.0, 640.0, 1e-3, 1279.999, 320.0]
    const = np.zeros(F, dtype=bool)
    const[0] = True
    return Scene(cam=cam, shutter=shutter, scanlines=np.array([0, 1280], dtype=np.int32),
                 interpolate_rotation=interpolate_rotation, poses=poses, points=points, obs_xy=xy,
                 obs_frame=fr, obs_point=pt, const_frames=const, name="edge")


def shuffled(scene: Scene, seed=3) -> Scene:
    """Same observations in a random (not frame-sorted) order."""
    perm = np.random.default_rng(seed).permutation(scene.num_obs)
    return Scene(cam=scene.cam, shutter=scene.shutter, scanlines=scene.scanlines,
                 interpolate_rotation=scene.interpolate_rotation, poses=scene.poses, points=scene.points,
                 obs_xy=scene.obs_xy[perm].copy(), obs_frame=scene.obs_frame[perm].copy(),
                 obs_point=scene.obs_point[perm].copy(), const_frames=scene.const_frames, name=scene.name + "-shuffled")


def damped_normal_equation_residual(scene, r, J, delta_poses, delta_points, radius, min_diag=1e-6, max_diag=1e32):
    """Size-independent property of one LM step (works at millions of observations, no sparse matrices): with the
    Jacobi scaling s = 1 / (1 + |column|) and D^2 = clamp(diag(J'^T J')) / radius of the Ceres-1.9-style step (J' = J s),
    y = delta / s must satisfy (J'^T J' + D^2) y = -J'^T r on the free parameters.  Returns
      * the relative residual |(J'^T J' + D^2) y + J'^T r| / |J'^T r|,
      * the model cost change -m.(r + m/2), m = J' y,
      * the largest COMPONENT-wise residual, each component relative to the magnitude of the terms summed into it
        (sum_n |J'_ni| (|m_n| + |r_n|) + |D^2 y|_i, no cancellation) -- a localised error (one pose component, one
        point) is not diluted by the millions of other parameters there,
    all computed here from (r, J) alone."""
    F, P = scene.num_frames, scene.num_points
    fr, pt = np.asarray(scene.obs_frame), np.asarray(scene.obs_point)
    # per-block row-major (Ceres' contract): J_pose0[2][6] | J_pose1[2][6] | J_point[2][3]
    Jc = np.concatenate([J[:, :12].reshape(-1, 2, 6), J[:, 12:24].reshape(-1, 2, 6)], axis=2)     # [N, 2, 12]
    Jp = J[:, 24:].reshape(-1, 2, 3)
    free_c = np.repeat(~np.asarray(scene.const_frames, dtype=bool), 12).reshape(F, 12)

    def jt(v):                                   # J^T v for v [N, 2]  ->  ([F, 12], [P, 3])
        tc = np.einsum("nrk,nr->nk", Jc, v)
        tp = np.einsum("nrk,nr->nk", Jp, v)
        gc = np.stack([np.bincount(fr, weights=tc[:, k], minlength=F) for k in range(12)], axis=1)
        gp = np.stack([np.bincount(pt, weights=tp[:, k], minlength=P) for k in range(3)], axis=1)
        return gc * free_c, gp

    col_c = np.stack([np.bincount(fr, weights=(Jc[:, :, k] ** 2).sum(axis=1), minlength=F) for k in range(12)], axis=1)
    col_p = np.stack([np.bincount(pt, weights=(Jp[:, :, k] ** 2).sum(axis=1), minlength=P) for k in range(3)], axis=1)
    col_c = col_c * free_c
    sc_, sp_ = 1.0 / (1.0 + np.sqrt(col_c)), 1.0 / (1.0 + np.sqrt(col_p))
    d2c = np.clip(col_c * sc_ ** 2, min_diag, max_diag) / radius
    d2p = np.clip(col_p * sp_ ** 2, min_diag, max_diag) / radius
    dc, dp = np.asarray(delta_poses)[:F] * free_c, np.asarray(delta_points)
    m = np.einsum("nrk,nk->nr", Jc, dc[fr]) + np.einsum("nrk,nk->nr", Jp, dp[pt])          # J delta = J' y
    hc, hp = jt(m)
    gc, gp = jt(r)
    # scaled space: s J^T (J delta) + D^2 (delta / s) + s J^T r
    res_c = (sc_ * hc + d2c * dc / sc_ + sc_ * gc) * free_c
    res_p = sp_ * hp + d2p * dp / sp_ + sp_ * gp
    g_norm = np.sqrt(np.sum((sc_ * gc) ** 2) + np.sum((sp_ * gp) ** 2))
    rel = np.sqrt(np.sum(res_c ** 2) + np.sum(res_p ** 2)) / g_norm
    mcc = -float(np.sum(m * (r + 0.5 * m)))
    mag = np.abs(m) + np.abs(r)
    ac = np.stack([np.bincount(fr, weights=np.einsum("nr,nr->n", np.abs(Jc[:, :, k]), mag), minlength=F) for k in range(12)], axis=1)
    ap = np.stack([np.bincount(pt, weights=np.einsum("nr,nr->n", np.abs(Jp[:, :, k]), mag), minlength=P) for k in range(3)], axis=1)
    den_c = sc_ * ac + np.abs(d2c * dc / sc_)
    den_p = sp_ * ap + np.abs(d2p * dp / sp_)
    comp = max(float(np.max(np.abs(res_c) / np.where(den_c > 0, den_c, 1.0), initial=0.0)),
               float(np.max(np.abs(res_p) / np.where(den_p > 0, den_p, 1.0), initial=0.0)))
    return float(rel), mcc, comp
