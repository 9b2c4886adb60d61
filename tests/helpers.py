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
