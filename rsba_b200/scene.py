"""Synthetic rolling-shutter scenes.

The reference ships no scene generator (only the switches ``tracks.synthetic``
``SfmOptions.h:40`` and the noise knobs ``SfmOptions.h:15-18``), so the benchmark scene is
ours and is fully defined here: counter-based RNG (Philox) with a fixed seed, the camera
model of ``mat/cam.h`` and the session layout of ``CeresHandler::Add``
(``CeresHandler.h:208-255``: observations inserted frame-major, every observation tied to
``(frame.poses[0], frame.poses[1], track.pt)``).

This module is host-side workload construction.  It does not import the oracle and it is
not the checker; the projection below is only used to synthesise observations.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SEED = 20240917
IMAGE_W, IMAGE_H = 1280, 720
# fx fy k1 k2 p1 p2 k3 cx cy  (mat/cam.h:23-34); focal/k1 as in test/mat_test.cc:213
DEFAULT_CAM = np.array([860.0, 860.0, 1e-3, 0.0, 0.0, 0.0, 0.0, 640.0, 360.0])
SHUTTER_GLOBAL, SHUTTER_HORIZONTAL, SHUTTER_VERTICAL = 0, 1, 2  # mat/cam.h:37-41

# BASELINE.json configs: name -> (frames, points, observations per point)
CONFIGS = {
    "C1": (10, 500, 10),
    "C2": (100, 20_000, 25),
    "C3": (1_000, 200_000, 25),
    "C5": (4_000, 1_000_000, 20),
    # not in BASELINE.json: the sizes of C3 with FULL covisibility (every frame pair shares points) -- an orbit
    # around one object, the loop-closure / turntable case in which the reduced camera matrix is dense
    "C3dense": (1_000, 200_000, 25),
    "C2dense": (100, 20_000, 25),
}
ORBIT = {"C3dense", "C2dense"}


@dataclass
class Scene:
    cam: np.ndarray                 # [9]
    shutter: int
    scanlines: np.ndarray           # [2] int32
    interpolate_rotation: bool
    poses: np.ndarray               # [F, 12] initial estimate: pose0[6] | pose1[6]
    points: np.ndarray              # [P, 3] initial estimate
    obs_xy: np.ndarray              # [N, 2]
    obs_frame: np.ndarray           # [N] int32, non-decreasing (frame-major insertion order)
    obs_point: np.ndarray           # [N] int32
    const_frames: np.ndarray        # [F] bool -- SetParameterBlockConstant on both pose blocks
    poses_true: np.ndarray | None = None
    points_true: np.ndarray | None = None
    name: str = ""
    meta: dict = field(default_factory=dict)

    @property
    def num_frames(self) -> int:
        return int(self.poses.shape[0])

    @property
    def num_points(self) -> int:
        return int(self.points.shape[0])

    @property
    def num_obs(self) -> int:
        return int(self.obs_xy.shape[0])

    def subscene(self, num_frames: int) -> "Scene":
        """First ``num_frames`` frames with the points they alone observe renumbered; a
        bounded sample of the same workload (used for the CPU baseline)."""
        keep = self.obs_frame < num_frames
        pts, inv = np.unique(self.obs_point[keep], return_inverse=True)
        return Scene(
            cam=self.cam.copy(), shutter=self.shutter, scanlines=self.scanlines.copy(),
            interpolate_rotation=self.interpolate_rotation,
            poses=self.poses[:num_frames].copy(), points=self.points[pts].copy(),
            obs_xy=self.obs_xy[keep].copy(), obs_frame=self.obs_frame[keep].copy(),
            obs_point=inv.astype(np.int32), const_frames=self.const_frames[:num_frames].copy(),
            poses_true=None if self.poses_true is None else self.poses_true[:num_frames].copy(),
            points_true=None if self.points_true is None else self.points_true[pts].copy(),
            name=f"{self.name}[:{num_frames}]", meta=dict(self.meta))


# --------------------------------------------------------------------------------------
# camera model, vectorised (value only) -- used to synthesise observations
# --------------------------------------------------------------------------------------
def rotate(aa: np.ndarray, pt: np.ndarray) -> np.ndarray:
    """Angle-axis rotation of ``pt`` [M,3] by ``aa`` [M,3] (Rodrigues, small-angle branch)."""
    theta2 = np.einsum("ij,ij->i", aa, aa)
    big = theta2 > np.finfo(np.float64).eps
    theta = np.sqrt(np.where(big, theta2, 1.0))
    w = aa / theta[:, None]
    c, s = np.cos(theta), np.sin(theta)
    wxp = np.cross(w, pt)
    tmp = np.einsum("ij,ij->i", w, pt) * (1.0 - c)
    full = pt * c[:, None] + wxp * s[:, None] + w * tmp[:, None]
    small = pt + np.cross(aa, pt)
    return np.where(big[:, None], full, small)


def project(cam: np.ndarray, pose: np.ndarray, X: np.ndarray):
    """``w2i`` (mat/cam.h:401-419) on arrays: returns (proj [M,2], z [M])."""
    pt = rotate(pose[:, :3], X - pose[:, 3:6])
    z = pt[:, 2]
    zs = np.where(np.abs(z) < 1e-300, 1e-300, z)
    xp, yp = pt[:, 0] / zs, pt[:, 1] / zs
    k1, k2, p1, p2, k3 = cam[2], cam[3], cam[4], cam[5], cam[6]
    r2 = xp * xp + yp * yp
    d = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3))
    xy = xp * yp
    px = d * xp + (2.0 * p1 * xy + p2 * (r2 + 2.0 * xp * xp))
    py = d * yp + (p1 * (r2 + 2.0 * yp * yp) + 2.0 * p2 * xy)
    return np.stack([px * cam[0] + cam[7], py * cam[1] + cam[8]], axis=1), z


def interpolate_rs(pose01: np.ndarray, x: np.ndarray, shutter: int, scanlines, interp_rot: bool):
    """``interpolate_rs`` (mat/cam.h:316-349) on arrays; tau from ``x`` (F2), clamped."""
    p0, p1 = pose01[:, :6], pose01[:, 6:]
    if shutter == SHUTTER_GLOBAL:
        return p0.copy()
    tau = np.clip((x - scanlines[0]) / float(scanlines[1] - scanlines[0]), 0.0, 1.0)[:, None]
    out = p0 + (p1 - p0) * tau
    if not interp_rot:
        out[:, :3] = p0[:, :3]
    return out


def rs_project(cam, pose01, X, shutter, scanlines, interp_rot, iters=12):
    """Exact RS projection by fixed point on the scan-line time, like ``reproject``
    (struct/VideoSfM.cc:139-155)."""
    proj = np.tile(np.array([cam[7], cam[8]]), (X.shape[0], 1))
    z = None
    for _ in range(iters):
        pose = interpolate_rs(pose01, proj[:, 0], shutter, scanlines, interp_rot)
        new, z = project(cam, pose, X)
        done = np.max(np.sum((new - proj) ** 2, axis=1)) <= 1e-12
        proj = new
        if done:
            break
    return proj, z


# --------------------------------------------------------------------------------------
def make_scene(num_frames: int, num_points: int, obs_per_point: int, seed: int = SEED,
               name: str = "", noise_px: float = 0.5, interpolate_rotation: bool = True,
               shutter: int = SHUTTER_HORIZONTAL) -> Scene:
    """Deterministic synthetic RS scene with exactly num_points*obs_per_point observations."""
    F, P, K = int(num_frames), int(num_points), int(obs_per_point)
    if K > F:
        raise ValueError("obs_per_point cannot exceed the number of frames")
    rng = np.random.Generator(np.random.Philox(seed))
    cam = DEFAULT_CAM.copy()
    scan = np.array([0, IMAGE_W], dtype=np.int32)

    # ---- trajectory: forward motion + sinusoidal sway, smooth small rotations
    t = np.arange(F + 1, dtype=np.float64)
    ph = rng.uniform(0, 2 * np.pi, size=6)
    centre = np.stack([0.30 * np.sin(2 * np.pi * t / 180.0 + ph[0]),
                       0.15 * np.sin(2 * np.pi * t / 110.0 + ph[1]),
                       0.05 * t], axis=1)
    rot = 0.05 * np.stack([np.sin(2 * np.pi * t / 140.0 + ph[2]),
                           np.sin(2 * np.pi * t / 200.0 + ph[3]),
                           np.sin(2 * np.pi * t / 90.0 + ph[4])], axis=1)
    centre -= centre[0]
    rot -= rot[0]                      # frame 0 starts at the origin (CeresHandler.h:132-140)
    pose0 = np.concatenate([rot, centre], axis=1)          # [F+1, 6]
    pose1 = pose0[:-1] + 0.5 * (pose0[1:] - pose0[:-1])    # intra-frame motion
    poses_true = np.concatenate([pose0[:-1], pose1], axis=1)
    poses_true[0] = 0.0                                    # frame 0: both control poses zero

    # ---- points: 4..12 m ahead of a home frame, seen by the K frames nearest to it
    home = rng.integers(0, F, size=P)
    first = np.clip(home - K // 2, 0, F - K)               # window [first, first+K)
    points_true = np.empty((P, 3))
    todo = np.arange(P)
    margin = 8.0
    for _ in range(200):
        if todo.size == 0:
            break
        m = todo.size
        depth = rng.uniform(4.0, 12.0, size=m)
        u = rng.uniform(-0.62, 0.62, size=m) * (IMAGE_W / 2) / cam[0]
        v = rng.uniform(-0.62, 0.62, size=m) * (IMAGE_H / 2) / cam[1]
        hp = poses_true[home[todo], :6]
        local = np.stack([u * depth, v * depth, depth], axis=1)
        X = rotate(-hp[:, :3], local) + hp[:, 3:6]         # c2w (mat/cam.h:117-126)
        ok = np.ones(m, dtype=bool)
        for k in range(K):
            pj, z = rs_project(cam, poses_true[first[todo] + k], X, shutter, scan,
                               interpolate_rotation, iters=4)
            ok &= (z > 0.5) & (pj[:, 0] > margin) & (pj[:, 0] < IMAGE_W - margin) \
                & (pj[:, 1] > margin) & (pj[:, 1] < IMAGE_H - margin)
        points_true[todo[ok]] = X[ok]
        todo = todo[~ok]
    if todo.size:
        raise RuntimeError("scene generator failed to place all points")

    # ---- observations, frame-major then point-major inside a frame
    obs_point = np.repeat(np.arange(P, dtype=np.int64), K)
    obs_frame = (first[:, None] + np.arange(K)[None, :]).reshape(-1)
    order = np.lexsort((obs_point, obs_frame))
    obs_point, obs_frame = obs_point[order], obs_frame[order]
    proj, z = rs_project(cam, poses_true[obs_frame], points_true[obs_point], shutter, scan,
                         interpolate_rotation)
    assert np.all(z > 0.25)
    obs_xy = proj + rng.normal(0.0, noise_px, size=proj.shape)

    # ---- initial estimate: truth + noise; frame 0 is the gauge and stays put
    poses = poses_true.copy()
    for blk in (0, 6):                 # N(0,1e-3) rad on rotations, N(0,1e-2 m) on centres
        poses[:, blk:blk + 3] += rng.normal(0, 1e-3, size=(F, 3))
        poses[:, blk + 3:blk + 6] += rng.normal(0, 1e-2, size=(F, 3))
    poses[0] = 0.0
    points = points_true + rng.normal(0, 2e-2, size=(P, 3))
    const_frames = np.zeros(F, dtype=bool)
    const_frames[0] = True

    return Scene(cam=cam, shutter=shutter, scanlines=scan, interpolate_rotation=interpolate_rotation,
                 poses=np.ascontiguousarray(poses), points=np.ascontiguousarray(points),
                 obs_xy=np.ascontiguousarray(obs_xy), obs_frame=obs_frame.astype(np.int32),
                 obs_point=obs_point.astype(np.int32), const_frames=const_frames,
                 poses_true=poses_true, points_true=points_true, name=name,
                 meta={"seed": seed, "frames": F, "points": P, "obs_per_point": K,
                       "noise_px": noise_px, "image": [IMAGE_W, IMAGE_H]})


def make_orbit_scene(num_frames: int, num_points: int, obs_per_point: int, seed: int = SEED, name: str = "",
                     noise_px: float = 0.5) -> Scene:
    """Fully covisible scene: the camera orbits an object of radius 1.2 m at 6 m distance, every point is seen
    from ``obs_per_point`` frames drawn at random over the WHOLE sequence, so any two frames share points
    (expected P (K/F)^2 of them) and the reduced camera matrix is dense.  Same camera model, noise and
    frame-0 gauge as make_scene."""
    F, P, K = int(num_frames), int(num_points), int(obs_per_point)
    rng = np.random.Generator(np.random.Philox(seed + 1))
    cam = DEFAULT_CAM.copy()
    scan = np.array([0, IMAGE_W], dtype=np.int32)
    shutter, interp = SHUTTER_HORIZONTAL, True
    R = 6.0
    theta = 2.0 * np.pi * np.arange(F + 1, dtype=np.float64) / F        # unwrapped: pose1 blends towards the next frame
    centre = np.stack([R * np.sin(theta), 0.02 * np.sin(7 * theta), R - R * np.cos(theta)], axis=1)   # frame 0 at the origin
    rot = np.stack([np.zeros(F + 1), theta, np.zeros(F + 1)], axis=1)   # looks at the object centre (0, 0, R)
    pose0 = np.concatenate([rot, centre], axis=1)
    pose1 = pose0[:-1] + 0.5 * (pose0[1:] - pose0[:-1])
    poses_true = np.concatenate([pose0[:-1], pose1], axis=1)
    poses_true[0] = 0.0
    # points: uniform in a ball around the object centre
    d = rng.normal(size=(P, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    points_true = np.array([0.0, 0.0, R]) + 1.2 * d * rng.uniform(0, 1, size=(P, 1)) ** (1 / 3)
    # K distinct frames per point, anywhere in the sequence (argpartition of random keys = sampling w/o replacement)
    obs_frame = np.empty((P, K), dtype=np.int64)
    for b in range(0, P, 20000):
        e = min(P, b + 20000)
        obs_frame[b:e] = np.argpartition(rng.random((e - b, F)), K - 1, axis=1)[:, :K]
    obs_point = np.repeat(np.arange(P, dtype=np.int64), K)
    obs_frame = obs_frame.reshape(-1)
    order = np.lexsort((obs_point, obs_frame))
    obs_point, obs_frame = obs_point[order], obs_frame[order]
    proj, z = rs_project(cam, poses_true[obs_frame], points_true[obs_point], shutter, scan, interp)
    assert np.all(z > 0.25) and np.all((proj[:, 0] > 0) & (proj[:, 0] < IMAGE_W) & (proj[:, 1] > 0) & (proj[:, 1] < IMAGE_H))
    obs_xy = proj + rng.normal(0.0, noise_px, size=proj.shape)
    poses = poses_true.copy()
    for blk in (0, 6):
        poses[:, blk:blk + 3] += rng.normal(0, 1e-3, size=(F, 3))
        poses[:, blk + 3:blk + 6] += rng.normal(0, 1e-2, size=(F, 3))
    poses[0] = 0.0
    points = points_true + rng.normal(0, 2e-2, size=(P, 3))
    const_frames = np.zeros(F, dtype=bool)
    const_frames[0] = True
    return Scene(cam=cam, shutter=shutter, scanlines=scan, interpolate_rotation=interp,
                 poses=np.ascontiguousarray(poses), points=np.ascontiguousarray(points),
                 obs_xy=np.ascontiguousarray(obs_xy), obs_frame=obs_frame.astype(np.int32),
                 obs_point=obs_point.astype(np.int32), const_frames=const_frames,
                 poses_true=poses_true, points_true=points_true, name=name,
                 meta={"seed": seed, "frames": F, "points": P, "obs_per_point": K, "noise_px": noise_px,
                       "image": [IMAGE_W, IMAGE_H], "pattern": "orbit"})


def make_config(name: str, cache: bool = True, **kw) -> Scene:
    """BASELINE.json config by name.  Large scenes are cached as .npz under
    $RSBA_SCENE_CACHE (default /tmp/rsba_scene_cache): generation is deterministic."""
    import os
    F, P, K = CONFIGS[name]
    gen = make_orbit_scene if name in ORBIT else make_scene
    if kw or not cache or F * P < 1_000_000:
        return gen(F, P, K, name=name, **kw)
    d = os.environ.get("RSBA_SCENE_CACHE", "/tmp/rsba_scene_cache")
    path = os.path.join(d, f"{name}_seed{SEED}.npz")
    if os.path.exists(path):
        try:
            g = np.load(path)
            return Scene(cam=g["cam"], shutter=int(g["shutter"]), scanlines=g["scanlines"],
                         interpolate_rotation=bool(g["interpolate_rotation"]), poses=g["poses"],
                         points=g["points"], obs_xy=g["obs_xy"], obs_frame=g["obs_frame"],
                         obs_point=g["obs_point"], const_frames=g["const_frames"],
                         poses_true=g["poses_true"], points_true=g["points_true"], name=name,
                         meta={"seed": SEED, "frames": F, "points": P, "obs_per_point": K})
        except Exception:
            pass
    sc = gen(F, P, K, name=name)
    try:
        os.makedirs(d, exist_ok=True)
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, cam=sc.cam, shutter=np.int32(sc.shutter), scanlines=sc.scanlines,
                 interpolate_rotation=np.int32(sc.interpolate_rotation), poses=sc.poses, points=sc.points,
                 obs_xy=sc.obs_xy, obs_frame=sc.obs_frame, obs_point=sc.obs_point,
                 const_frames=sc.const_frames, poses_true=sc.poses_true, points_true=sc.points_true)
        os.replace(tmp, path)
    except OSError:
        pass
    return sc
