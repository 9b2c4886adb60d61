"""CPU suite: a third, independent pin of the oracle's residuals and Jacobians.

The oracle's Jacobian comes from forward-mode `Jet<15>` arithmetic in a stand-in for Ceres (oracle/shim); it is checked
against the reference's golden vectors and central finite differences in test_oracle_cpu.py.  Here the functor is restated
once more in torch.float64 and differentiated by REVERSE-mode autograd -- no code shared with the shim, the port or the
CUDA kernel:
  interpolate_rs  cam.h:316-349  (tau from observed x, clamped; GLOBAL copies pose0)
  interpolate     cam.h:294-311  (component-wise lerp; the rotation only if useSlerp)
  w2c             cam.h:355-366  (ceres::AngleAxisRotatePoint: Rodrigues, first-order branch for tiny angles)
  w2i / c2i       cam.h:372-419  (z < 1e-8 fails; dehomogenise, distort, focal length, principal point)
  distort         cam.h:49-72
  residual        video_bundler_free.h:45-65 (projection - observation)
"""
import numpy as np
import pytest
import torch

from conftest import rel_block_err
from helpers import edge_scene, small_scene


def _rotate(aa, p):
    """ceres::AngleAxisRotatePoint (third-party; call site cam.h:365)."""
    theta2 = (aa * aa).sum(-1, keepdim=True)
    big = theta2 > np.finfo(np.float64).eps
    t2 = torch.where(big, theta2, torch.ones_like(theta2))        # keeps sqrt's gradient finite in the other branch
    theta = torch.sqrt(t2)
    w = aa / theta
    c, s = torch.cos(theta), torch.sin(theta)
    wxp = torch.linalg.cross(w, p)
    wdp = (w * p).sum(-1, keepdim=True)
    rodrigues = p * c + wxp * s + w * wdp * (1.0 - c)
    small = p + torch.linalg.cross(aa, p)
    return torch.where(big, rodrigues, small)


def _residuals(sc, p0, p1, X, xy, cam_rows=None):
    cam = torch.tensor(np.asarray(sc.cam, dtype=np.float64)) if cam_rows is None else cam_rows.T.unsqueeze(-1)
    x_obs = xy[:, 0:1]
    if int(sc.shutter) == 0:                                       # GLOBAL
        pose = p0
    else:
        s0, s1 = float(sc.scanlines[0]), float(sc.scanlines[1])
        tau = ((x_obs - s0) / (s1 - s0)).clamp(0.0, 1.0)           # obs = {x, x}: x for both shutter directions
        centre = p0[:, 3:] + (p1[:, 3:] - p0[:, 3:]) * tau
        rot = p0[:, :3] + (p1[:, :3] - p0[:, :3]) * tau if sc.interpolate_rotation else p0[:, :3]
        pose = torch.cat([rot, centre], dim=1)
    pt = _rotate(pose[:, :3], X - pose[:, 3:])
    z = pt[:, 2:3]
    valid = (z >= 1e-8).squeeze(1)
    zs = torch.where(z >= 1e-8, z, torch.ones_like(z))
    u = pt[:, :2] / zs
    xp, yp = u[:, 0:1], u[:, 1:2]
    fx, fy, k1, k2, p1_, p2_, k3, cx, cy = [cam[i] for i in range(9)]
    r2 = xp * xp + yp * yp
    d = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3))
    xy_ = xp * yp
    px = d * xp + (2.0 * p1_ * xy_ + p2_ * (r2 + 2.0 * xp * xp))
    py = d * yp + (p1_ * (r2 + 2.0 * yp * yp) + 2.0 * p2_ * xy_)
    proj = torch.cat([px * fx + cx, py * fy + cy], dim=1)
    return proj - xy, valid


def _autograd_eval_cam(sc):
    """d residual / d (fx fy k1 k2 p1 p2 k3 cx cy): the 9-block of the uncalibrated variant <2; 9, 6, 6, 3>
    (VideoSfmBaRs.h:38-49, ReprojectionError::operator()(camera, pose, point, residuals) video_bundler_free.h:33-41)."""
    fr = torch.as_tensor(np.asarray(sc.obs_frame, dtype=np.int64))
    pi = torch.as_tensor(np.asarray(sc.obs_point, dtype=np.int64))
    poses = torch.tensor(np.asarray(sc.poses, dtype=np.float64))
    points = torch.tensor(np.asarray(sc.points, dtype=np.float64))
    xy = torch.tensor(np.asarray(sc.obs_xy, dtype=np.float64))
    n = xy.shape[0]
    cam = torch.tensor(np.asarray(sc.cam, dtype=np.float64)).repeat(n, 1).requires_grad_(True)   # one copy per observation
    res, valid = _residuals(sc, poses[fr, :6], poses[fr, 6:], points[pi], xy, cam_rows=cam)
    Jc = np.zeros((n, 18))
    for row in range(2):
        (g,) = torch.autograd.grad(res[:, row].sum(), (cam,), retain_graph=True)
        Jc[:, 9 * row:9 * row + 9] = g.numpy()
    return Jc, valid.numpy()


def _autograd_eval(sc):
    fr = torch.as_tensor(np.asarray(sc.obs_frame, dtype=np.int64))
    pi = torch.as_tensor(np.asarray(sc.obs_point, dtype=np.int64))
    poses = torch.tensor(np.asarray(sc.poses, dtype=np.float64))
    points = torch.tensor(np.asarray(sc.points, dtype=np.float64))
    xy = torch.tensor(np.asarray(sc.obs_xy, dtype=np.float64))
    p0 = poses[fr, :6].clone().requires_grad_(True)                # one copy per observation: rows are independent
    p1 = poses[fr, 6:].clone().requires_grad_(True)
    X = points[pi].clone().requires_grad_(True)
    res, valid = _residuals(sc, p0, p1, X, xy)
    n = res.shape[0]
    J = np.zeros((n, 30))
    for row in range(2):
        g0, g1, gx = torch.autograd.grad(res[:, row].sum(), (p0, p1, X), retain_graph=True, allow_unused=True)
        J[:, 6 * row:6 * row + 6] = g0.numpy()                     # J_pose0[2][6] | J_pose1[2][6] | J_point[2][3]
        if g1 is not None:                                         # (a global shutter never reads pose1)
            J[:, 12 + 6 * row:12 + 6 * row + 6] = g1.numpy()
        J[:, 24 + 3 * row:24 + 3 * row + 3] = gx.numpy()
    return res.detach().numpy(), J, valid.numpy()


SCENES = [("C1", lambda: small_scene())] + [
    (f"edge_shutter{s}_rot{int(r)}", (lambda s=s, r=r: edge_scene(s, r))) for s in (0, 1, 2) for r in (False, True)]


@pytest.mark.parametrize("name,make", SCENES, ids=[n for n, _ in SCENES])
def test_oracle_matches_reverse_mode_autograd(oracle_built, name, make):
    sc = make()
    res_t, J_t, valid_t = _autograd_eval(sc)
    impls = ["port"] + (["ref"] if oracle_built.ref_available() else [])
    for impl in impls:
        res, J, valid = oracle_built.evaluate(sc, impl=impl)
        res, J = np.asarray(res).reshape(-1, 2), np.asarray(J).reshape(-1, 30)
        ok = np.asarray(valid).reshape(-1) == 1
        assert np.array_equal(ok, valid_t), f"{impl}: the z >= 1e-8 gate disagrees"
        assert ok.sum() > 0.5 * ok.size or name.startswith("edge")
        # relative 1e-6 is the north star's bar; two FP64 evaluations of the same formulas agree far better
        assert np.max(np.abs(res[ok] - res_t[ok]) / np.maximum(1.0, np.abs(res_t[ok]))) <= 1e-9, impl
        assert rel_block_err(J[ok], J_t[ok]).max() <= 1e-9, impl


def test_pose1_block_vanishes_for_a_global_shutter(oracle_built):
    sc = edge_scene(0, True)
    _, J_t, valid_t = _autograd_eval(sc)
    assert not J_t[valid_t, 12:24].any()
    _, J, valid = oracle_built.evaluate(sc, impl="port")
    J = np.asarray(J).reshape(-1, 30)
    assert not J[np.asarray(valid).reshape(-1) == 1, 12:24].any()


@pytest.mark.parametrize("name,make", SCENES, ids=[n for n, _ in SCENES])
def test_intrinsics_block_matches_reverse_mode_autograd(oracle_built, name, make):
    sc = make()
    Jc_t, valid_t = _autograd_eval_cam(sc)
    Jc = np.asarray(oracle_built.intrinsics_jacobian(sc)).reshape(-1, 18)           # closed form, travels everywhere
    assert rel_block_err(Jc[valid_t], Jc_t[valid_t]).max() <= 1e-9
    assert not Jc[~valid_t].any()
    if oracle_built.ref_available():                                                # the reference functor under Jet<24>
        _, _, Jc_ref, valid = oracle_built.evaluate_cam_ref(sc)
        ok = np.asarray(valid).reshape(-1) == 1
        assert np.array_equal(ok, valid_t)
        assert rel_block_err(np.asarray(Jc_ref).reshape(-1, 18)[ok], Jc_t[ok]).max() <= 1e-9


@pytest.mark.parametrize("a", [0.5, 2.0])
def test_huber_correction_is_the_gradient_of_the_robustified_cost(oracle_built, a):
    """ceres::HuberLoss on every residual block (CeresHandler.h:85-90).  Ceres' Corrector is third-party and restated in
    oracle.apply_huber (rho'' <= 0: residual and Jacobian times sqrt(rho')); what can be pinned without Ceres is that the
    corrected quantities are consistent with the DEFINITION of the robustified problem: cost = 1/2 sum rho(|r_i|^2) with
    rho(s) = s (s <= a^2), 2 a sqrt(s) - a^2 otherwise, and J_w^T r_w = its exact gradient (reverse-mode autograd)."""
    from rsba_b200.scene import Scene
    sc = small_scene()
    xy = sc.obs_xy.copy()
    xy[::9] += 25.0                                   # gross outliers: both branches of rho are exercised
    sc = Scene(**{**sc.__dict__, "obs_xy": xy})
    r0, J0, v0 = oracle_built.evaluate(sc, impl="port")
    r0, J0 = np.asarray(r0).reshape(-1, 2), np.asarray(J0).reshape(-1, 30)
    ok = np.asarray(v0).reshape(-1) == 1
    rw, Jw, cost_w = oracle_built.apply_huber(r0, J0, a)
    s_np = np.sum(r0 * r0, axis=1)
    assert (s_np[ok] > a * a).sum() > 100 and (s_np[ok] <= a * a).sum() > 100

    fr = torch.as_tensor(np.asarray(sc.obs_frame, dtype=np.int64))
    pi = torch.as_tensor(np.asarray(sc.obs_point, dtype=np.int64))
    poses = torch.tensor(np.asarray(sc.poses, dtype=np.float64), requires_grad=True)
    points = torch.tensor(np.asarray(sc.points, dtype=np.float64), requires_grad=True)
    res, valid = _residuals(sc, poses[fr, :6], poses[fr, 6:], points[pi], torch.tensor(xy))
    assert np.array_equal(valid.numpy(), ok)
    sq = (res * res).sum(1)
    rho = torch.where(sq > a * a, 2.0 * a * torch.sqrt(torch.where(sq > a * a, sq, torch.ones_like(sq))) - a * a, sq)
    cost = 0.5 * rho[valid].sum()
    g_poses, g_points = torch.autograd.grad(cost, (poses, points))
    assert abs(float(cost.detach()) - cost_w) <= 1e-12 * cost_w

    # J_w^T r_w scattered to the parameter blocks (J = J_pose0[2][6] | J_pose1[2][6] | J_point[2][3])
    gp = np.zeros_like(g_poses.numpy())
    gx = np.zeros_like(g_points.numpy())
    f_np, p_np = np.asarray(sc.obs_frame), np.asarray(sc.obs_point)
    for row in range(2):
        np.add.at(gp[:, :6], f_np[ok], Jw[ok, 6 * row:6 * row + 6] * rw[ok, row:row + 1])
        np.add.at(gp[:, 6:], f_np[ok], Jw[ok, 12 + 6 * row:12 + 6 * row + 6] * rw[ok, row:row + 1])
        np.add.at(gx, p_np[ok], Jw[ok, 24 + 3 * row:24 + 3 * row + 3] * rw[ok, row:row + 1])
    scale = max(np.abs(g_poses.numpy()).max(), np.abs(g_points.numpy()).max())
    assert np.abs(gp - g_poses.numpy()).max() <= 1e-9 * scale
    assert np.abs(gx - g_points.numpy()).max() <= 1e-9 * scale
