"""The C++ host side (include/rsba_cuda_handler.hpp): a CeresHandler-shaped Add()/solve() driven
exactly as VideoSfMHandler::BA drives the reference (VideoSfMHandler.cc:586-592), through the
pointer-identity API, must reproduce the bulk-API solve."""
import os
import subprocess

import numpy as np
import pytest

from rsba_b200.scene import Scene, make_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "tools", "handler_check")


def write_scene(path, sc):
    with open(path, "wb") as f:
        np.array([sc.num_frames, sc.num_points, sc.num_obs, sc.shutter], dtype=np.int64).tofile(f)
        np.asarray(sc.cam, dtype=np.float64).tofile(f)
        np.asarray(sc.scanlines, dtype=np.int32).tofile(f)
        np.ascontiguousarray(sc.poses, dtype=np.float64).tofile(f)
        np.ascontiguousarray(sc.points, dtype=np.float64).tofile(f)
        np.ascontiguousarray(sc.obs_xy, dtype=np.float64).tofile(f)
        np.ascontiguousarray(sc.obs_frame, dtype=np.int32).tofile(f)
        np.ascontiguousarray(sc.obs_point, dtype=np.int32).tofile(f)


def test_handler_binary_is_built():
    """build() compiles the header against plain structs with g++ (no GPU needed for that)."""
    assert os.path.exists(BIN), "run __graft_entry__.build()"


@pytest.mark.gpu
def test_handler_add_solve_matches_bulk_api(tmp_path):
    import rsba_b200.api as api
    sc = make_scene(12, 400, 8, name="handler")
    src, dst = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(src, sc)
    r = subprocess.run([BIN, src, dst, "1", "6"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(dst)
    usable, iters, c0, c1 = out[:4]
    poses = out[4:4 + 12 * sc.num_frames].reshape(-1, 12)
    points = out[4 + 12 * sc.num_frames:].reshape(-1, 3)
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        s = pb.solve(api.default_options(max_num_iterations=6))
        po, pt = pb.get_parameters()
    assert usable == 1 and int(iters) == s.iterations
    assert abs(c0 - s.initial_cost) <= 1e-12 * s.initial_cost
    assert abs(c1 - s.final_cost) <= 1e-9 * s.final_cost
    # points that no frame observes are never handed to the handler and keep their values
    seen = np.zeros(sc.num_points, bool)
    seen[sc.obs_point] = True
    assert np.linalg.norm(poses - po) <= 1e-7 * np.linalg.norm(po)
    assert np.linalg.norm(points[seen] - pt[seen]) <= 1e-7 * np.linalg.norm(pt[seen])
    assert not poses[0].any()                      # fixFirstNCameras = 1: frame 0 untouched


@pytest.mark.gpu
def test_handler_windowed_ba_freezes_old_tracks(tmp_path):
    """startFrame > 0 (windowedBA, VideoSfMHandler.cc:185): frames before the window are not added,
    tracks seen before it are SetParameterBlockConstant (CeresHandler.h:288-300)."""
    sc = make_scene(24, 800, 8, name="window")
    src, dst = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(src, sc)
    start = 12
    r = subprocess.run([BIN, src, dst, "0", "5", str(start)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(dst)
    assert out[0] == 1 and out[3] < out[2]
    poses = out[4:4 + 12 * sc.num_frames].reshape(-1, 12)
    points = out[4 + 12 * sc.num_frames:].reshape(-1, 3)
    assert np.array_equal(poses[:start], sc.poses[:start])          # outside the window
    old = np.zeros(sc.num_points, bool)
    old[sc.obs_point[sc.obs_frame < start]] = True
    assert np.array_equal(points[old], sc.points[old])               # frozen tracks
    assert (~old).sum() > 20 and np.any(points[~old] != sc.points[~old])


@pytest.mark.gpu
def test_handler_global_shutter_single_pose_frames(tmp_path):
    """A GLOBAL-shutter session with one pose per frame (ReprojectionError <2; 6, 3>, CeresHandler.h:265-286):
    the handler completes each frame with a constant stand-in block; result == bulk API with the
    second control poses masked constant."""
    import rsba_b200.api as api
    sc = make_scene(10, 300, 6, name="gs-handler", shutter=0)
    src, dst = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(src, sc)
    r = subprocess.run([BIN, src, dst, "1", "8", "0", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(dst)
    poses = out[4:4 + 12 * sc.num_frames].reshape(-1, 12)
    points = out[4 + 12 * sc.num_frames:].reshape(-1, 3)
    mask = np.full(sc.num_frames, 0xFC0, dtype=np.uint16)            # second control pose constant everywhere
    mask[0] = 0xFFF
    with api.Problem(0) as pb:
        pb.set_camera(sc.cam, 0, sc.scanlines, True)
        pb.set_scene(sc.obs_xy, sc.obs_frame, sc.obs_point, sc.num_frames, sc.num_points, mask)
        pb.set_parameters(sc.poses, sc.points)
        s = pb.solve(api.default_options(max_num_iterations=8))
        po, pt = pb.get_parameters()
    assert out[0] == 1 and abs(out[3] - s.final_cost) <= 1e-9 * s.final_cost
    seen = np.zeros(sc.num_points, bool)
    seen[sc.obs_point] = True
    assert np.linalg.norm(poses[:, :6] - po[:, :6]) <= 1e-7 * np.linalg.norm(po[:, :6])
    assert np.linalg.norm(points[seen] - pt[seen]) <= 1e-7 * np.linalg.norm(pt[seen])


@pytest.mark.gpu
def test_handler_uncalibrated_session(tmp_path):
    """opt.model.calibrated = false: CeresHandler::Add uses CreateWithCam with sess.cam as a parameter block
    (CeresHandler.h:256-264); the handler's result == bulk API with free intrinsics, sess.cam updated in place."""
    import rsba_b200.api as api
    sc = make_scene(12, 400, 8, name="uncal-handler")
    src, dst = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(src, sc)
    r = subprocess.run([BIN, src, dst, "1", "8", "0", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(dst)
    nf, npnt = sc.num_frames, sc.num_points
    poses = out[4:4 + 12 * nf].reshape(-1, 12)
    cam = out[4 + 12 * nf + 3 * npnt:]
    assert cam.size == 9
    with api.Problem(0) as pb:
        pb.set_intrinsics_free(True)
        pb.load_scene(sc)
        s = pb.solve(api.default_options(max_num_iterations=8))
        po, _ = pb.get_parameters()
        cam_bulk = pb.get_camera()
    assert out[0] == 1 and abs(out[3] - s.final_cost) <= 1e-9 * s.final_cost
    assert np.linalg.norm(poses - po) <= 1e-7 * np.linalg.norm(po)
    assert np.linalg.norm(cam - cam_bulk) <= 1e-9 * np.linalg.norm(cam_bulk) and np.abs(cam - sc.cam).max() > 0


@pytest.mark.gpu
def test_handler_velocity_priors_with_the_default_free_ratio(tmp_path):
    """opt.ceres.constFrameVelocity != 0 with the default interFrameRatio == 1: the handler's own copy of the
    ratio is a variable, lower-bounded block shared by every prior (CeresHandler.h:156-180) and is updated in
    place; result == bulk API with rsba_cuda_set_inter_frame_ratio_free."""
    import rsba_b200.api as api
    sc = make_scene(12, 400, 8, name="velo-handler")
    src, dst = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(src, sc)
    r = subprocess.run([BIN, src, dst, "1", "8", "0", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(dst)
    nf, npnt = sc.num_frames, sc.num_points
    poses = out[4:4 + 12 * nf].reshape(-1, 12)
    ratio = out[4 + 12 * nf + 3 * npnt:]
    assert ratio.size == 1
    mask = np.zeros(nf, dtype=np.uint16)
    mask[0] = 0xFFF                                          # fixFirstNCameras = 1; frame 1's prior fixes frame 0 too
    with api.Problem(0) as pb:
        pb.set_camera(sc.cam, sc.shutter, sc.scanlines, sc.interpolate_rotation)
        pb.set_scene(sc.obs_xy, sc.obs_frame, sc.obs_point, nf, npnt, mask)
        pb.set_parameters(sc.poses, sc.points)
        pb.set_motion_priors([1] * (nf - 1), [10.0] * (nf - 1), [1.0] * (nf - 1), list(range(1, nf)), list(range(nf - 1)))
        pb.set_inter_frame_ratio_free(True, 1.0)
        s = pb.solve(api.default_options(max_num_iterations=8))
        po, _ = pb.get_parameters()
        ratio_bulk = pb.inter_frame_ratio()
    assert out[0] == 1 and abs(out[3] - s.final_cost) <= 1e-9 * s.final_cost
    assert np.linalg.norm(poses - po) <= 1e-7 * np.linalg.norm(po)
    assert abs(ratio[0] - ratio_bulk) <= 1e-9 and ratio[0] != 1.0


@pytest.mark.gpu
def test_handler_good_pose_priors(tmp_path):
    """opt.ceres.trustPriorCamRotation / trustPriorCamPosition with f.priorPoses (CeresHandler.h:188-204): one
    GoodPosePrior per control pose of every frame >= fixFirstNCameras; the prior blocks are free parameter
    blocks (the reference never fixes them) and are written back in place; result == bulk API."""
    import rsba_b200.api as api
    sc = make_scene(10, 300, 8, name="good-handler")
    src, dst = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(src, sc)
    r = subprocess.run([BIN, src, dst, "1", "8", "0", "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(dst)
    nf, npnt = sc.num_frames, sc.num_points
    poses = out[4:4 + 12 * nf].reshape(-1, 12)
    blocks = out[4 + 12 * nf + 3 * npnt:].reshape(nf, 2, 6)
    k, c = np.arange(nf)[:, None, None], np.arange(6)[None, None, :]
    prior0 = sc.poses.reshape(nf, 2, 6) + np.where(c < 3, 1e-3, 2e-2) * ((k + c) % 3 - 1)
    frames = [f for f in range(1, nf) for _ in (0, 1)]
    which = [w for _ in range(1, nf) for w in (0, 1)]
    mask = np.zeros(nf, dtype=np.uint16)
    mask[0] = 0xFFF
    with api.Problem(0) as pb:
        pb.set_camera(sc.cam, sc.shutter, sc.scanlines, sc.interpolate_rotation)
        pb.set_scene(sc.obs_xy, sc.obs_frame, sc.obs_point, nf, npnt, mask)
        pb.set_parameters(sc.poses, sc.points)
        pb.set_pose_priors(frames, which, [20.0] * len(frames), [6.0] * len(frames), prior0[1:].reshape(-1, 6))
        s = pb.solve(api.default_options(max_num_iterations=8))
        po, _ = pb.get_parameters()
        val, _ = pb.pose_priors()
    assert out[0] == 1 and abs(out[3] - s.final_cost) <= 1e-9 * s.final_cost
    assert np.linalg.norm(poses - po) <= 1e-7 * np.linalg.norm(po)
    assert np.array_equal(blocks[0], prior0[0])                      # frame 0 < fixFirstNCameras: no prior, untouched
    assert np.linalg.norm(blocks[1:].reshape(-1, 6) - val) <= 1e-7 * np.linalg.norm(val)
    assert np.abs(blocks[1:] - prior0[1:]).max() > 1e-6              # the free prior blocks moved


@pytest.mark.gpu
def test_handler_frame_with_pose_priors_only(tmp_path):
    """A frame that carries GoodPosePrior blocks and NO observation (nor a motion prior) is a legal Ceres problem
    (CeresHandler.h:188-204 adds the priors before the residual loop): its control poses are registered through
    rsba_cuda_add_frame_blocks and the solve moves them half-way to the (free) prior blocks -- both ends of a
    GoodPosePrior are parameter blocks, so pose and prior meet in the middle of the weighted residual."""
    sc = make_scene(10, 300, 8, name="good-handler")
    src, dst = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(src, sc)
    r = subprocess.run([BIN, src, dst, "1", "8", "0", "5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(dst)
    nf, npnt = sc.num_frames, sc.num_points
    assert out[0] == 1 and out[3] < out[2]
    poses = out[4:4 + 12 * nf].reshape(nf, 2, 6)
    blocks = out[4 + 12 * nf + 3 * npnt:].reshape(nf, 2, 6)
    # last frame: no observation -> the only residual on its poses is the prior, which the solve drives to zero
    assert np.abs(poses[-1] - blocks[-1]).max() <= 1e-6
    assert np.abs(poses[-1] - sc.poses[-1].reshape(2, 6)).max() > 1e-5      # ... by moving both ends
    assert not poses[0].any()


@pytest.mark.gpu
def test_session_sweeps_on_the_device_match_the_host_predicate(tmp_path):
    """include/rsba_cuda_session.hpp: validateFrame (one rsba_cuda_validate sweep per frame) against the host
    validate() of rsba_cuda_handler.hpp on every observation -- a third of them pushed 60 px off --, evalTracks'
    bookkeeping on top of it, and reprojectPoints landing on the untouched observations."""
    sc = make_scene(10, 300, 8, name="session")
    # at the TRUE parameters only the 0.5 px observation noise separates projection and observation
    sc = Scene(**{**sc.__dict__, "poses": sc.poses_true, "points": sc.points_true})
    src, dst = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(src, sc)
    r = subprocess.run([BIN, src, dst, "1", "1", "0", "6"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("session")][0].split()
    got = {line[i]: int(line[i + 1]) for i in range(1, len(line), 2)}
    assert got["checked"] == sc.num_obs and got["mismatches"] == 0
    # the pushed observations fail the predicate and (polarity: drop what FAILS) lose their track
    order = np.argsort(sc.obs_frame, kind="stable")
    pushed = 0
    for f in range(sc.num_frames):
        n = int((sc.obs_frame == f).sum())
        pushed += sum(1 for oi in range(n) if (f + oi) % 3 == 0)
    assert got["dropped"] >= pushed and got["kept"] == sc.num_obs - got["dropped"]
    assert got["dropped"] <= pushed + sc.num_obs // 50          # + a few noisy ones beyond 4 px
    assert got["reproj_bad"] <= sc.num_obs // 50
    assert order.size == sc.num_obs


@pytest.mark.gpu
def test_session_soa_bulk_path_equals_the_pointer_path(tmp_path):
    """SessionSoA gather -> upload -> rsba_cuda_solve -> download -> scatter gives what Handler::Add + solve gives."""
    sc = make_scene(12, 400, 8, name="soa-gpu")
    src = str(tmp_path / "scene.bin")
    write_scene(src, sc)
    outs = []
    for mode in ("0", "7"):
        dst = str(tmp_path / f"out{mode}.bin")
        r = subprocess.run([BIN, src, dst, "1", "6", "0", mode], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(np.fromfile(dst))
    a, b = outs
    assert a[0] == 1 and b[0] == 1 and a[1] == b[1]
    assert abs(a[3] - b[3]) <= 1e-9 * a[3]
    assert np.linalg.norm(a[4:] - b[4:]) <= 1e-7 * np.linalg.norm(a[4:])


@pytest.mark.gpu
def test_handler_over_every_gpu_of_the_box_from_one_thread(tmp_path):
    """Handler<..., MultiGpuProblem> (rsba_cuda_create_multi underneath) with all visible devices -- one on the
    single-GPU test box, where it must equal the plain handler bit for bit; N > 1 elsewhere, to 1e-9."""
    sc = make_scene(16, 600, 8, name="multi-handler")
    src = str(tmp_path / "scene.bin")
    write_scene(src, sc)
    outs = []
    for mode in ("0", "8"):
        dst = str(tmp_path / f"out{mode}.bin")
        r = subprocess.run([BIN, src, dst, "1", "6", "0", mode], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        outs.append(np.fromfile(dst))
        if mode == "8":
            assert any(ln.startswith("gpus ") for ln in r.stdout.splitlines())
    a, b = outs
    assert a[1] == b[1] and abs(a[3] - b[3]) <= 1e-9 * a[3]
    assert np.linalg.norm(a[4:] - b[4:]) <= 1e-9 * np.linalg.norm(a[4:])
