"""CPU suite: include/rsba_cuda_handler.hpp instantiated with the REFERENCE'S OWN types -- the Thrift-generated
``vision::sfm::gen::Session`` (src/rsba/gen-cpp/sfm_types.{h,cpp}, compiled in place against the inert Thrift
stand-in tests/shim/thrift) and ``vision::SfmOptions`` (src/rsba/SfmOptions.h) -- with a recording problem type in
place of the GPU one: which ceres::Problem calls does Add() make (CeresHandler.h:92-392)?  Covered: the residual
loop in insertion order, fixFirstNCameras, the match fallback for observations without a track (:223-241),
revalidateReprojections (:244-248), pose initialisation by velocity extrapolation (:99-144), windowed BA
(:288-300), SubsetParameterization (:362-371), motion + pose priors incl. a frame that carries priors only.

Runs only where /root/reference exists (this container); the decisions are checked against the oracle's
restatement of validate() (oracle.validate_sweep), which is pinned to the reference's primitives."""
import os
import subprocess

import numpy as np
import pytest

from rsba_b200.scene import Scene, make_scene
from test_gpu_handler import write_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
BIN = os.path.join(ROOT, "tests", "tools", "handler_ref_check")

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "rsba", "gen-cpp", "sfm_types.cpp")),
                                reason="the reference tree is not here (GPU box): nothing to compile against")


@pytest.fixture(scope="module")
def tool():
    src = os.path.join(ROOT, "tests", "tools", "handler_ref_check.cc")
    deps = [src] + [os.path.join(ROOT, "include", h) for h in
                    ("rsba_cuda_handler.hpp", "rsba_cuda_session.hpp", "rsba_cuda_functors.hpp", "rsba_reproj_math.h",
                     "rsba_cuda.h")]
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(p) for p in deps):
        subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "tests", "shim"), "-I", os.path.join(ROOT, "oracle", "shim"), "-I", REF,
                        src, os.path.join(REF, "rsba", "gen-cpp", "sfm_types.cpp"), "-o", BIN,
                        "-L", os.path.join(ROOT, "rsba_b200", "lib"), "-lrsba_cuda",
                        "-Wl,-rpath," + os.path.join(ROOT, "rsba_b200", "lib")], check=True)
    return BIN


def scene_with_outliers():
    sc = make_scene(8, 120, 6, name="handler-ref")
    xy = sc.obs_xy.copy()
    bad = np.arange(sc.num_obs) % 7 == 3
    xy[bad] += 60.0                                    # far outside sqrt(sqrdThreshold) = 4 px
    return Scene(**{**sc.__dict__, "obs_xy": xy}), bad


def run(tool, tmp_path, sc, mode):
    src = str(tmp_path / f"{mode}.bin")
    write_scene(src, sc)
    r = subprocess.run([tool, src, mode], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return [ln.split() for ln in r.stdout.strip().split("\n")]


def frame_major(sc):
    """observations in the order CeresHandler::Add visits them: frame by frame, insertion order inside a frame"""
    return np.argsort(sc.obs_frame, kind="stable")


def test_plain_residual_loop_and_first_camera(tool, tmp_path):
    sc = make_scene(8, 120, 6, name="handler-ref")
    log = run(tool, tmp_path, sc, "plain")
    assert not [ln for ln in log if ln[0] == "exception"]
    rs = [ln for ln in log if ln[0] == "rs"]
    order = frame_major(sc)
    assert len(rs) == sc.num_obs
    for ln, i in zip(rs, order):
        f, p = int(sc.obs_frame[i]), int(sc.obs_point[i])
        assert ln[1:4] == [f"f{f}p0", f"f{f}p1", f"x{p}"]
        assert float(ln[4]) == sc.obs_xy[i, 0] and float(ln[5]) == sc.obs_xy[i, 1]
    consts = [ln[1] for ln in log if ln[0] == "const"]
    assert consts == ["f0p0", "f0p1"]                  # fixFirstNCameras = 1 (CeresHandler.h:342-346)
    assert log[1][0] == "camera" and log[1][3:] == ["shutter", str(sc.shutter), "scan", str(sc.scanlines[0]),
                                                    str(sc.scanlines[1]), "interp", "1"]


def test_revalidate_reprojections_skips_what_validate_rejects(tool, tmp_path, oracle_built):
    sc, bad = scene_with_outliers()
    ok, _ = oracle_built.validate_sweep(sc)            # the reference's validate(), restated on the port
    assert (ok == 0).sum() >= bad.sum() > 5
    log = run(tool, tmp_path, sc, "revalidate")
    rs = [ln for ln in log if ln[0] == "rs"]
    keep = [i for i in frame_major(sc) if ok[i]]
    assert len(rs) == len(keep)
    for ln, i in zip(rs, keep):
        assert ln[3] == f"x{int(sc.obs_point[i])}" and float(ln[4]) == sc.obs_xy[i, 0]


def test_match_fallback_finds_the_track_of_a_matched_observation(tool, tmp_path, oracle_built):
    """CeresHandler.h:223-241 with useOnlyValidMatches = false: an observation without a track takes the track of
    a matched observation iff that track's point validates against THIS observation."""
    sc, _ = scene_with_outliers()
    ok, _ = oracle_built.validate_sweep(sc)
    # the driver's edit, restated: per point, the observation list in frame-major insertion order
    order = frame_major(sc)
    pos_in_frame = np.zeros(sc.num_obs, dtype=int)
    count = {}
    for i in order:
        f = int(sc.obs_frame[i])
        pos_in_frame[i] = count.get(f, 0)
        count[f] = pos_in_frame[i] + 1
    per_point = {}
    for i in order:
        per_point.setdefault(int(sc.obs_point[i]), []).append(i)
    has_track = np.ones(sc.num_obs, bool)
    match = {}
    for p, lst in per_point.items():
        for a, i in enumerate(lst):
            if (int(sc.obs_frame[i]) + pos_in_frame[i]) % 3 == 0 and len(lst) >= 2:
                has_track[i] = False
                match[i] = lst[(a + 1) % len(lst)]
    expect = []
    for i in order:
        if has_track[i]:
            expect.append(i)
        elif has_track[match[i]] and ok[i]:           # same point: validate(point of the matched track, this obs)
            expect.append(i)
    assert 0 < (~has_track).sum() and len(expect) < sc.num_obs
    log = run(tool, tmp_path, sc, "matches")
    rs = [ln for ln in log if ln[0] == "rs"]
    assert len(rs) == len(expect)
    for ln, i in zip(rs, expect):
        assert ln[3] == f"x{int(sc.obs_point[i])}" and float(ln[4]) == sc.obs_xy[i, 0]


def test_pose_initialisation_extrapolates_the_velocity(tool, tmp_path):
    sc = make_scene(8, 120, 6, name="handler-ref")
    log = run(tool, tmp_path, sc, "init")
    init = [ln for ln in log if ln[0] == "initialised"][0]
    assert init[1:3] == ["1", "2"]
    got = np.array([float(v) for v in init[3:]]).reshape(2, 6)
    F = sc.num_frames
    want = sc.poses[F - 2].reshape(2, 6) + (sc.poses[F - 2].reshape(2, 6) - sc.poses[F - 3].reshape(2, 6))
    assert np.array_equal(got, want)                   # minus6 + plus6 (CeresHandler.h:110-113)
    # the new poses are the blocks of that frame's residuals
    last = [ln for ln in log if ln[0] == "rs" and ln[1] == f"f{F - 1}p0"]
    assert len(last) == int((sc.obs_frame == F - 1).sum()) and all(ln[2] == f"f{F - 1}p1" for ln in last)


def test_windowed_ba_freezes_old_tracks(tool, tmp_path):
    sc = make_scene(8, 120, 6, name="handler-ref")
    log = run(tool, tmp_path, sc, "window")
    start = sc.num_frames // 2
    adds = [int(ln[1]) for ln in log if ln[0] == "add"]
    assert adds == list(range(start, sc.num_frames))
    old = set(int(p) for p in sc.obs_point[sc.obs_frame < start])
    in_window = set(int(p) for p in sc.obs_point[sc.obs_frame >= start])
    consts = set(ln[1] for ln in log if ln[0] == "const" and ln[1].startswith("x"))
    assert consts == {f"x{p}" for p in old & in_window}


def test_fix_rotation_uses_subset_parameterization(tool, tmp_path):
    sc = make_scene(8, 120, 6, name="handler-ref")
    log = run(tool, tmp_path, sc, "rotation")
    subs = [ln for ln in log if ln[0] == "subset"]
    assert [ln[1] for ln in subs] == [f"f{k}p{i}" for k in range(sc.num_frames) for i in (0, 1)]
    assert all(ln[2:] == ["0", "1", "2"] for ln in subs)
    assert not [ln for ln in log if ln[0] == "const"]


def test_priors_and_a_frame_that_carries_priors_only(tool, tmp_path):
    sc = make_scene(8, 120, 6, name="handler-ref")
    log = run(tool, tmp_path, sc, "priors")
    assert not [ln for ln in log if ln[0] == "exception"]
    F = sc.num_frames
    motion = [ln for ln in log if ln[0] == "motion"]
    assert [ln[4:] for ln in motion] == [[f"f{k}p0", f"f{k}p1", f"f{k - 1}p0", f"f{k - 1}p1"] for k in range(1, F)]
    assert all(ln[1] == "2" and float(ln[2]) == 3.0 and float(ln[3]) == 1.0 for ln in motion)   # acceleration prior
    assert [ln for ln in log if ln[0] == "ratio_block"]               # interFrameRatio == 1: a free block (:156-180)
    pp = [ln for ln in log if ln[0] == "poseprior"]
    assert [ln[3:] for ln in pp] == [[f"q{k}p{i}", f"f{k}p{i}"] for k in range(1, F) for i in (0, 1)]
    # the last frame has no observation: its two control poses are still registered as a frame
    tail = log[[i for i, ln in enumerate(log) if ln[0] == "add"][-1]:]
    assert ["frame", f"f{F - 1}p0", f"f{F - 1}p1"] in tail and not [ln for ln in tail if ln[0] == "rs"]
    # frame 0 is the fixed camera; the prior of frame 1 also fixes it through its previous-frame blocks (:182-186)
    assert [ln[1] for ln in log if ln[0] == "const"].count("f0p0") >= 1


def test_session_soa_gathers_in_add_order_and_scatters_back(tool, tmp_path):
    """rsba_cuda_session.hpp's SessionSoA on the reference's own gen::Session: the flat arrays hold every
    observation with a usable track, frame by frame in insertion order (the order CeresHandler::Add feeds ceres,
    CeresHandler.h:208), fixFirstNCameras becomes the frame mask, and scatter() writes poses / points back."""
    sc = make_scene(6, 40, 4, name="soa")
    out = run(tool, tmp_path, sc, "soa")
    head = out[0]
    order = frame_major(sc)
    keep = [i for i in order if sc.obs_point[i] != 1]          # track 1 was made invalid
    first = order[0]                                              # frame 0, observation 0 lost its track
    keep = [i for i in keep if i != first]
    assert head == ["soa", str(sc.num_frames), str(len(set(sc.obs_point[keep]))), str(len(keep))]
    obs = [ln for ln in out if ln[0] == "obs"]
    in_frame = np.zeros(sc.num_obs, int)
    for f in range(sc.num_frames):
        idx = order[sc.obs_frame[order] == f]
        in_frame[idx] = np.arange(idx.size)
    assert len(obs) == len(keep)
    for ln, i in zip(obs, keep):
        assert [int(ln[1]), int(ln[2]), int(ln[3])] == [sc.obs_frame[i], in_frame[i], sc.obs_point[i]]
        assert float(ln[4]) == sc.obs_xy[i, 0] and float(ln[5]) == sc.obs_xy[i, 1]
    masks = {int(ln[1]): int(ln[2]) for ln in out if ln[0] == "mask"}
    assert masks[0] == 0xFFF and all(masks[k] == 0 for k in range(1, sc.num_frames))
    for ln in out:
        if ln[0] == "pose":
            k = int(ln[1])
            assert float(ln[2]) == sc.poses[k, 0] + 0.5 and float(ln[3]) == sc.poses[k, 11] + 0.5
        if ln[0] == "pt":
            q = int(ln[1])
            moved = q in set(sc.obs_point[keep])
            assert float(ln[2]) == sc.points[q, 2] - (0.25 if moved else 0.0)


def test_eval_tracks_bookkeeping_follows_the_reference(tool, tmp_path):
    """applyEvalTracks against a transcription of VideoSfMHandler::evalTracks (VideoSfMHandler.cc:381-405) on the
    reference's gen::Session: an observation with a track whose predicate equals `drop_when` (TRUE in the reference
    as published, :390) loses its track flag, its reference leaves the track's list, and a valid track that falls
    below minReprojections turns invalid; both polarities are driven (even frames: the reference's, odd: the other)."""
    sc = make_scene(6, 40, 4, name="evaltracks")
    out = run(tool, tmp_path, sc, "evaltracks")
    order = frame_major(sc)
    F = sc.num_frames
    frames = [[] for _ in range(F)]           # per frame: [track, has_track]
    tracks = {p: {"valid": True, "obs": []} for p in range(sc.num_points)}
    for i in order:
        f = int(sc.obs_frame[i])
        tracks[int(sc.obs_point[i])]["obs"].append((f, len(frames[f])))
        frames[f].append([int(sc.obs_point[i]), True])
    want_counts = []
    for k in range(F):
        n_obs = n_tr = 0
        for oi, (tr, has) in enumerate(frames[k]):
            if not has:
                continue
            ok = (k + oi) % 3 == 0
            if ok != (k % 2 == 0):
                continue
            frames[k][oi][1] = False
            n_obs += 1
            t = tracks[tr]
            if (k, oi) in t["obs"]:
                t["obs"].remove((k, oi))
                if t["valid"] and len(t["obs"]) < 4:
                    t["valid"] = False
                    n_tr += 1
        want_counts.append((n_obs, n_tr))
    got_counts = [(int(ln[2]), int(ln[3])) for ln in out if ln[0] == "frame"]
    assert got_counts == want_counts
    for ln in out:
        if ln[0] == "o":
            assert bool(int(ln[3])) == frames[int(ln[1])][int(ln[2])][1]
        if ln[0] == "t":
            t = tracks[int(ln[1])]
            assert bool(int(ln[2])) == t["valid"]
            assert [tuple(int(v) for v in r.split(":")) for r in ln[3:]] == t["obs"]
