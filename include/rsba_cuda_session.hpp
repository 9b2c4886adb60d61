// rsba_cuda_session.hpp -- header-only C++11: the session <-> SoA marshaller and the track bookkeeping that
// surround a bundle adjustment in henrique/rsba's driver (paths under src/rsba/), over the C ABI of rsba_cuda.h.
//
//   SessionSoA<Session>      gathers a window of frames of a Thrift-shaped sfm::Session (frames[k].poses,
//                            frames[k].obs[i].{x, y, track}, getTrack(id).pt) into the flat arrays of
//                            rsba_cuda_set_scene / rsba_cuda_set_parameters and scatters the optimised poses and
//                            points back -- the bulk twin of the pointer-identity path of rsba_cuda_handler.hpp
//                            (CeresHandler.h:208-302 feeds the same blocks to ceres::Problem one residual at a time).
//   evalTracks(...)          VideoSfMHandler::evalTracks (VideoSfMHandler.cc:377-410): the validate() predicate of
//                            struct/VideoSfM.cc:159-169 for every tracked observation of a frame as ONE device sweep
//                            (rsba_cuda_validate) instead of a host loop, then the reference's bookkeeping on the
//                            session: o.__isset.track, the track's observation list, Track::valid against
//                            opt.tracks.minReprojections.
//   applyEvalTracks(...)     that bookkeeping alone, from a vector of predicate values (host only; what the CPU
//                            tests drive with the reference's own Thrift types).
//   reprojectPoints(...)     reproject(sess, f, opt, pt, obs) of struct/VideoSfM.cc:139-155 for many points of a
//                            frame in one call (rsba_cuda_reproject).
// Template parameters: any types with the members named above -- tests/tools/session_ref_check.cc instantiates them
// with the reference's gen::Session / SfmOptions, tests/tools/handler_check.cc with plain look-alikes.
// There is no CPU fallback for the sweeps: without a CUDA device the calls fail with RSBA_ERR_NO_DEVICE.
#ifndef RSBA_CUDA_SESSION_HPP_
#define RSBA_CUDA_SESSION_HPP_

#include <cstddef>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "rsba_cuda.h"

namespace rsba_cuda {

inline void session_check(int rc) {
  if (rc != RSBA_OK) throw std::runtime_error(std::string("rsba_cuda: ") + rsba_cuda_last_error());
}

// ------------------------------------------------------------------ Session <-> SoA
template <typename Session>
struct SessionSoA {
  std::size_t start = 0, end = 0;            // frame window [start, end]
  std::vector<double> obs_xy;                // [N][2]
  std::vector<int> obs_frame, obs_point;     // [N] frame = frameKey - start, point = index into points
  std::vector<int> obs_key;                  // [N] index of the observation inside its frame
  std::vector<double> poses;                 // [F][12] pose0 | pose1 (a single-pose frame: pose1 = pose0)
  std::vector<double> points;                // [P][3]
  std::vector<int> point_track;              // [P] track key of point p
  std::vector<unsigned short> pose_mask;     // [F] constant-scalar bits (0xFFF = both control poses fixed)
  std::vector<unsigned char> point_const;    // [P]

  int num_frames() const { return (int)(end - start + 1); }
  int num_points() const { return (int)point_track.size(); }
  long num_obs() const { return (long)obs_frame.size(); }

  // Every observation of frames [s, e] whose track carries a 3-D point (and is valid when only_valid is set: the
  // rule of CeresHandler.h:217-220).  Frames < fix_first are constant (CeresHandler.h:342-346); a track with an
  // observation in a frame before window_start is constant (windowed BA, :288-300).
  void gather(Session& sess, std::size_t s, std::size_t e, bool only_valid, std::size_t fix_first = 0,
              std::size_t window_start = 0) {
    start = s; end = e;
    obs_xy.clear(); obs_frame.clear(); obs_point.clear(); obs_key.clear(); points.clear(); point_track.clear();
    point_const.clear();
    const int F = num_frames();
    poses.assign((std::size_t)12 * F, 0.0);
    pose_mask.assign(F, 0);
    std::map<int, int> point_of_track;
    for (std::size_t fk = s; fk <= e; ++fk) {
      auto& f = sess.frames[fk];
      if (f.poses.empty()) throw std::runtime_error("rsba_cuda: empty frame");
      if (f.poses.size() > 2) throw std::runtime_error("rsba_cuda: frames with one pose per scan line (fullDoF) are not supported");
      for (int k = 0; k < 6; ++k) {
        poses[12 * (fk - s) + k] = f.poses[0][k];
        poses[12 * (fk - s) + 6 + k] = f.poses[f.poses.size() == 2 ? 1 : 0][k];
      }
      if (fk < fix_first) pose_mask[fk - s] = 0xFFF;
      else if (f.poses.size() == 1) pose_mask[fk - s] = 0xFC0;     // the stand-in second pose never moves
      for (std::size_t oi = 0; oi < f.obs.size(); ++oi) {
        auto& o = f.obs[oi];
        if (!o.__isset.track) continue;
        auto& t = sess.getTrack(o.track);
        if (!t.__isset.pt || (only_valid && !t.valid)) continue;
        auto it = point_of_track.find(o.track);
        if (it == point_of_track.end()) {
          it = point_of_track.emplace(o.track, (int)point_track.size()).first;
          point_track.push_back(o.track);
          for (int k = 0; k < 3; ++k) points.push_back(t.pt[k]);
          bool old = false;
          if (window_start > 0)
            for (auto& ref : t.obs)
              if ((std::size_t)ref.frame < window_start) { old = true; break; }
          point_const.push_back(old ? 1 : 0);
        }
        obs_xy.push_back(o.x); obs_xy.push_back(o.y);
        obs_frame.push_back((int)(fk - s));
        obs_point.push_back(it->second);
        obs_key.push_back((int)oi);
      }
    }
  }

  // camera + scene + parameters into a problem handle (rsba_cuda_set_camera / set_scene / set_parameters)
  template <typename Options>
  void upload(rsba_problem* h, const Session& sess, const Options& opt) const {
    const int scan[2] = {(int)sess.scanlines[0], (int)sess.scanlines[1]};
    session_check(rsba_cuda_set_camera(h, sess.cam.data(), (int)sess.rs, scan, opt.model.interpolateRotation ? 1 : 0));
    session_check(rsba_cuda_set_scene(h, num_obs(), obs_xy.data(), obs_frame.data(), obs_point.data(), num_frames(),
                                      num_points(), pose_mask.data(), point_const.data()));
    session_check(rsba_cuda_set_parameters(h, poses.data(), points.data()));
  }

  // optimised parameters back into the session's blocks
  void download(rsba_problem* h) { session_check(rsba_cuda_get_parameters(h, poses.data(), points.data())); }
  void scatter(Session& sess) const {
    for (std::size_t fk = start; fk <= end; ++fk) {
      auto& f = sess.frames[fk];
      for (std::size_t pi = 0; pi < f.poses.size(); ++pi)
        for (int k = 0; k < 6; ++k) f.poses[pi][k] = poses[12 * (fk - start) + 6 * pi + k];
    }
    for (std::size_t p = 0; p < point_track.size(); ++p) {
      auto& t = sess.getTrack(point_track[p]);
      for (int k = 0; k < 3; ++k) t.pt[k] = points[3 * p + k];
    }
  }
};

// ------------------------------------------------------------------ evalTracks
struct EvalTracksCount {
  unsigned observations = 0, tracks = 0;   // "bad reprojections and bad tracks removed" (VideoSfMHandler.cc:407-409)
};

// The bookkeeping of VideoSfMHandler::evalTracks (VideoSfMHandler.cc:381-405) for frame `frameKey`.  ok[oi] is
// validate(sess, f, opt, t.pt, obs) of observation oi (ignored where the observation has no track).
// drop_when: the predicate value on which the reference drops the observation.  As published, :390 reads
// `if (validate(...)) { o.__isset.track = false; ... }`, i.e. TRUE -- that is the default here, bit for bit what
// the reference does; a caller who wants the evidently intended clean-up of FAILED re-projections passes false.
template <typename Session, typename Options>
EvalTracksCount applyEvalTracks(Session& sess, std::size_t frameKey, const std::vector<unsigned char>& ok,
                                const Options& opt, bool drop_when = true) {
  EvalTracksCount n;
  auto& f = sess.frames[frameKey];
  if (ok.size() < f.obs.size()) throw std::runtime_error("rsba_cuda: evalTracks needs one predicate value per observation");
  for (std::size_t oi = 0; oi < f.obs.size(); ++oi) {
    auto& o = f.obs[oi];
    if (!o.__isset.track) continue;
    auto& t = sess.getTrack(o.track);
    if ((ok[oi] != 0) != drop_when) continue;
    o.__isset.track = false;
    n.observations++;
    for (std::size_t i = 0; i < t.obs.size(); ++i) {
      if (t.obs[i].frame == (int)frameKey && t.obs[i].obs == (int)oi) {
        t.obs.erase(t.obs.begin() + i);
        if (t.valid && t.obs.size() < (std::size_t)opt.tracks.minReprojections) {
          t.valid = false;
          n.tracks++;
        }
        break;
      }
    }
  }
  return n;
}

// validate() of every tracked observation of the frame as one device sweep; ok has one entry per observation of
// the frame (0 where the observation has no track).  `h` is any problem handle of the target device: its scene is
// replaced by this frame's observations.
template <typename Session, typename Options>
std::vector<unsigned char> validateFrame(rsba_problem* h, Session& sess, std::size_t frameKey, const Options& opt) {
  auto& f = sess.frames[frameKey];
  std::vector<unsigned char> ok(f.obs.size(), 0);
  if (f.__isset.cam) throw std::runtime_error("rsba_cuda: per-frame intrinsics (f.cam) are not on the device path");
  if (f.poses.empty()) throw std::runtime_error("empty frame");                 // struct/VideoSfM.cc:104
  if (f.poses.size() > 2) throw std::runtime_error("rsba_cuda: frames with one pose per scan line (fullDoF) are not supported");
  std::vector<double> xy, pts;
  std::vector<int> fr, pt, key;
  for (std::size_t oi = 0; oi < f.obs.size(); ++oi) {
    auto& o = f.obs[oi];
    if (!o.__isset.track) continue;
    auto& t = sess.getTrack(o.track);
    xy.push_back(o.x); xy.push_back(o.y);
    fr.push_back(0);
    pt.push_back((int)key.size());
    for (int k = 0; k < 3; ++k) pts.push_back(t.pt[k]);
    key.push_back((int)oi);
  }
  if (key.empty()) return ok;
  double poses[12];
  for (int k = 0; k < 6; ++k) {
    poses[k] = f.poses[0][k];
    poses[6 + k] = f.poses[f.poses.size() == 2 ? 1 : 0][k];
  }
  const bool rs = f.poses.size() == 2;     // getPose: one pose = that pose, whatever the session's shutter (:104-106)
  const int scan[2] = {(int)sess.scanlines[0], (int)sess.scanlines[1]};
  session_check(rsba_cuda_set_camera(h, sess.cam.data(), rs ? (int)sess.rs : 0, scan, opt.model.interpolateRotation ? 1 : 0));
  session_check(rsba_cuda_set_scene(h, (long)key.size(), xy.data(), fr.data(), pt.data(), 1, (int)key.size(), nullptr, nullptr));
  session_check(rsba_cuda_set_parameters(h, poses, pts.data()));
  std::vector<unsigned char> got(key.size(), 0);
  session_check(rsba_cuda_validate(h, (double)opt.tracks.sqrdThreshold, (double)opt.tracks.minDistanceToCamera, got.data(), nullptr));
  for (std::size_t j = 0; j < key.size(); ++j) ok[key[j]] = got[j];
  return ok;
}

// VideoSfMHandler::evalTracks(sess, frameKey) with the predicate evaluated on the device
template <typename Session, typename Options>
EvalTracksCount evalTracks(rsba_problem* h, Session& sess, std::size_t frameKey, const Options& opt, bool drop_when = true) {
  return applyEvalTracks(sess, frameKey, validateFrame(h, sess, frameKey, opt), opt, drop_when);
}

// reproject(sess, f, opt, pt, obs) (struct/VideoSfM.cc:139-155) of the 3-D points of `tracks` onto frame frameKey:
// proj_xy [n][2], ok [n].  `h` must hold a scene with this session's camera (its parameters are replaced).
template <typename Session, typename Options>
void reprojectPoints(rsba_problem* h, Session& sess, std::size_t frameKey, const Options& opt, const std::vector<int>& tracks,
                     std::vector<double>* proj_xy, std::vector<unsigned char>* ok) {
  auto& f = sess.frames[frameKey];
  if (f.poses.empty()) throw std::runtime_error("empty frame");
  if (f.poses.size() > 2) throw std::runtime_error("rsba_cuda: frames with one pose per scan line (fullDoF) are not supported");
  const std::size_t n = tracks.size();
  proj_xy->assign(2 * n, 0.0);
  ok->assign(n, 0);
  if (n == 0) return;
  std::vector<double> pts(3 * n), xy(2 * n, 0.0);
  std::vector<int> fr(n, 0), pt(n);
  for (std::size_t j = 0; j < n; ++j) {
    auto& t = sess.getTrack(tracks[j]);
    for (int k = 0; k < 3; ++k) pts[3 * j + k] = t.pt[k];
    pt[j] = (int)j;
  }
  double poses[12];
  for (int k = 0; k < 6; ++k) {
    poses[k] = f.poses[0][k];
    poses[6 + k] = f.poses[f.poses.size() == 2 ? 1 : 0][k];
  }
  const bool rs = f.poses.size() == 2;
  const int scan[2] = {(int)sess.scanlines[0], (int)sess.scanlines[1]};
  session_check(rsba_cuda_set_camera(h, sess.cam.data(), rs ? (int)sess.rs : 0, scan, opt.model.interpolateRotation ? 1 : 0));
  session_check(rsba_cuda_set_scene(h, (long)n, xy.data(), fr.data(), pt.data(), 1, (int)n, nullptr, nullptr));
  session_check(rsba_cuda_set_parameters(h, poses, pts.data()));
  session_check(rsba_cuda_reproject(h, (long)n, fr.data(), pt.data(), (double)opt.tracks.sqrdThreshold, proj_xy->data(), ok->data()));
}

}  // namespace rsba_cuda
#endif  // RSBA_CUDA_SESSION_HPP_
