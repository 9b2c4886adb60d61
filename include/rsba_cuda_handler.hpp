// rsba_cuda_handler.hpp -- header-only C++11 host side over the C ABI (rsba_cuda.h).
//
// Two layers, both mirroring what henrique/rsba's driver uses (paths under src/rsba/):
//
//  * rsba_cuda::Problem  -- the slice of ceres::Problem / ceres::Solve that CeresHandler touches
//      AddResidualBlock(RsBundleAdjustment::Create(sess,opt,obs), loss, pose0, pose1, point)
//                                                   CeresHandler.h:250-255  -> AddRsResidualBlock
//      SetParameterBlockConstant(double*)           CeresHandler.h:283,299,344-345
//      SetParameterization(p, SubsetParameterization(6, constant))   CeresHandler.h:350-381
//      Evaluate(EvaluateOptions(), &cost, ...)      CeresHandler.h:386
//      ceres::Solve(options, &problem, &summary)    CeresHandler.h:419
//    Block identity is pointer identity and the caller owns the parameter memory, exactly as
//    with Ceres; results are written back in place.
//
//  * rsba_cuda::Handler<Session, Options> -- a CeresHandler-shaped class (Add(frameKey, sess, uninitialized),
//    solve(options)) for the 3-D-point configurations (CeresHandler.h:92-392): pose initialisation /
//    velocity extrapolation of a frame without poses (:99-144), motion and pose priors (:146-205), the
//    residual loop incl. the match fallback for observations without a usable track (:223-241) and
//    revalidateReprojections (:244-248), constancy / SubsetParameterization rules (:288-300, :335-382).  It is
//    a template over the session / option types: tests/tools/handler_ref_check.cc instantiates it with the
//    reference's OWN Thrift-generated gen::Session (gen-cpp/sfm_types.h) and SfmOptions (SfmOptions.h),
//    tests/tools/handler_check.cc with plain look-alike structs.  What the device path does not cover
//    THROWS std::runtime_error -- nothing is skipped silently (the reference aborts on its own unsupported
//    modes, CeresHandler.h:247): the structure-less feature-ray mode (:303-332), per-frame intrinsics blocks
//    f.cam, frames with more than two poses (fullDoF), constVelocity, and the SphericalPrior scale hack the
//    reference adds while it initialises frame 1 from an all-zero frame 0 (:36-52, :127-130).
#ifndef RSBA_CUDA_HANDLER_HPP_
#define RSBA_CUDA_HANDLER_HPP_

#include <cmath>
#include <cstddef>
#include <deque>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "rsba_cuda.h"
#include "rsba_cuda_functors.hpp"

namespace rsba_cuda {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

class Problem {
 public:
  explicit Problem(int device = 0) {
    check(rsba_cuda_create(&h_, device));
  }
  ~Problem() { rsba_cuda_destroy(h_); }
  Problem(const Problem&) = delete;
  Problem& operator=(const Problem&) = delete;

  // sess.cam, sess.rs, sess.scanlines, opt.model.interpolateRotation  (VideoSfmBaRs.h:16-22,32-33)
  void SetCamera(const double cam9[9], int shutter, const int scanlines[2], bool interpolate_rotation) {
    check(rsba_cuda_set_camera(h_, cam9, shutter, scanlines, interpolate_rotation ? 1 : 0));
  }
  // lossFunction = new ceres::HuberLoss(a), applied to every residual block (CeresHandler.h:85-90)
  void SetHuberLoss(double a) { check(rsba_cuda_set_loss(h_, a)); }
  // CreateWithCam <2; 9, 6, 6, 3>: the shared intrinsics block is optimised too (CeresHandler.h:256-264)
  void AddRsResidualBlockWithIntrinsics(const double observed[2], double* cam, double* pose0, double* pose1, double* point) {
    check(rsba_cuda_add_rs_residual_with_intrinsics(h_, observed, cam, pose0, pose1, point));
    ++num_residual_blocks_;
    note(cam); note(pose0); note(pose1); note(point);
  }
  void AddRsResidualBlock(const double observed[2], double* pose0, double* pose1, double* point) {
    check(rsba_cuda_add_rs_residual(h_, observed, pose0, pose1, point));
    ++num_residual_blocks_;
    note(pose0); note(pose1); note(point);
  }
  // ceres::Problem::AddParameterBlock for the two control poses of a frame (a frame with pose priors only)
  void AddFrameBlocks(double* pose0, double* pose1) {
    check(rsba_cuda_add_frame_blocks(h_, pose0, pose1));
    note(pose0); note(pose1);
  }
  // RsConstVeloPrior (1) / RsConstAccelerationPrior (2) with a constant ratio (CeresHandler.h:148-186)
  void AddMotionPrior(int kind, double scale, double ratio, double* pose0, double* end0, double* pose1, double* end1) {
    check(rsba_cuda_add_motion_prior(h_, kind, scale, ratio, pose0, end0, pose1, end1));
    note(pose0); note(end0); note(pose1); note(end1);
  }
  // GoodPosePrior between a prior block and a control pose; both are parameter blocks (CeresHandler.h:188-204)
  void AddPosePrior(double rotation, double position, double* prior_block, double* pose_block) {
    check(rsba_cuda_add_pose_prior(h_, rotation, position, prior_block, pose_block));
    note(prior_block); note(pose_block);
  }
  // the shared `&opt.ceres.interFrameRatio` block left variable, with its lower bound (CeresHandler.h:156-180)
  void SetInterFrameRatioBlock(double* ratio) { check(rsba_cuda_set_inter_frame_ratio_block(h_, ratio)); note(ratio); }
  void SetParameterBlockConstant(double* block) { check(rsba_cuda_set_block_constant(h_, block)); }
  void SetSubsetConstant(double* pose_block, const std::vector<int>& constant) {
    check(rsba_cuda_set_subset_constant(h_, pose_block, (int)constant.size(), constant.data()));
  }
  long NumResidualBlocks() const { return num_residual_blocks_; }
  long NumParameterBlocks() const { return (long)blocks_.size(); }      // ceres::Problem::NumParameterBlocks (CeresHandler.h:97,335)

  // problem.Evaluate: cost = 1/2 sum r^2; false if a functor returned false (cam.h:410-412)
  bool Evaluate(double* cost, std::vector<double>* residuals = nullptr, std::vector<double>* jacobian = nullptr) {
    if (residuals) residuals->assign(2 * (size_t)num_residual_blocks_, 0.0);
    if (jacobian) jacobian->assign(30 * (size_t)num_residual_blocks_, 0.0);
    const int rc = rsba_cuda_evaluate(h_, cost, residuals ? residuals->data() : nullptr,
                                      jacobian ? jacobian->data() : nullptr, nullptr);
    if (rc == RSBA_ERR_EVALUATION_FAILED) return false;
    check(rc);
    return true;
  }

  // ceres::Solve.  Like Ceres, failure is reported through the summary (usable == 0), not thrown.
  rsba_solve_summary Solve(const rsba_solve_options& options) {
    rsba_solve_summary s;
    const int rc = rsba_cuda_solve(h_, &options, &s);
    if (rc != RSBA_OK && rc != RSBA_ERR_EVALUATION_FAILED && rc != RSBA_ERR_LINEAR_SOLVER) check(rc);
    return s;
  }

  static rsba_solve_options DefaultOptions() {
    rsba_solve_options o;
    rsba_cuda_default_options(&o);
    return o;
  }
  rsba_problem* handle() { return h_; }

 private:
  static void check(int rc) {
    if (rc != RSBA_OK) throw Error(rc, rsba_cuda_last_error());
  }
  void note(const double* block) { blocks_.insert(block); }
  rsba_problem* h_ = nullptr;
  long num_residual_blocks_ = 0;
  std::set<const double*> blocks_;      // block identity = pointer identity, as in ceres::Problem
};

// The same surface over rsba_cuda_create_multi: ONE host thread, several GPUs of the node.  Every builder call is
// forwarded to the handle of every device (each keeps the share of the points it owns), Solve() runs the ranks'
// LM loops on worker threads inside the library and returns rank 0's summary; results are written back into the
// caller's blocks by rank 0.  `Handler<Session, Options, MultiGpuProblem> cs(opt, startFrame, n_gpus)` is what lets
// the reference's single-threaded VideoSfMHandler::BA (VideoSfMHandler.cc:574-631) use the whole box.
class MultiGpuProblem {
 public:
  // n_devices GPUs: devices 0 .. n_devices - 1
  explicit MultiGpuProblem(int n_devices = 1) {
    if (n_devices < 1) throw Error(RSBA_ERR_INVALID_ARGUMENT, "MultiGpuProblem: at least one device");
    std::vector<int> dev(n_devices);
    for (int k = 0; k < n_devices; ++k) dev[k] = k;
    check(rsba_cuda_create_multi(&m_, dev.data(), n_devices));
  }
  ~MultiGpuProblem() { rsba_cuda_destroy_multi(m_); }
  MultiGpuProblem(const MultiGpuProblem&) = delete;
  MultiGpuProblem& operator=(const MultiGpuProblem&) = delete;

  void SetCamera(const double cam9[9], int shutter, const int scanlines[2], bool interpolate_rotation) {
    each([&](rsba_problem* h) { return rsba_cuda_set_camera(h, cam9, shutter, scanlines, interpolate_rotation ? 1 : 0); });
  }
  void SetHuberLoss(double a) { each([&](rsba_problem* h) { return rsba_cuda_set_loss(h, a); }); }
  void AddRsResidualBlockWithIntrinsics(const double observed[2], double* cam, double* pose0, double* pose1, double* point) {
    each([&](rsba_problem* h) { return rsba_cuda_add_rs_residual_with_intrinsics(h, observed, cam, pose0, pose1, point); });
    ++num_residual_blocks_;
    note(cam); note(pose0); note(pose1); note(point);
  }
  void AddRsResidualBlock(const double observed[2], double* pose0, double* pose1, double* point) {
    each([&](rsba_problem* h) { return rsba_cuda_add_rs_residual(h, observed, pose0, pose1, point); });
    ++num_residual_blocks_;
    note(pose0); note(pose1); note(point);
  }
  void AddFrameBlocks(double* pose0, double* pose1) {
    each([&](rsba_problem* h) { return rsba_cuda_add_frame_blocks(h, pose0, pose1); });
    note(pose0); note(pose1);
  }
  void AddMotionPrior(int kind, double scale, double ratio, double* pose0, double* end0, double* pose1, double* end1) {
    each([&](rsba_problem* h) { return rsba_cuda_add_motion_prior(h, kind, scale, ratio, pose0, end0, pose1, end1); });
    note(pose0); note(end0); note(pose1); note(end1);
  }
  void AddPosePrior(double rotation, double position, double* prior_block, double* pose_block) {
    each([&](rsba_problem* h) { return rsba_cuda_add_pose_prior(h, rotation, position, prior_block, pose_block); });
    note(prior_block); note(pose_block);
  }
  void SetInterFrameRatioBlock(double* ratio) {
    each([&](rsba_problem* h) { return rsba_cuda_set_inter_frame_ratio_block(h, ratio); });
    note(ratio);
  }
  void SetParameterBlockConstant(double* block) { each([&](rsba_problem* h) { return rsba_cuda_set_block_constant(h, block); }); }
  void SetSubsetConstant(double* pose_block, const std::vector<int>& constant) {
    each([&](rsba_problem* h) { return rsba_cuda_set_subset_constant(h, pose_block, (int)constant.size(), constant.data()); });
  }
  long NumResidualBlocks() const { return num_residual_blocks_; }
  long NumParameterBlocks() const { return (long)blocks_.size(); }
  rsba_solve_summary Solve(const rsba_solve_options& options) {
    rsba_solve_summary s;
    const int rc = rsba_cuda_multi_solve(m_, &options, &s);
    if (rc != RSBA_OK && rc != RSBA_ERR_EVALUATION_FAILED && rc != RSBA_ERR_LINEAR_SOLVER) check(rc);
    return s;
  }
  static rsba_solve_options DefaultOptions() {
    rsba_solve_options o;
    rsba_cuda_default_options(&o);
    return o;
  }
  int size() const { return rsba_cuda_multi_size(m_); }
  rsba_problem* handle(int rank = 0) { return rsba_cuda_multi_handle(m_, rank); }

 private:
  static void check(int rc) {
    if (rc != RSBA_OK) throw Error(rc, rsba_cuda_last_error());
  }
  template <typename Fn>
  void each(Fn fn) {
    for (int r = 0; r < rsba_cuda_multi_size(m_); ++r) check(fn(rsba_cuda_multi_handle(m_, r)));
  }
  void note(const double* block) { blocks_.insert(block); }
  rsba_multi* m_ = nullptr;
  long num_residual_blocks_ = 0;
  std::set<const double*> blocks_;
};

// validate(sess, f, opt, pt, obs) of the reference (struct/VideoSfM.cc:159-169) on the host: the pose at the
// observation's own scan line (getPose, :102-131; x or y by shutter direction), the point at least
// minDistanceToCamera away from the camera centre, in front of the camera, and re-projected within
// sqrt(sqrdThreshold) pixels (::vision::validate, mat/cam.h:445-457).  Same arithmetic as the device sweep
// rsba_cuda_validate (validate_kernel).
template <typename Session, typename Frame, typename Options>
inline bool validate(const Session& sess, const Frame& f, const Options& opt, const double pt[3], const double obs[2]) {
  if (f.poses.empty()) throw std::runtime_error("empty frame");                 // struct/VideoSfM.cc:104
  if (f.poses.size() > 2) throw std::runtime_error("rsba_cuda: frames with one pose per scan line (fullDoF) are not supported");
  rsba::CameraModel cm;
  const double* cam = f.__isset.cam ? f.cam.data() : sess.cam.data();
  for (int k = 0; k < 9; ++k) cm.cam[k] = cam[k];
  const bool rs = f.poses.size() == 2;
  cm.shutter = rs ? (int)sess.rs : 0;
  cm.scan0 = rs ? (double)sess.scanlines[0] : 0.0;
  cm.scan_span = rs ? (double)(sess.scanlines[1] - sess.scanlines[0]) : 1.0;
  cm.interp_rot = opt.model.interpolateRotation ? 1 : 0;
  cm.huber = 0.0;
  cm.cam_offset = -1;
  double frame[12];
  for (int k = 0; k < 6; ++k) {
    frame[k] = f.poses[0][k];
    frame[6 + k] = f.poses[rs ? 1 : 0][k];
  }
  const rsba::Proj p = rsba::reproject<false, true>(cm, obs[0], obs[1], frame, pt[0], pt[1], pt[2], nullptr, nullptr);
  double tau = 0.0;
  if (cm.shutter != 0) {
    tau = ((cm.shutter == 2 ? obs[1] : obs[0]) - cm.scan0) / cm.scan_span;
    tau = tau < 0.0 ? 0.0 : (tau > 1.0 ? 1.0 : tau);
  }
  double d2 = 0.0;
  for (int k = 0; k < 3; ++k) {
    const double d = frame[3 + k] + (frame[9 + k] - frame[3 + k]) * tau - pt[k];
    d2 += d * d;
  }
  return std::sqrt(d2) >= (double)opt.tracks.minDistanceToCamera && p.ok &&
         p.r0 * p.r0 + p.r1 * p.r1 < opt.tracks.sqrdThreshold;
}

// CeresHandler-shaped front end (CeresHandler.h:75-426).  Session must offer what the reference's
// sfm::Session offers on this path: frames[k].poses (vector<vector<double>>, size 2 for a rolling-
// shutter frame), frames[k].obs[i].{x, y, track, __isset.track}, getTrack(id) -> {pt, valid,
// __isset.pt, obs[j].frame}, cam, rs, scanlines.  Options: model.{use3Dpoints, calibrated,
// constVelocity, interpolateRotation}, ceres.{huberLoss, const3d, fixFirstNCameras, fixScale,
// fixRotation, fixPosition, useOnlyValidMatches, revalidateReprojections, constFrameVelocity,
// constFrameAcceleration, interFrameRatio, trustPriorCamRotation, trustPriorCamPosition},
// model.rolling_shutter, tracks.{sqrdThreshold, minDistanceToCamera}; frames[k].priorPoses, .cam,
// .__isset.{poses, priorPoses, cam}, obs[i].matches[j].{frame, obs}.
// ProblemT: rsba_cuda::Problem (the GPU), or anything with the same member functions -- the host-only
// tests pass a recorder to check WHICH calls Add() makes without a device.
template <typename Session, typename Options, typename ProblemT = Problem>
class Handler {
 public:
  ProblemT problem;
  Options opt;
  std::size_t startFrame;

  explicit Handler(const Options& o, std::size_t start = 0, int device = 0) : problem(device), opt(o), startFrame(start) {
    if (opt.ceres.huberLoss > 0) problem.SetHuberLoss(opt.ceres.huberLoss);      // CeresHandler.h:85-90
    if (!opt.model.use3Dpoints) throw std::runtime_error("rsba_cuda: structure-less (feature ray) mode is out of scope");
  }

  void Add(const std::size_t frameKey, Session& sess, bool uninitialized = false) {
    if (!camera_set_) {
      const int scan[2] = {(int)sess.scanlines[0], (int)sess.scanlines[1]};
      problem.SetCamera(sess.cam.data(), (int)sess.rs, scan, opt.model.interpolateRotation);
      camera_set_ = true;
    }
    auto& f = sess.frames[frameKey];
    const long formerParamNum = problem.NumParameterBlocks();               // CeresHandler.h:97

    if (!f.__isset.poses) {                                                  // :99-144 initialise the camera frame
      if (frameKey > 0) {
        auto& f_1 = sess.frames[frameKey - 1];
        f.poses = f_1.poses;                                                 // use the last pose as reference
        if (frameKey > 1) {                                                  // extrapolate the linear velocity
          auto& f_2 = sess.frames[frameKey - 2];
          for (std::size_t pi = 0; pi < f.poses.size(); ++pi)
            for (int k = 0; k < 6; ++k)                                      // minus6 + plus6 (:110-113)
              f.poses[pi][k] = f_1.poses[pi][k] + (f_1.poses[pi][k] - f_2.poses[pi][k]);
        } else {
          bool originFrame = true;
          for (auto& pose : f_1.poses)
            for (double p : pose)
              if (p != 0) originFrame = false;
          for (auto& pose : f.poses) {
            pose[3] += 1e-4;
            pose[4] += 1e-4;
            pose[5] += 1e-4;
          }
          if (frameKey == 1 && originFrame)
            throw std::runtime_error("rsba_cuda: the SphericalPrior that CeresHandler adds to frame 1 while it initialises "
                                     "it from an all-zero frame 0 (CeresHandler.h:36-52,127-130) is not on the device "
                                     "path; initialise the first two frames before the first GPU bundle adjustment");
        }
      } else {                                                               // frameKey == 0: zeros
        f.poses.assign(opt.model.rolling_shutter ? 2 : 1, std::vector<double>(6, 0.0));
      }
      uninitialized = true;
      f.__isset.poses = true;
    }

    // "good initial guess" priors (:188-204); a size mismatch re-seeds the poses from the priors
    const bool pose_priors = frameKey >= (std::size_t)opt.ceres.fixFirstNCameras &&
                             (opt.ceres.trustPriorCamRotation != 0 || opt.ceres.trustPriorCamPosition != 0) &&
                             f.__isset.priorPoses && f.priorPoses.size() > 0;
    if (pose_priors && f.poses.size() != f.priorPoses.size()) f.poses = f.priorPoses;
    // A frame with ONE pose is the reference's global-shutter case: ReprojectionError <2; 6, 3> on
    // getPose() = f.poses[0] (:265-286, struct/VideoSfM.cc:104-106).  The device path always has two
    // control-pose blocks per frame; with a GLOBAL shutter the second one does not enter the projection
    // (mat/cam.h:321-322), so a constant stand-in block completes the frame.
    double* second_pose = nullptr;
    if (f.poses.size() == 1) {
      if ((int)sess.rs != 0) throw std::runtime_error("rsba_cuda: single-pose frames need a GLOBAL-shutter session");
      auto it = gs_second_.find(frameKey);
      if (it == gs_second_.end()) {
        gs_store_.push_back(f.poses[0]);
        it = gs_second_.emplace(frameKey, gs_store_.back().data()).first;
      }
      second_pose = it->second;
    } else if (f.poses.size() == 2) {
      second_pose = f.poses[1].data();
    } else {
      throw std::runtime_error("rsba_cuda: frames must hold one (global shutter) or two (rolling shutter) poses");
    }
    if (f.poses.size() == 2 && opt.model.constVelocity) throw std::runtime_error("rsba_cuda: constVelocity (the reference aborts here too)");
    if (f.__isset.cam) throw std::runtime_error("rsba_cuda: per-frame intrinsics (f.cam) are not on the device path");

    // motion prior between this frame and the previous one (:147-186)
    if (frameKey >= (std::size_t)opt.ceres.fixFirstNCameras && frameKey > 0 &&
        (opt.ceres.constFrameVelocity != 0 || opt.ceres.constFrameAcceleration != 0)) {
      auto& f_1 = sess.frames[frameKey - 1];
      if (f.poses.size() == 2 && f_1.poses.size() == 2) {
        const bool accel = opt.ceres.constFrameAcceleration != 0;
        // the ratio block is this handler's own copy of the options, as in the reference (`opt` is copied,
        // :79); it is constant only when it differs from 1 (:178-180)
        if (opt.ceres.interFrameRatio == 1 && !ratio_block_set_) {
          problem.SetInterFrameRatioBlock(&opt.ceres.interFrameRatio);
          ratio_block_set_ = true;
        }
        problem.AddMotionPrior(accel ? 2 : 1, accel ? opt.ceres.constFrameAcceleration : opt.ceres.constFrameVelocity,
                               opt.ceres.interFrameRatio, f.poses[0].data(), f.poses[1].data(), f_1.poses[0].data(),
                               f_1.poses[1].data());
        if (frameKey - 1 < (std::size_t)opt.ceres.fixFirstNCameras)          // :182-186
          for (auto& pose : f_1.poses) problem.SetParameterBlockConstant(pose.data());
      }
    }
    if (pose_priors) {
      problem.AddFrameBlocks(f.poses[0].data(), second_pose);   // the prior names one block: tell the library the frame
      for (std::size_t i = 0; i < f.poses.size(); ++i)
        problem.AddPosePrior(opt.ceres.trustPriorCamRotation, opt.ceres.trustPriorCamPosition,
                             f.priorPoses[i].data(), f.poses[i].data());
    }

    for (auto& o : f.obs) {                                   // :208
      const double obs[2] = {o.x, o.y};
      decltype(&sess.getTrack(0)) t = nullptr;
      if (o.__isset.track) {                                  // :217-220
        t = &sess.getTrack(o.track);
        if (!t->__isset.pt || (opt.ceres.useOnlyValidMatches && !t->valid)) t = nullptr;
      }
      if (!t && (uninitialized || !opt.ceres.useOnlyValidMatches)) {   // :223-241 also add bad reprojections:
        for (auto& ref : o.matches) {                                   // the track of a matched observation, if
          const auto& f2 = sess.frames[ref.frame];                      // its point re-projects onto this one
          const auto& o2 = f2.obs[ref.obs];
          if (o2.__isset.track) {
            t = &sess.getTrack(o2.track);
            if (!t->__isset.pt || (opt.ceres.useOnlyValidMatches && !t->valid) ||
                !validate(sess, f, opt, t->pt.data(), obs)) {
              t = nullptr;
            } else {
              break;   // found!
            }
          }
        }
      }
      if (!t || !(t->valid || !opt.ceres.useOnlyValidMatches)) continue;   // :243
      if (opt.ceres.revalidateReprojections && !validate(sess, f, opt, t->pt.data(), obs)) continue;   // :244-248
      if (opt.model.calibrated) {
        problem.AddRsResidualBlock(obs, f.poses[0].data(), second_pose, t->pt.data());   // :250-255 / :265-270
      } else {                                                                           // :256-264 / :271-277
        problem.AddRsResidualBlockWithIntrinsics(obs, sess.cam.data(), f.poses[0].data(), second_pose, t->pt.data());
      }
      if (f.poses.size() == 1 && frameKey < (std::size_t)opt.ceres.fixFirstNCameras)     // :280-283
        problem.SetParameterBlockConstant(f.poses[0].data());
      bool fixedOldTrack = false;                             // :288-300
      if (startFrame > 0)
        for (auto& ref : t->obs)
          if ((std::size_t)ref.frame < startFrame) { fixedOldTrack = true; break; }
      if (fixedOldTrack || opt.ceres.const3d) problem.SetParameterBlockConstant(t->pt.data());
    }

    if (problem.NumParameterBlocks() <= formerParamNum) return;            // :335
    if (f.poses.size() == 1) {
      problem.AddFrameBlocks(f.poses[0].data(), second_pose);
      problem.SetParameterBlockConstant(second_pose);                      // the stand-in never moves
    }
    if (frameKey < (std::size_t)opt.ceres.fixFirstNCameras) {             // :342-346
      if (f.poses.size() == 2) {
        problem.AddFrameBlocks(f.poses[0].data(), f.poses[1].data());
        problem.SetParameterBlockConstant(f.poses[0].data());
        problem.SetParameterBlockConstant(f.poses[1].data());
      }   // else already constant
    } else if (opt.ceres.fixScale && (frameKey == 0 || frameKey == sess.frames.size() - 1)) {   // :350-361
      problem.SetSubsetConstant(frameKey == 0 ? f.poses[0].data() : f.poses.back().data(), {3, 4, 5});
    } else if (opt.ceres.fixRotation) {                                    // :362-371
      for (auto& pose : f.poses) problem.SetSubsetConstant(pose.data(), {0, 1, 2});
    } else if (opt.ceres.fixPosition) {                                    // :372-382
      for (auto& pose : f.poses) problem.SetSubsetConstant(pose.data(), {3, 4, 5});
    }
  }

  // CeresHandler::solve (:394-426): SPARSE_SCHUR, progress to stdout, 50 iterations by default
  rsba_solve_summary solve(const rsba_solve_options* options = nullptr) {
    rsba_solve_options tmp = ProblemT::DefaultOptions();
    if (!options) {
      tmp.verbose = 1;
      tmp.max_num_iterations = 50;
      options = &tmp;
    }
    return problem.Solve(*options);
  }

 private:
  bool camera_set_ = false, ratio_block_set_ = false;
  std::deque<std::vector<double>> gs_store_;        // stand-in second poses of single-pose frames
  std::map<std::size_t, double*> gs_second_;
};

}  // namespace rsba_cuda
#endif  // RSBA_CUDA_HANDLER_HPP_
