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
//  * rsba_cuda::Handler<Session, Options> -- a CeresHandler-shaped class (Add(frameKey, sess),
//    solve(options)) for the rolling-shutter, calibrated, 3-D-point configuration
//    (CeresHandler.h:208-302, 335-382).  It is a template over the session / option types, so
//    it compiles against the Thrift-generated gen::Session of the reference
//    as well as against the plain structs of tests/tools/handler_check.cc.  Configurations the
//    device path does not cover yet throw std::runtime_error (the reference aborts on its own
//    unsupported modes, CeresHandler.h:247).
#ifndef RSBA_CUDA_HANDLER_HPP_
#define RSBA_CUDA_HANDLER_HPP_

#include <cstddef>
#include <deque>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "rsba_cuda.h"

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
  }
  void AddRsResidualBlock(const double observed[2], double* pose0, double* pose1, double* point) {
    check(rsba_cuda_add_rs_residual(h_, observed, pose0, pose1, point));
    ++num_residual_blocks_;
  }
  // RsConstVeloPrior (1) / RsConstAccelerationPrior (2) with a constant ratio (CeresHandler.h:148-186)
  void AddMotionPrior(int kind, double scale, double ratio, double* pose0, double* end0, double* pose1, double* end1) {
    check(rsba_cuda_add_motion_prior(h_, kind, scale, ratio, pose0, end0, pose1, end1));
  }
  // GoodPosePrior between a prior block and a control pose; both are parameter blocks (CeresHandler.h:188-204)
  void AddPosePrior(double rotation, double position, double* prior_block, double* pose_block) {
    check(rsba_cuda_add_pose_prior(h_, rotation, position, prior_block, pose_block));
  }
  // the shared `&opt.ceres.interFrameRatio` block left variable, with its lower bound (CeresHandler.h:156-180)
  void SetInterFrameRatioBlock(double* ratio) { check(rsba_cuda_set_inter_frame_ratio_block(h_, ratio)); }
  void SetParameterBlockConstant(double* block) { check(rsba_cuda_set_block_constant(h_, block)); }
  void SetSubsetConstant(double* pose_block, const std::vector<int>& constant) {
    check(rsba_cuda_set_subset_constant(h_, pose_block, (int)constant.size(), constant.data()));
  }
  long NumResidualBlocks() const { return num_residual_blocks_; }

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
  rsba_problem* h_ = nullptr;
  long num_residual_blocks_ = 0;
};

// CeresHandler-shaped front end (CeresHandler.h:75-426).  Session must offer what the reference's
// sfm::Session offers on this path: frames[k].poses (vector<vector<double>>, size 2 for a rolling-
// shutter frame), frames[k].obs[i].{x, y, track, __isset.track}, getTrack(id) -> {pt, valid,
// __isset.pt, obs[j].frame}, cam, rs, scanlines.  Options: model.{use3Dpoints, calibrated,
// constVelocity, interpolateRotation}, ceres.{huberLoss, const3d, fixFirstNCameras, fixScale,
// fixRotation, fixPosition, useOnlyValidMatches, constFrameVelocity, constFrameAcceleration,
// interFrameRatio, trustPriorCamRotation, trustPriorCamPosition}; frames[k].priorPoses + __isset.priorPoses.
template <typename Session, typename Options>
class Handler {
 public:
  Problem problem;
  Options opt;
  std::size_t startFrame;

  explicit Handler(const Options& o, std::size_t start = 0, int device = 0) : problem(device), opt(o), startFrame(start) {
    if (opt.ceres.huberLoss > 0) problem.SetHuberLoss(opt.ceres.huberLoss);      // CeresHandler.h:85-90
    if (!opt.model.use3Dpoints) throw std::runtime_error("rsba_cuda: structure-less (feature ray) mode is out of scope");
  }

  void Add(const std::size_t frameKey, Session& sess) {
    if (!camera_set_) {
      const int scan[2] = {(int)sess.scanlines[0], (int)sess.scanlines[1]};
      problem.SetCamera(sess.cam.data(), (int)sess.rs, scan, opt.model.interpolateRotation);
      camera_set_ = true;
    }
    auto& f = sess.frames[frameKey];
    // "good initial guess" priors (CeresHandler.h:188-204); a size mismatch re-seeds the poses from the priors
    const bool pose_priors = frameKey >= (std::size_t)opt.ceres.fixFirstNCameras &&
                             (opt.ceres.trustPriorCamRotation != 0 || opt.ceres.trustPriorCamPosition != 0) &&
                             f.__isset.priorPoses && f.priorPoses.size() > 0;
    if (pose_priors && f.poses.size() != f.priorPoses.size()) f.poses = f.priorPoses;
    // A frame with ONE pose is the reference's global-shutter case: ReprojectionError <2; 6, 3> on
    // getPose() = f.poses[0] (CeresHandler.h:265-286, struct/VideoSfM.cc:104-106).  The device path
    // always has two control-pose blocks per frame; with a GLOBAL shutter the second one does not
    // enter the projection (mat/cam.h:321-322), so a constant stand-in block completes the frame.
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
    bool added = false;
    // motion prior between this frame and the previous one (CeresHandler.h:147-186)
    if (frameKey >= (std::size_t)opt.ceres.fixFirstNCameras && frameKey > 0 &&
        (opt.ceres.constFrameVelocity != 0 || opt.ceres.constFrameAcceleration != 0)) {
      auto& f_1 = sess.frames[frameKey - 1];
      if (f.poses.size() == 2 && f_1.poses.size() == 2) {
        const bool accel = opt.ceres.constFrameAcceleration != 0;
        // the ratio block is this handler's own copy of the options, as in the reference (`opt` is copied,
        // CeresHandler.h:79); it is constant only when it differs from 1 (:178-180)
        if (opt.ceres.interFrameRatio == 1 && !ratio_block_set_) {
          problem.SetInterFrameRatioBlock(&opt.ceres.interFrameRatio);
          ratio_block_set_ = true;
        }
        problem.AddMotionPrior(accel ? 2 : 1, accel ? opt.ceres.constFrameAcceleration : opt.ceres.constFrameVelocity,
                               opt.ceres.interFrameRatio, f.poses[0].data(), f.poses[1].data(), f_1.poses[0].data(),
                               f_1.poses[1].data());
        added = true;
        if (frameKey - 1 < (std::size_t)opt.ceres.fixFirstNCameras)          // :182-186
          for (auto& pose : f_1.poses) problem.SetParameterBlockConstant(pose.data());
      }
    }
    if (pose_priors)
      for (std::size_t i = 0; i < f.poses.size(); ++i)
        problem.AddPosePrior(opt.ceres.trustPriorCamRotation, opt.ceres.trustPriorCamPosition,
                             f.priorPoses[i].data(), f.poses[i].data());
    for (auto& o : f.obs) {                                   // CeresHandler.h:208
      if (!o.__isset.track) continue;
      auto* t = &sess.getTrack(o.track);
      if (!t->__isset.pt || (opt.ceres.useOnlyValidMatches && !t->valid)) continue;
      const double obs[2] = {o.x, o.y};
      if (opt.model.calibrated) {
        problem.AddRsResidualBlock(obs, f.poses[0].data(), second_pose, t->pt.data());   // :250-255 / :265-270
      } else {                                                                           // :256-264 / :271-277
        if (f.__isset.cam) throw std::runtime_error("rsba_cuda: per-frame intrinsics (f.cam) are not on the device path");
        problem.AddRsResidualBlockWithIntrinsics(obs, sess.cam.data(), f.poses[0].data(), second_pose, t->pt.data());
      }
      added = true;
      bool fixedOldTrack = false;                             // :288-300
      if (startFrame > 0)
        for (auto& ref : t->obs)
          if ((std::size_t)ref.frame < startFrame) { fixedOldTrack = true; break; }
      if (fixedOldTrack || opt.ceres.const3d) problem.SetParameterBlockConstant(t->pt.data());
    }
    if (!added) return;
    if (f.poses.size() == 1) problem.SetParameterBlockConstant(second_pose);   // the stand-in never moves
    if (frameKey < (std::size_t)opt.ceres.fixFirstNCameras) {             // :342-346, :280-283
      problem.SetParameterBlockConstant(f.poses[0].data());
      if (f.poses.size() == 2) problem.SetParameterBlockConstant(f.poses[1].data());
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
    rsba_solve_options tmp = Problem::DefaultOptions();
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
