// Test driver for include/rsba_cuda_handler.hpp: a CeresHandler-shaped Add()/solve() over plain
// structs that have the members rsba's Thrift-generated gen::Session offers on this path
// (gen-cpp/sfm_types.h: Frame.poses, Frame.obs, Observation.{x,y,track}, Track.{pt,valid,obs},
// Session.{cam,rs,scanlines,frames}).  Reads a flat scene file written by tests/test_gpu_handler.py,
// runs the same calls VideoSfMHandler::BA makes (VideoSfMHandler.cc:586-592), writes the result.
//   usage: handler_check <scene.bin> <out.bin> <fixFirstNCameras> <maxIter> [startFrame] [gs]
//   (gs = 1: global-shutter session, every frame holds ONE pose -- the first control pose of the file;
//    gs = 2: uncalibrated, opt.model.calibrated = false -- the optimised sess.cam is appended to the output;
//    gs = 3: constant-velocity priors with the default, free interFrameRatio -- the ratio is appended;
//    gs = 4: GoodPosePrior on every frame past the first, priorPoses = the initial poses + a fixed offset --
//            the (free, hence moved) prior blocks are appended;
//    gs = 5: as 4, and the LAST frame loses its observations: a frame that carries pose priors only, which Ceres
//            accepts (its poses move to where the priors pull them);
//    gs = 6: the sweeps of include/rsba_cuda_session.hpp -- every third observation is pushed 60 px away, then per frame
//            validateFrame (device) is compared with the host validate() of rsba_cuda_handler.hpp, evalTracks runs with
//            the reference's polarity, reprojectPoints is compared with the observations; prints "session ..." lines;
//    gs = 7: bulk path -- SessionSoA gather / upload / rsba_cuda_solve / download / scatter instead of Add();
//    gs = 8: Handler over MultiGpuProblem with every visible GPU (one host thread, N devices))
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "rsba_cuda_handler.hpp"
#include "rsba_cuda_session.hpp"

namespace mock {
struct IsSetObs { bool track = false; };
struct IsSetFrame { bool cam = false, priorPoses = false, poses = false; };
struct ObservationRef { int frame = 0, obs = 0; };
struct Observation { double x = 0, y = 0; int track = -1; IsSetObs __isset; std::vector<ObservationRef> matches; };
struct IsSetTrack { bool pt = false; };
struct Track { std::vector<double> pt; bool valid = true; IsSetTrack __isset; std::vector<ObservationRef> obs; };
struct Frame { std::vector<std::vector<double>> poses, priorPoses; std::vector<Observation> obs; std::vector<double> cam; IsSetFrame __isset; };
struct Session {
  std::vector<double> cam;
  int rs = 1;
  std::vector<int> scanlines;
  std::vector<Frame> frames;
  std::map<int, Track> tracks;
  Track& getTrack(int id) { return tracks.at(id); }
};
struct Options {
  struct { bool use3Dpoints = true, calibrated = true, constVelocity = false, interpolateRotation = true, rolling_shutter = true; } model;
  struct { double sqrdThreshold = 16.0; unsigned minDistanceToCamera = 0, minReprojections = 3; } tracks;
  struct {
    double huberLoss = 0, constFrameVelocity = 0, constFrameAcceleration = 0, interFrameRatio = 1;
    double trustPriorCamRotation = 0, trustPriorCamPosition = 0;
    bool const3d = false, fixScale = false, fixRotation = false, fixPosition = false, useOnlyValidMatches = false;
    bool revalidateReprojections = false;
    unsigned fixFirstNCameras = 1;
  } ceres;
};
}  // namespace mock

static void rd(FILE* f, void* p, size_t n) {
  if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  long hdr[4];  // frames, points, observations, shutter
  rd(f, hdr, sizeof(hdr));
  const long F = hdr[0], P = hdr[1], N = hdr[2];
  mock::Session sess;
  sess.rs = (int)hdr[3];
  sess.cam.resize(9);
  rd(f, sess.cam.data(), 9 * sizeof(double));
  sess.scanlines.resize(2);
  rd(f, sess.scanlines.data(), 2 * sizeof(int));
  std::vector<double> poses(12 * F), points(3 * P), xy(2 * N);
  std::vector<int> fr(N), pt(N);
  rd(f, poses.data(), poses.size() * sizeof(double));
  rd(f, points.data(), points.size() * sizeof(double));
  rd(f, xy.data(), xy.size() * sizeof(double));
  rd(f, fr.data(), N * sizeof(int));
  rd(f, pt.data(), N * sizeof(int));
  fclose(f);
  const bool gs = argc > 6 && atoi(argv[6]) == 1;
  const bool uncal = argc > 6 && atoi(argv[6]) == 2;
  const bool velo = argc > 6 && atoi(argv[6]) == 3;   // constant-velocity priors, default (free) interFrameRatio
  const bool prior_only = argc > 6 && atoi(argv[6]) == 5;
  const bool good = (argc > 6 && atoi(argv[6]) == 4) || prior_only;   // GoodPosePrior blocks
  if (gs) sess.rs = 0;
  sess.frames.resize(F);
  for (long k = 0; k < F; ++k) {
    sess.frames[k].poses.assign(gs ? 1 : 2, std::vector<double>(6));
    sess.frames[k].__isset.poses = true;
    for (int c = 0; c < 6; ++c) {
      sess.frames[k].poses[0][c] = poses[12 * k + c];
      if (!gs) sess.frames[k].poses[1][c] = poses[12 * k + 6 + c];
    }
  }
  for (long p = 0; p < P; ++p) {
    mock::Track t;
    t.pt.assign(points.begin() + 3 * p, points.begin() + 3 * p + 3);
    t.__isset.pt = true;
    sess.tracks[(int)p] = t;
  }
  for (long i = 0; i < N; ++i) {
    mock::Observation o;
    o.x = xy[2 * i];
    o.y = xy[2 * i + 1];
    o.track = pt[i];
    o.__isset.track = true;
    mock::ObservationRef ref;
    ref.frame = fr[i];
    ref.obs = (int)sess.frames[fr[i]].obs.size();
    sess.tracks[pt[i]].obs.push_back(ref);
    sess.frames[fr[i]].obs.push_back(o);
  }
  mock::Options opt;
  opt.ceres.fixFirstNCameras = (unsigned)atoi(argv[3]);
  opt.model.calibrated = !uncal;
  if (velo) opt.ceres.constFrameVelocity = 10.0;
  if (good) {
    opt.ceres.trustPriorCamRotation = 20.0;
    opt.ceres.trustPriorCamPosition = 6.0;
    for (long k = 0; k < F; ++k) {
      sess.frames[k].priorPoses = sess.frames[k].poses;
      for (auto& pose : sess.frames[k].priorPoses)
        for (int c = 0; c < 6; ++c) pose[c] += (c < 3 ? 1e-3 : 2e-2) * ((k + c) % 3 - 1);
      sess.frames[k].__isset.priorPoses = true;
    }
  }
  if (prior_only) sess.frames[F - 1].obs.clear();
  double ratio_out = 1.0;
  const size_t startFrame = argc > 5 ? (size_t)atol(argv[5]) : 0;

  const int mode = argc > 6 ? atoi(argv[6]) : 0;
  if (mode == 6) {
    try {
      for (long k = 0; k < F; ++k)
        for (size_t oi = 0; oi < sess.frames[k].obs.size(); ++oi)
          if ((k + (long)oi) % 3 == 0) { sess.frames[k].obs[oi].x += 60.0; sess.frames[k].obs[oi].y -= 45.0; }
      rsba_cuda::Problem pb(0);
      long mismatches = 0, checked = 0, dropped = 0, bad_tracks = 0, reproj_bad = 0;
      for (long k = 0; k < F; ++k) {
        const std::vector<unsigned char> ok = rsba_cuda::validateFrame(pb.handle(), sess, (size_t)k, opt);
        for (size_t oi = 0; oi < sess.frames[k].obs.size(); ++oi) {
          const mock::Observation& o = sess.frames[k].obs[oi];
          const double obs[2] = {o.x, o.y};
          const bool want = rsba_cuda::validate(sess, sess.frames[k], opt, sess.getTrack(o.track).pt.data(), obs);
          mismatches += (want != (ok[oi] != 0));
          ++checked;
        }
        // re-projection of the frame's tracks: lands on the (unperturbed) observation within the noise
        std::vector<int> trk;
        for (const auto& o : sess.frames[k].obs) trk.push_back(o.track);
        std::vector<double> proj;
        std::vector<unsigned char> pok;
        rsba_cuda::reprojectPoints(pb.handle(), sess, (size_t)k, opt, trk, &proj, &pok);
        for (size_t oi = 0; oi < trk.size(); ++oi) {
          if ((k + (long)oi) % 3 == 0) continue;
          const double dx = proj[2 * oi] - sess.frames[k].obs[oi].x, dy = proj[2 * oi + 1] - sess.frames[k].obs[oi].y;
          if (!pok[oi] || dx * dx + dy * dy > 25.0) ++reproj_bad;
        }
        const rsba_cuda::EvalTracksCount n = rsba_cuda::evalTracks(pb.handle(), sess, (size_t)k, opt, false);
        dropped += n.observations;
        bad_tracks += n.tracks;
      }
      long still = 0;
      for (long k = 0; k < F; ++k)
        for (const auto& o : sess.frames[k].obs) still += o.__isset.track;
      printf("session checked %ld mismatches %ld dropped %ld bad_tracks %ld kept %ld reproj_bad %ld\n", checked, mismatches,
             dropped, bad_tracks, still, reproj_bad);
      return 0;
    } catch (const std::exception& e) {
      fprintf(stderr, "handler_check: %s\n", e.what());
      return 1;
    }
  }
  rsba_solve_summary s;
  try {
    if (mode == 7) {
      rsba_cuda::SessionSoA<mock::Session> soa;
      soa.gather(sess, 0, (size_t)F - 1, opt.ceres.useOnlyValidMatches, opt.ceres.fixFirstNCameras, startFrame);
      rsba_cuda::Problem pb(0);
      soa.upload(pb.handle(), sess, opt);
      rsba_solve_options o = rsba_cuda::Problem::DefaultOptions();
      o.max_num_iterations = atoi(argv[4]);
      s = pb.Solve(o);
      soa.download(pb.handle());
      soa.scatter(sess);
    } else if (mode == 8) {
      const int n_gpus = rsba_cuda_device_count();
      rsba_cuda::Handler<mock::Session, mock::Options, rsba_cuda::MultiGpuProblem> cs(opt, startFrame, n_gpus);
      for (size_t fi = startFrame; fi < (size_t)F; ++fi) cs.Add(fi, sess);
      rsba_solve_options o = rsba_cuda::Problem::DefaultOptions();
      o.max_num_iterations = atoi(argv[4]);
      s = cs.solve(&o);
      printf("gpus %d\n", cs.problem.size());
    } else {
    rsba_cuda::Handler<mock::Session, mock::Options> cs(opt, startFrame);
    for (size_t fi = startFrame; fi < (size_t)F; ++fi) cs.Add(fi, sess);    // VideoSfMHandler.cc:586-590
    rsba_solve_options o = rsba_cuda::Problem::DefaultOptions();
    o.max_num_iterations = atoi(argv[4]);
    s = cs.solve(&o);                                                        // VideoSfMHandler.cc:592
    ratio_out = cs.opt.ceres.interFrameRatio;
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "handler_check: %s\n", e.what());
    return 1;
  }
  printf("usable %d iterations %d initial %.17g final %.17g blocks %ld: %s\n", s.usable, s.iterations, s.initial_cost,
         s.final_cost, s.num_residual_blocks, s.message);
  FILE* g = fopen(argv[2], "wb");
  if (!g) return 2;
  double head[4] = {(double)s.usable, (double)s.iterations, s.initial_cost, s.final_cost};
  fwrite(head, sizeof(double), 4, g);
  for (long k = 0; k < F; ++k) {
    fwrite(sess.frames[k].poses[0].data(), sizeof(double), 6, g);
    fwrite(sess.frames[k].poses.back().data(), sizeof(double), 6, g);   // (global shutter: pose0 again)
  }
  for (long p = 0; p < P; ++p) fwrite(sess.tracks[(int)p].pt.data(), sizeof(double), 3, g);
  if (uncal) fwrite(sess.cam.data(), sizeof(double), 9, g);
  if (velo) fwrite(&ratio_out, sizeof(double), 1, g);
  if (good)
    for (long k = 0; k < F; ++k)
      for (auto& pose : sess.frames[k].priorPoses) fwrite(pose.data(), sizeof(double), 6, g);
  fclose(g);
  return 0;
}
