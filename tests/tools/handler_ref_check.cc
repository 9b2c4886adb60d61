// Host-only test driver: include/rsba_cuda_handler.hpp instantiated with the REFERENCE'S OWN types --
// vision::sfm::gen::Session / Frame / Observation / Track from src/rsba/gen-cpp/sfm_types.{h,cpp} (Thrift
// generated; compiled in place against the inert Thrift stand-in of tests/shim/thrift) and vision::SfmOptions
// from src/rsba/SfmOptions.h (compiled in place; its <ceres/ceres.h> include resolves to oracle/shim).
// The problem type is a RECORDER: every call that CeresHandler::Add would make on ceres::Problem
// (CeresHandler.h:92-392) is logged, so tests/test_handler_reference_types.py can check the branches of Add()
// -- pose initialisation, match fallback, revalidateReprojections, constancy rules -- without a GPU.
//   usage: handler_ref_check <scene.bin> <mode>       (scene format: tests/test_gpu_handler.py)
// Built only where /root/reference exists (this container); the GPU box runs tests/tools/handler_check instead.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "rsba/SfmOptions.h"
#include "rsba/gen-cpp/sfm_types.h"

#include "rsba_cuda_handler.hpp"
#include "rsba_cuda_session.hpp"

namespace gen = vision::sfm::gen;

// struct/VideoSfM.h:112-124 (that header itself needs glog + OpenCV): the session with its track accessor
struct Session : public gen::Session {
  gen::Track& getTrack(const size_t trackKey) { return tracks[trackKey]; }
  const gen::Track& getTrack(const size_t trackKey) const { return tracks[trackKey]; }
};

// The log is kept with raw block addresses and printed at the end, when every block has a name (the poses of
// a frame that Add() itself initialises do not exist before the call).
static std::map<const double*, std::string> g_names;
static std::vector<std::string> g_log;
static std::string name(const double* p) {
  char buf[32];
  snprintf(buf, sizeof(buf), "@%p@", (const void*)p);
  return buf;
}
static void logf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
static void logf(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_log.push_back(buf);
}
static void flush_log() {
  for (std::string line : g_log) {
    for (size_t a = line.find('@'); a != std::string::npos; a = line.find('@', a)) {
      const size_t b = line.find('@', a + 1);
      const void* ptr = nullptr;
      sscanf(line.substr(a + 1, b - a - 1).c_str(), "%p", (void**)&ptr);
      auto it = g_names.find((const double*)ptr);
      const std::string nm = it == g_names.end() ? "?" : it->second;
      line.replace(a, b - a + 1, nm);
      a += nm.size();
    }
    printf("%s\n", line.c_str());
  }
}

struct Recorder {
  explicit Recorder(int) {}
  void SetCamera(const double cam[9], int shutter, const int scan[2], bool interp) {
    logf("camera %.17g %.17g shutter %d scan %d %d interp %d", cam[0], cam[8], shutter, scan[0], scan[1], (int)interp);
  }
  void SetHuberLoss(double a) { logf("huber %.17g", a); }
  void AddRsResidualBlock(const double obs[2], double* p0, double* p1, double* pt) {
    logf("rs %s %s %s %.17g %.17g", name(p0).c_str(), name(p1).c_str(), name(pt).c_str(), obs[0], obs[1]);
    note(p0); note(p1); note(pt);
  }
  void AddRsResidualBlockWithIntrinsics(const double obs[2], double* cam, double* p0, double* p1, double* pt) {
    logf("rscam %s %s %s %s %.17g %.17g", name(cam).c_str(), name(p0).c_str(), name(p1).c_str(), name(pt).c_str(), obs[0], obs[1]);
    note(cam); note(p0); note(p1); note(pt);
  }
  void AddFrameBlocks(double* p0, double* p1) { logf("frame %s %s", name(p0).c_str(), name(p1).c_str()); note(p0); note(p1); }
  void AddMotionPrior(int kind, double scale, double ratio, double* a, double* b, double* c, double* d) {
    logf("motion %d %.17g %.17g %s %s %s %s", kind, scale, ratio, name(a).c_str(), name(b).c_str(), name(c).c_str(), name(d).c_str());
    note(a); note(b); note(c); note(d);
  }
  void AddPosePrior(double r, double t, double* prior, double* pose) {
    logf("poseprior %.17g %.17g %s %s", r, t, name(prior).c_str(), name(pose).c_str());
    note(prior); note(pose);
  }
  void SetInterFrameRatioBlock(double* r) { logf("ratio_block"); note(r); }
  void SetParameterBlockConstant(double* b) { logf("const %s", name(b).c_str()); }
  void SetSubsetConstant(double* b, const std::vector<int>& c) {
    std::string line = "subset " + name(b);
    for (int k : c) line += " " + std::to_string(k);
    g_log.push_back(line);
  }
  long NumParameterBlocks() const { return (long)blocks.size(); }
  rsba_solve_summary Solve(const rsba_solve_options&) { return rsba_solve_summary(); }
  static rsba_solve_options DefaultOptions() { return rsba_solve_options(); }
  void note(const double* b) { blocks[b] = 1; }
  std::map<const double*, int> blocks;
};

static void rd(FILE* f, void* p, size_t n) {
  if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string mode = argv[2];
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  long hdr[4];
  rd(f, hdr, sizeof(hdr));
  const long F = hdr[0], P = hdr[1], N = hdr[2];
  Session sess;
  sess.rs = (gen::RollingShutter::type)hdr[3];
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

  sess.frames.resize(F);
  for (long k = 0; k < F; ++k) {
    gen::Frame& fk = sess.frames[k];
    fk.poses.assign(2, std::vector<double>(6));
    for (int c = 0; c < 6; ++c) {
      fk.poses[0][c] = poses[12 * k + c];
      fk.poses[1][c] = poses[12 * k + 6 + c];
    }
    fk.__isset.poses = true;
  }
  sess.tracks.resize(P);
  for (long p = 0; p < P; ++p) {
    gen::Track& t = sess.tracks[p];
    t.pt.assign(points.begin() + 3 * p, points.begin() + 3 * p + 3);
    t.__isset.pt = true;
    t.valid = true;
  }
  for (long i = 0; i < N; ++i) {
    gen::Observation o;
    o.x = xy[2 * i];
    o.y = xy[2 * i + 1];
    o.track = pt[i];
    o.__isset.track = true;
    gen::ObservationRef ref;
    ref.frame = fr[i];
    ref.obs = (int)sess.frames[fr[i]].obs.size();
    sess.tracks[pt[i]].obs.push_back(ref);
    sess.frames[fr[i]].obs.push_back(o);
  }

  vision::SfmOptions opt;            // the reference's defaults (SfmOptions.h)
  opt.ceres.fixFirstNCameras = 1;
  size_t startFrame = 0;
  if (mode == "matches") {
    // every observation loses its track and keeps a match to ANOTHER observation of the same point: the
    // fallback of CeresHandler.h:223-241 must find the track through the match iff the point re-projects
    opt.ceres.useOnlyValidMatches = false;
    for (long p = 0; p < P; ++p) {
      const std::vector<gen::ObservationRef>& refs = sess.tracks[p].obs;
      for (size_t a = 0; a < refs.size(); ++a) {
        gen::Observation& o = sess.frames[refs[a].frame].obs[refs[a].obs];
        if ((refs[a].frame + refs[a].obs) % 3 != 0 || refs.size() < 2) continue;
        o.__isset.track = false;
        o.matches.push_back(refs[(a + 1) % refs.size()]);
        o.__isset.matches = true;
      }
    }
  } else if (mode == "revalidate") {
    opt.ceres.revalidateReprojections = true;          // CeresHandler.h:244-248
  } else if (mode == "init") {
    sess.frames[F - 1].poses.clear();                   // CeresHandler.h:99-144: extrapolated from the two before
    sess.frames[F - 1].__isset.poses = false;
  } else if (mode == "window") {
    startFrame = (size_t)(F / 2);                       // windowed BA: CeresHandler.h:288-300
  } else if (mode == "rotation") {
    opt.ceres.fixRotation = true;
    opt.ceres.fixFirstNCameras = 0;
  } else if (mode == "priors") {
    opt.ceres.constFrameAcceleration = 3.0;
    opt.ceres.trustPriorCamRotation = 20.0;
    opt.ceres.trustPriorCamPosition = 6.0;
    for (long k = 0; k < F; ++k) {
      sess.frames[k].priorPoses = sess.frames[k].poses;
      sess.frames[k].__isset.priorPoses = true;
    }
    sess.frames[F - 1].obs.clear();                     // a frame that carries pose priors only
  } else if (mode == "soa") {
    // rsba_cuda_session.hpp with the reference's own types: gather all frames, print what would be uploaded,
    // move every parameter, scatter, print the session's blocks
    sess.tracks[1].valid = false;                       // dropped with only_valid (CeresHandler.h:217-220)
    sess.frames[0].obs[0].__isset.track = false;        // an observation without a track
    rsba_cuda::SessionSoA<Session> soa;
    soa.gather(sess, 0, (size_t)F - 1, true, 1, 0);
    printf("soa %d %d %ld\n", soa.num_frames(), soa.num_points(), soa.num_obs());
    for (long i = 0; i < soa.num_obs(); ++i)
      printf("obs %d %d %d %.17g %.17g\n", soa.obs_frame[i], soa.obs_key[i], soa.point_track[soa.obs_point[i]], soa.obs_xy[2 * i],
             soa.obs_xy[2 * i + 1]);
    for (int k = 0; k < soa.num_frames(); ++k) printf("mask %d %u\n", k, (unsigned)soa.pose_mask[k]);
    for (double& v : soa.poses) v += 0.5;
    for (double& v : soa.points) v -= 0.25;
    soa.scatter(sess);
    for (long k = 0; k < F; ++k) printf("pose %ld %.17g %.17g\n", k, sess.frames[k].poses[0][0], sess.frames[k].poses[1][5]);
    for (long q = 0; q < P; ++q) printf("pt %ld %.17g\n", q, sess.tracks[q].pt[2]);
    return 0;
  } else if (mode == "evaltracks") {
    // the bookkeeping of VideoSfMHandler::evalTracks (VideoSfMHandler.cc:381-405) on the reference's types, driven
    // by a fixed predicate pattern; the default drops on TRUE, as the reference's line 390 reads
    opt.tracks.minReprojections = 4;
    for (long k = 0; k < F; ++k) {
      std::vector<unsigned char> ok(sess.frames[k].obs.size());
      for (size_t oi = 0; oi < ok.size(); ++oi) ok[oi] = ((k + (long)oi) % 3 == 0) ? 1 : 0;
      const rsba_cuda::EvalTracksCount n = rsba_cuda::applyEvalTracks(sess, (size_t)k, ok, opt, (k % 2) == 0);
      printf("frame %ld %u %u\n", k, n.observations, n.tracks);
    }
    for (long k = 0; k < F; ++k)
      for (size_t oi = 0; oi < sess.frames[k].obs.size(); ++oi)
        printf("o %ld %zu %d\n", k, oi, (int)sess.frames[k].obs[oi].__isset.track);
    for (long q = 0; q < P; ++q) {
      printf("t %ld %d", q, (int)sess.tracks[q].valid);
      for (const auto& r : sess.tracks[q].obs) printf(" %d:%d", r.frame, r.obs);
      printf("\n");
    }
    return 0;
  } else if (mode != "plain") {
    return 2;
  }
  // names of the parameter blocks (after the mode's edits; poses of the "init" frame are named below)
  for (long k = 0; k < F; ++k)
    for (size_t i = 0; i < sess.frames[k].poses.size(); ++i) g_names[sess.frames[k].poses[i].data()] = "f" + std::to_string(k) + "p" + std::to_string(i);
  for (long k = 0; k < F; ++k)
    for (size_t i = 0; i < sess.frames[k].priorPoses.size(); ++i) g_names[sess.frames[k].priorPoses[i].data()] = "q" + std::to_string(k) + "p" + std::to_string(i);
  for (long p = 0; p < P; ++p) g_names[sess.tracks[p].pt.data()] = "x" + std::to_string(p);
  g_names[sess.cam.data()] = "cam";

  try {
    rsba_cuda::Handler<Session, vision::SfmOptions, Recorder> cs(opt, startFrame);
    for (size_t fi = startFrame; fi < (size_t)F; ++fi) {                 // VideoSfMHandler.cc:586-590
      logf("add %zu", fi);
      cs.Add(fi, sess);
    }
  } catch (const std::exception& e) {
    logf("exception %s", e.what());
  }
  if (mode == "init") {
    const gen::Frame& fl = sess.frames[F - 1];
    for (size_t i = 0; i < fl.poses.size(); ++i) g_names[fl.poses[i].data()] = "f" + std::to_string(F - 1) + "p" + std::to_string(i);
    std::string line = "initialised " + std::to_string((int)fl.__isset.poses) + " " + std::to_string(fl.poses.size());
    char buf[40];
    for (const auto& pose : fl.poses)
      for (double v : pose) { snprintf(buf, sizeof(buf), " %.17g", v); line += buf; }
    g_log.push_back(line);
  }
  flush_log();
  return 0;
}
