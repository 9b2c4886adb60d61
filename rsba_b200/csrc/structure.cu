// Host analysis of a scene's structure (structure.cuh).  No device calls in this file.
#include "structure.cuh"

#include "api_guard.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/rsba_cuda.h"

namespace rsba {

void set_last_error(const std::string& msg);   // problem.cu
void compute_point_owners(int n_frames, int n_points, long n_obs, const int* obs_frame_sorted, const int* obs_point,
                          int world, std::vector<int>* owner);   // problem.cu

namespace {

int fail(std::string* error, int code, const char* msg) {
  if (error) *error = msg;
  return code;
}
}  // namespace

int analyze_structure(const SceneTopology& sc, HostStructure* out, std::string* error,
                      void (*lap_fn)(const char*, void*), void* lap_ctx) {
  auto lap = [&](const char* what) { if (lap_fn) lap_fn(what, lap_ctx); };
  const long N = sc.n_obs;
  const int F = sc.n_frames, P = sc.n_points;
  const bool free_cam = sc.free_cam, free_ratio = sc.free_ratio && !sc.prior_pairs.empty();
  const bool pseudo = sc.free_cam || sc.free_ratio;
  const int Fc = sc.n_cam_frames();    // + the pseudo-frame (free intrinsics / free interFrameRatio)
  const int* fr = sc.obs_frame;
  const int* pt = sc.obs_point;
  if (F > 65536) return fail(error, RSBA_ERR_INVALID_ARGUMENT, "more than 65536 frames: the tile index (T x T) would not fit; shard the sequence");
  if (N > 2147483647L) return fail(error, RSBA_ERR_INVALID_ARGUMENT, "more than 2^31 observations on one GPU: shard the scene over more GPUs");

  HostThreads pool(N, sc.world);

  // point-major CSR by a stable parallel counting sort (observation order inside a point stays frame order):
  // every thread histograms a contiguous range of observations, the per-(thread, point) start offsets follow from a
  // prefix over points and threads, and each thread scatters its own range.  The frame of each entry travels with
  // it, so that the incidence passes below read sequentially instead of gathering fr[pt_obs[x]].
  std::vector<int>& pt_ptr = out->pt_ptr;
  HostVec<int>& pt_obs = out->pt_obs;
  pt_ptr.assign(P + 1, 0);
  pt_obs.resize(N);
  HostVec<int2> pt_both(N);   // (observation, frame): scattered as ONE 8-byte store
  {
    const int nt = pool.n;
    std::vector<std::vector<int>> hist(nt);
    pool.run([&](int t) {
      hist[t].assign(P, 0);
      const long b = N * t / nt, e = N * (t + 1) / nt;
      int* h = hist[t].data();
      for (long i = b; i < e; ++i) h[pt[i]]++;
    });
    // (a single pass over P x threads counters: 1.6 M at C3 with 8 threads)
    int run = 0;
    for (int p = 0; p < P; ++p) {
      pt_ptr[p] = run;
      for (int t = 0; t < nt; ++t) { const int c = hist[t][p]; hist[t][p] = run; run += c; }
    }
    pt_ptr[P] = run;
    pool.run([&](int t) {
      const long b = N * t / nt, e = N * (t + 1) / nt;
      int* cur = hist[t].data();
      for (long i = b; i < e; ++i) pt_both[cur[pt[i]]++] = make_int2((int)i, fr[i]);
    });
    pool.run([&](int t) {
      const long b = N * t / nt, e = N * (t + 1) / nt;
      for (long i = b; i < e; ++i) pt_obs[i] = pt_both[i].x;
    });
  }
  // frame chunks of <= 128 observations
  std::vector<int>& chunk_frame = out->chunk_frame;
  std::vector<int>& chunk_beg = out->chunk_beg;
  std::vector<int>& chunk_cnt = out->chunk_cnt;
  std::vector<int>& frame_chunk_ptr = out->frame_chunk_ptr;
  chunk_frame.clear(); chunk_beg.clear(); chunk_cnt.clear();
  frame_chunk_ptr.assign(Fc + 1, 0);
  {
    long i = 0;
    for (int f = 0; f < F; ++f) {
      frame_chunk_ptr[f] = (int)chunk_frame.size();
      const long j = std::upper_bound(fr + i, fr + N, f) - fr;   // (sorted by frame)
      for (long b = i; b < j; b += 128) {
        chunk_frame.push_back(f);
        chunk_beg.push_back((int)b);
        chunk_cnt.push_back((int)std::min<long>(128, j - b));
      }
      i = j;
    }
    for (int f = F; f <= Fc; ++f) frame_chunk_ptr[f] = (int)chunk_frame.size();   // the pseudo-frame has no observations
  }
  // groups of whole points (their observations are one contiguous run of the point-major order)
  {
    std::vector<int2>& groups = out->point_groups;
    std::vector<int>& big = out->point_big;
    groups.clear(); big.clear();
    int lo = 0;
    long acc = 0;
    auto flush = [&](int hi) { if (hi > lo) groups.push_back(make_int2(lo, hi)); lo = hi; acc = 0; };
    for (int p = 0; p < P; ++p) {
      const long n = pt_ptr[p + 1] - pt_ptr[p];
      if (n > kPointGroupObs) {            // its observations sit between the neighbours': the group ends here
        flush(p);
        big.push_back(p);
        lo = p + 1;
        continue;
      }
      if (acc + n > kPointGroupObs || p - lo >= kPointGroupPoints) flush(p);
      acc += n;
    }
    flush(P);
  }
  lap("point CSR + frame chunks");
  // ---- Schur SYRK structure: (frame tile, point) incidences, tile pairs, work items
  const int T = (int)((12L * Fc + kTile - 1) / kTile);
  out->T = T;
  const int H = 2 * T;                                   // sub-tiles of 4 frames
  const int Hreal = (Fc + kSubFrames - 1) / kSubFrames;  // ... that hold at least one frame
  const int cam_sub = F / kSubFrames, cam_slot = F % kSubFrames;   // where the pseudo-frame sits
  HostVec<int>& inc_point = out->inc_point;
  HostVec<int>& inc_tile = out->inc_tile;
  HostVec<int>& slot_beg = out->slot_beg;
  std::vector<int>& pt_inc_ptr = out->pt_inc_ptr;
  std::vector<int>& cam_inc = out->cam_inc;
  HostVec<unsigned char>& slot_cnt = out->slot_cnt;
  pt_inc_ptr.assign(P + 1, 0);
  cam_inc.assign(std::max(P, 1), -1);
  // threads own contiguous point ranges holding about the same number of observations each
  std::vector<int> pt_cut(pool.n + 1, P);
  pt_cut[0] = 0;
  for (int t = 1; t < pool.n; ++t)
    pt_cut[t] = (int)(std::lower_bound(pt_ptr.begin(), pt_ptr.end(), (int)(N * t / pool.n)) - pt_ptr.begin());
  for (int t = 1; t <= pool.n; ++t) pt_cut[t] = std::min(P, std::max(pt_cut[t], pt_cut[t - 1]));
  // pass 1: incidences per point (distinct sub-tiles among its frames; + the pseudo-frame's sub-tile when the
  // intrinsics are free: every eliminated point couples with them, and its panel in that sub-tile -- shared with
  // the last real frames when F is not a multiple of 4 -- gets the pseudo-frame rows)
  pool.run([&](int t) {
    for (int p = pt_cut[t]; p < pt_cut[t + 1]; ++p) {
      if (sc.point_const[p]) continue;  // constant points are not eliminated: no Schur term
      const int b = pt_ptr[p], e = pt_ptr[p + 1];
      int n = 0, last_tile = -1;
      for (int x = b; x < e; ++x) {
        const int A = pt_both[x].y / kSubFrames;
        if (A != last_tile) { ++n; last_tile = A; }
      }
      if (free_cam && e > b && last_tile != cam_sub) ++n;
      pt_inc_ptr[p + 1] = n;
    }
  });
  {
    long run = 0;
    for (int p = 0; p < P; ++p) { run += pt_inc_ptr[p + 1]; pt_inc_ptr[p + 1] = (int)std::min<long>(run, 2147483647L); }
    if (run * kPanelDoubles + kPanelDoubles > 2147483647L)
      return fail(error, RSBA_ERR_INVALID_ARGUMENT, "Schur panel buffer exceeds 2^31 doubles: shard the scene over more GPUs");
  }
  const int n_inc = pt_inc_ptr[P];
  out->n_inc = n_inc;
  inc_point.resize(n_inc);
  inc_tile.resize(n_inc);
  pool.resize_fill(slot_beg, (size_t)n_inc * kSubFrames, -1);
  pool.resize_fill(slot_cnt, (size_t)n_inc * kSubFrames, (unsigned char)0);
  // pass 2: fill them
  std::vector<char> slot_overflow(pool.n, 0);
  pool.run([&](int t) {
    for (int p = pt_cut[t]; p < pt_cut[t + 1]; ++p) {
      if (pt_inc_ptr[p] == pt_inc_ptr[p + 1]) continue;
      const int b = pt_ptr[p], e = pt_ptr[p + 1];
      int last = pt_inc_ptr[p] - 1, last_tile = -1;
      for (int x = b; x < e; ++x) {
        const int f = pt_both[x].y, A = f / kSubFrames, fs = f % kSubFrames;
        if (A != last_tile) {
          ++last;
          last_tile = A;
          inc_point[last] = p;
          inc_tile[last] = A;
        }
        const size_t sl = (size_t)last * kSubFrames + fs;
        if (slot_cnt[sl] == 0) slot_beg[sl] = x;
        if (slot_cnt[sl] == 255) { slot_overflow[t] = 1; continue; }
        slot_cnt[sl]++;
      }
      if (free_cam && e > b) {
        if (last_tile != cam_sub) {
          ++last;
          inc_point[last] = p;
          inc_tile[last] = cam_sub;
        }
        cam_inc[p] = last;
      }
    }
  });
  for (char o : slot_overflow)
    if (o) return fail(error, RSBA_ERR_INVALID_ARGUMENT, "more than 255 observations of one point in one frame");
  lap("incidences");
  // threads own contiguous incidence ranges (cut at point borders: the pair passes below go point by point)
  std::vector<int> inc_cut(pool.n + 1, n_inc);
  for (int t = 0; t <= pool.n; ++t) inc_cut[t] = pt_inc_ptr[pt_cut[t]];
  // where each observation's 12 panel rows live; incidences with a doubly observed frame slot are
  // rebuilt by phi_build_kernel instead.  Which 2-frame halves of an incidence's 4 frame slots are populated
  // (bit 0: slots 0-1, bit 1: slots 2-3): a track that starts or ends inside a sub-tile leaves a half empty; the
  // entries of an off-diagonal pair are grouped by the (column side, row side) half masks so that the SYRK skips
  // the 24 x 24 patches that are structurally zero for a whole work item (k2_schur.cu).
  HostVec<int>& obs_phi_off = out->obs_phi_off;
  std::vector<int>& dup_inc = out->dup_inc;
  std::vector<unsigned char>& inc_half = out->inc_half;
  pool.resize_fill(obs_phi_off, (size_t)std::max<long>(N, 1), -1);
  inc_half.assign(std::max(n_inc, 1), 0);
  dup_inc.clear();
  {
    std::vector<std::vector<int>> dup_of(pool.n);
    pool.run([&](int t) {
      for (int i = inc_cut[t]; i < inc_cut[t + 1]; ++i) {
        const unsigned char* c = &slot_cnt[(size_t)i * kSubFrames];
        const int* sb = &slot_beg[(size_t)i * kSubFrames];
        unsigned m = 0;
        bool dup = false;
        for (int fs = 0; fs < kSubFrames; ++fs) {
          if (c[fs]) m |= 1u << (fs / 2);
          dup = dup || c[fs] > 1;
        }
        inc_half[i] = (unsigned char)m;
        if (dup) { dup_of[t].push_back(i); continue; }
        for (int fs = 0; fs < kSubFrames; ++fs)
          if (c[fs] == 1) obs_phi_off[pt_obs[sb[fs]]] = i * kPanelDoubles + fs * kFrameParams;
      }
      if (free_cam)   // the pseudo-frame rows of a point's panel are written by phi_cam, not through a slot
        for (int p = pt_cut[t]; p < pt_cut[t + 1]; ++p)
          if (cam_inc[p] >= 0) inc_half[cam_inc[p]] |= (unsigned char)(1u << (cam_slot / 2));
    });
    for (const auto& d : dup_of) dup_inc.insert(dup_inc.end(), d.begin(), d.end());   // ascending
  }
  // ---- sub-tile pairs and their entry lists, by a two-pass counting sort on the pair key a*H + b
  // (entries of one pair stay in point order).  The key table is dense while H^2 is small, else the
  // distinct keys are collected and sorted.
  const bool dense_keys = (long)H * H <= (1L << 24) && !sc.sparse_keys;
  std::vector<int> key_pair;          // dense: key -> pair id (or -1)
  std::vector<long> keys;             // sparse: sorted distinct keys
  auto for_each_pair_of_point = [&](int p, auto&& fn) {
    for (int x = pt_inc_ptr[p]; x < pt_inc_ptr[p + 1]; ++x)
      for (int y = x; y < pt_inc_ptr[p + 1]; ++y) fn((long)inc_tile[x] * H + inc_tile[y], x, y);
  };
  std::vector<long> marker_keys;
  for (int t = 0; t < Hreal; ++t) marker_keys.push_back((long)t * H + t);     // every diagonal sub-tile is a pair
  for (const auto& pr : sc.prior_pairs) {                                      // ... and every prior coupling
    const int a = std::min(pr.first, pr.second) / kSubFrames, b = std::max(pr.first, pr.second) / kSubFrames;
    marker_keys.push_back((long)a * H + b);
    if (free_ratio) {   // the ratio (pseudo-frame) couples with both frames of every prior
      marker_keys.push_back((long)a * H + cam_sub);
      marker_keys.push_back((long)b * H + cam_sub);
    }
  }
  std::vector<int>& pair_a = out->pair_a;
  std::vector<int>& pair_b = out->pair_b;
  pair_a.clear(); pair_b.clear();
  if (dense_keys) {
    // which keys occur: one byte per key and thread, OR-ed together
    std::vector<std::vector<char>> seen(pool.n);
    pool.run([&](int t) {
      seen[t].assign((size_t)H * H, 0);
      char* s = seen[t].data();
      for (int p = pt_cut[t]; p < pt_cut[t + 1]; ++p) for_each_pair_of_point(p, [&](long key, int, int) { s[key] = 1; });
    });
    std::vector<char>& present = seen[0];
    for (int t = 1; t < pool.n; ++t)
      for (size_t k = 0; k < (size_t)H * H; ++k) present[k] |= seen[t][k];
    for (long k : marker_keys) present[k] = 1;
    key_pair.assign((size_t)H * H, -1);
    for (long key = 0; key < (long)H * H; ++key)
      if (present[key]) {
        key_pair[key] = (int)pair_a.size();
        pair_a.push_back((int)(key / H));
        pair_b.push_back((int)(key % H));
      }
  } else {
    keys = marker_keys;
    for (int p = 0; p < P; ++p) for_each_pair_of_point(p, [&](long key, int, int) { keys.push_back(key); });
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    for (long key : keys) { pair_a.push_back((int)(key / H)); pair_b.push_back((int)(key % H)); }
  }
  auto pair_of_key = [&](long key) -> int {
    return dense_keys ? key_pair[key] : (int)(std::lower_bound(keys.begin(), keys.end(), key) - keys.begin());
  };
  lap("pairs");
  constexpr int kClasses = 9;   // (mask_a - 1) * 3 + (mask_b - 1); diagonal pairs use class 0 only
  auto class_of = [&](int x, int y) -> int {
    if (x == y) return 0;
    const int ma = inc_half[x] ? inc_half[x] : 3, mb = inc_half[y] ? inc_half[y] : 3;
    return (ma - 1) * 3 + (mb - 1);
  };
  // work items: <= kSchurSegPoints entries of one (pair, class) each, padded to a multiple of 8.
  // Entries per (pair, class) and thread first; a thread's entries of a class go behind those of the threads
  // before it, which keeps every class in point order whatever the thread count.
  const int n_pairs_h = (int)pair_a.size();
  const size_t n_cls = (size_t)n_pairs_h * kClasses;
  std::vector<std::vector<long>> cnt_of(pool.n);
  pool.run([&](int t) {
    cnt_of[t].assign(n_cls + 1, 0);
    long* c = cnt_of[t].data();
    for (int p = pt_cut[t]; p < pt_cut[t + 1]; ++p)
      for_each_pair_of_point(p, [&](long key, int x, int y) { c[(size_t)pair_of_key(key) * kClasses + class_of(x, y)]++; });
  });
  lap("entry counts");
  std::vector<int>& pair_item_ptr = out->pair_item_ptr;
  pair_item_ptr.assign(n_pairs_h + 1, 0);
  std::vector<int4>& items = out->items;
  items.clear();
  long pos = 0;
  for (int q = 0; q < n_pairs_h; ++q) {
    pair_item_ptr[q] = (int)items.size();
    const bool dg = pair_a[q] == pair_b[q];
    for (int cl = 0; cl < kClasses; ++cl) {
      const size_t k = (size_t)q * kClasses + cl;
      long left = 0;
      for (int t = 0; t < pool.n; ++t) { const long c = cnt_of[t][k]; cnt_of[t][k] = pos + left; left += c; }   // -> write cursors
      const int ma = cl / 3 + 1, mb = cl % 3 + 1;
      while (left > 0) {
        const int take = (int)std::min<long>(left, kSchurSegPoints);
        const int padded = (take + 7) / 8 * 8;
        if (pos + padded > 2147483647L) return fail(error, RSBA_ERR_INVALID_ARGUMENT, "Schur entry list exceeds 2^31");
        // w: bit 0 = diagonal pair; bits 4-5 / 8-9 = populated halves of the column (A) / row (B) side
        items.push_back(make_int4(q, (int)pos, padded, dg ? 1 : ((ma << 4) | (mb << 8))));
        pos += padded;
        left -= take;
      }
    }
  }
  pair_item_ptr[n_pairs_h] = (int)items.size();
  HostVec<int2>& entries = out->entries;
  lap("work items");
  pool.resize_fill(entries, (size_t)std::max<long>(pos, 1), make_int2(n_inc, n_inc));   // zero panel
  // (kSchurSegPoints is a multiple of 8: only the last segment of a class is padded, so a class is contiguous)
  pool.run([&](int t) {
    long* cur = cnt_of[t].data();
    for (int p = pt_cut[t]; p < pt_cut[t + 1]; ++p)
      for_each_pair_of_point(p, [&](long key, int x, int y) {
        entries[cur[(size_t)pair_of_key(key) * kClasses + class_of(x, y)]++] = make_int2(y, x);   // (row side B, column side A)
      });
  });
  out->n_items = (int)pair_item_ptr.back();
  if (items.empty()) items.push_back(make_int4(0, 0, 0, 1));
  lap("entry lists");
  // ---- ordering, symbolic factorisation, elimination levels (tile_plan.cu)
  // (the plan is a function of the WHOLE scene, so every rank of a multi-GPU run derives the same one)
  TilePlan& plan = out->plan;
  {
    std::vector<std::pair<int, int>> tp;
    if (sc.world > 1) {
      const long NG = sc.n_obs_global;
      std::vector<long> gptr(P + 1, 0);
      for (long i = 0; i < NG; ++i) gptr[sc.g_obs_point[i] + 1]++;
      for (int p = 0; p < P; ++p) gptr[p + 1] += gptr[p];
      std::vector<int> gtile(NG);
      {
        std::vector<long> cur(gptr.begin(), gptr.end() - 1);
        for (long i = 0; i < NG; ++i) gtile[cur[sc.g_obs_point[i]]++] = sc.g_obs_frame[i] / kFramesPerTile;
      }
      std::vector<char> seen((size_t)T * T, 0);
      std::vector<int> tiles;
      for (int p = 0; p < P; ++p) {
        if (sc.point_const[p]) continue;
        tiles.clear();
        for (long e = gptr[p]; e < gptr[p + 1]; ++e)
          if (tiles.empty() || tiles.back() != gtile[e]) tiles.push_back(gtile[e]);
        for (size_t x = 0; x < tiles.size(); ++x)
          for (size_t y = x; y < tiles.size(); ++y) seen[(size_t)tiles[x] * T + tiles[y]] = 1;
      }
      for (const auto& pr : sc.prior_pairs) {
        const int a = std::min(pr.first, pr.second) / kFramesPerTile, b = std::max(pr.first, pr.second) / kFramesPerTile;
        seen[(size_t)a * T + b] = 1;
        if (free_ratio) seen[(size_t)a * T + F / kFramesPerTile] = seen[(size_t)b * T + F / kFramesPerTile] = 1;
      }
      for (int a = 0; a < T; ++a)
        for (int b = a; b < T; ++b)
          if (seen[(size_t)a * T + b]) tp.emplace_back(a, b);
    } else {
      tp.reserve(pair_a.size());
      for (size_t k = 0; k < pair_a.size(); ++k) tp.emplace_back(pair_a[k] / 2, pair_b[k] / 2);
      std::sort(tp.begin(), tp.end());
      tp.erase(std::unique(tp.begin(), tp.end()), tp.end());
    }
    const int border_tile = pseudo ? F / kFramesPerTile : -1;   // couples with everything: eliminated last
    if (free_cam && sc.world > 1)
      for (int a = 0; a < T; ++a) tp.emplace_back(std::min(a, border_tile), std::max(a, border_tile));
    build_tile_plan(T, tp, sc.dense, sc.reorder, border_tile, &plan);
  }
  // where each off-diagonal tile (i, k) leaves its forward-substitution term: its index in row i's list
  out->fwd_slot.assign(std::max<size_t>(plan.trsm.size(), 1), 0);
  for (size_t t = 0; t < plan.trsm.size(); ++t) {
    const int i = plan.trsm[t].x, k = plan.trsm[t].y;
    int q = plan.lrow_ptr[i];
    while (q < plan.lrow_ptr[i + 1] && plan.lrow_cols[q] != k) ++q;
    if (q == plan.lrow_ptr[i + 1]) return fail(error, RSBA_ERR_STATE, "tile plan: trsm tile missing from the row list");
    out->fwd_slot[t] = q;
  }
  lap("tile plan (ND + symbolic)");
  return RSBA_OK;
}

}  // namespace rsba

// ------------------------------------------------------------------ device-free C ABI (include/rsba_cuda.h)
struct rsba_structure {
  rsba::HostStructure hs;
  // multi-GPU shard (world_size > 1): the rank's observations (positions in the caller's sorted list)
  std::vector<long> local_ids;
  std::vector<int> local_frame, local_point;
  std::vector<unsigned char> point_owned;
};

extern "C" {

int rsba_cuda_analyze_structure(long n_obs, const int* obs_frame, const int* obs_point, int n_frames, int n_points,
                                const unsigned char* const_point, int free_intrinsics, int free_ratio, int n_priors,
                                const int* prior_frame, const int* prior_prev, int dense, int reorder,
                                int sparse_keys, int rank, int world_size, rsba_structure** out) {
  return rsba::api_guard([&]() -> int {
  using namespace rsba;
  auto bad = [](const char* msg) { set_last_error(msg); return (int)RSBA_ERR_INVALID_ARGUMENT; };
  if (!out) return bad("out is NULL");
  *out = nullptr;
  if (n_obs < 0 || n_frames < 0 || n_points < 0 || n_priors < 0 || (n_obs > 0 && (!obs_frame || !obs_point)) ||
      (n_priors > 0 && (!prior_frame || !prior_prev)) || world_size < 1 || rank < 0 || rank >= world_size)
    return bad("bad arguments");
  for (long i = 0; i < n_obs; ++i) {
    if (obs_frame[i] < 0 || obs_frame[i] >= n_frames || obs_point[i] < 0 || obs_point[i] >= n_points)
      return bad("observation index out of range");
    if (i > 0 && obs_frame[i - 1] > obs_frame[i]) return bad("obs_frame must be sorted");
  }
  SceneTopology sc;
  sc.n_obs = n_obs; sc.n_frames = n_frames; sc.n_points = n_points;
  sc.obs_frame = obs_frame; sc.obs_point = obs_point;
  std::vector<unsigned char> cp(std::max(n_points, 1), 0);
  if (const_point) std::copy(const_point, const_point + n_points, cp.begin());
  sc.point_const = cp.data();
  sc.free_cam = free_intrinsics != 0; sc.free_ratio = free_ratio != 0;
  for (int i = 0; i < n_priors; ++i) {
    if (prior_frame[i] < 0 || prior_frame[i] >= n_frames || prior_prev[i] < 0 || prior_prev[i] >= n_frames)
      return bad("motion prior: frame index out of range");
    sc.prior_pairs.emplace_back(prior_frame[i], prior_prev[i]);
  }
  sc.n_obs_global = n_obs; sc.g_obs_frame = obs_frame; sc.g_obs_point = obs_point;
  sc.dense = dense != 0; sc.reorder = reorder != 0; sc.sparse_keys = sparse_keys != 0;
  rsba_structure* s = new rsba_structure;
  if (world_size > 1) {
    // this rank's share as materialize_local_share (problem.cu) forms it: all observations of the points it owns
    std::vector<int> owner;
    compute_point_owners(n_frames, n_points, n_obs, obs_frame, obs_point, world_size, &owner);
    s->point_owned.resize(n_points);
    for (int p = 0; p < n_points; ++p) s->point_owned[p] = owner[p] == rank;
    for (long i = 0; i < n_obs; ++i)
      if (s->point_owned[obs_point[i]]) {
        s->local_ids.push_back(i);
        s->local_frame.push_back(obs_frame[i]);
        s->local_point.push_back(obs_point[i]);
      }
    sc.world = world_size;
    sc.n_obs = (long)s->local_ids.size();
    sc.obs_frame = s->local_frame.data();
    sc.obs_point = s->local_point.data();
  } else {
    s->point_owned.assign(n_points, 1);
  }
  std::string err;
  // RSBA_CUDA_TRACE=1: wall-clock of each phase, to stderr (as rsba_cuda_solve prints it)
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [](const char* what, void* ctx) {
    auto* prev = static_cast<std::chrono::steady_clock::time_point*>(ctx);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[rsba_cuda] structure: %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - *prev).count());
    *prev = now;
  };
  const bool trace = getenv("RSBA_CUDA_TRACE") != nullptr;
  const int rc = analyze_structure(sc, &s->hs, &err, trace ? +lap : nullptr, &t_prev);
  if (rc) {
    delete s;
    set_last_error(err);
    return rc;
  }
  *out = s;
  return RSBA_OK;
  });
}

long rsba_cuda_structure_array(const rsba_structure* s, const char* name, const void** data, int* elem_bytes) {
  if (!s || !name) return -1;
  const rsba::HostStructure& h = s->hs;
  const std::string n(name);
  auto give = [&](const auto& v) -> long {
    if (data) *data = v.data();
    if (elem_bytes) *elem_bytes = (int)sizeof(v[0]);
    return (long)v.size();
  };
#define RSBA_ARR(field) if (n == #field) return give(h.field)
  RSBA_ARR(pt_ptr); RSBA_ARR(pt_obs); RSBA_ARR(chunk_frame); RSBA_ARR(chunk_beg); RSBA_ARR(chunk_cnt);
  RSBA_ARR(frame_chunk_ptr); RSBA_ARR(inc_point); RSBA_ARR(inc_tile); RSBA_ARR(slot_beg); RSBA_ARR(slot_cnt);
  RSBA_ARR(pt_inc_ptr); RSBA_ARR(cam_inc); RSBA_ARR(inc_half); RSBA_ARR(obs_phi_off); RSBA_ARR(dup_inc);
  RSBA_ARR(pair_a); RSBA_ARR(pair_b); RSBA_ARR(pair_item_ptr); RSBA_ARR(items); RSBA_ARR(entries); RSBA_ARR(fwd_slot);
  RSBA_ARR(point_groups); RSBA_ARR(point_big);
#undef RSBA_ARR
#define RSBA_PLAN(field) if (n == "plan." #field) return give(h.plan.field)
  RSBA_PLAN(tile_pos); RSBA_PLAN(pos_tile); RSBA_PLAN(nz_tiles); RSBA_PLAN(tile_slot); RSBA_PLAN(panels);
  RSBA_PLAN(panel_ptr); RSBA_PLAN(trsm); RSBA_PLAN(trsm_ptr); RSBA_PLAN(upd); RSBA_PLAN(lrow_ptr); RSBA_PLAN(lrow_cols);
#undef RSBA_PLAN
  if (n == "local_ids") return give(s->local_ids);
  if (n == "point_owned") return give(s->point_owned);
  if (n == "T" || n == "n_inc" || n == "n_items") {   // scalars come back as the count
    if (data) *data = nullptr;
    if (elem_bytes) *elem_bytes = 0;
    return n == "T" ? h.T : n == "n_inc" ? h.n_inc : h.n_items;
  }
  rsba::set_last_error("rsba_cuda_structure_array: unknown array name");
  return -1;
}

void rsba_cuda_structure_free(rsba_structure* s) { delete s; }

}  // extern "C"
