// Host-side symbolic analysis of the reduced camera system (see tile_plan.cuh).
#include "tile_plan.cuh"

#include <algorithm>
#include <cstdint>
#include <initializer_list>
#include <queue>

namespace rsba {
namespace {

constexpr int kLeafTiles = 8;   // sub-graphs of at most this many tiles are ordered naturally

using Adj = std::vector<std::vector<int>>;

// BFS inside the sub-graph `in` (in[v] == tag) from `start`; returns the level sets.
std::vector<std::vector<int>> bfs_levels(const Adj& adj, const std::vector<int>& in, int tag, int start,
                                         std::vector<int>& seen, int stamp) {
  std::vector<std::vector<int>> levels(1, std::vector<int>(1, start));
  seen[start] = stamp;
  while (true) {
    std::vector<int> next;
    for (int v : levels.back())
      for (int w : adj[v])
        if (in[w] == tag && seen[w] != stamp) {
          seen[w] = stamp;
          next.push_back(w);
        }
    if (next.empty()) break;
    levels.push_back(std::move(next));
  }
  return levels;
}

struct Dissector {
  const Adj& adj;
  std::vector<int> in, seen;
  int next_tag = 1, stamp = 0;
  std::vector<int> order;
  explicit Dissector(const Adj& a) : adj(a), in(a.size(), 0), seen(a.size(), 0) {}

  void natural(std::vector<int>& nodes) {
    std::sort(nodes.begin(), nodes.end());
    order.insert(order.end(), nodes.begin(), nodes.end());
  }

  void run(std::vector<int> nodes) {
    if ((int)nodes.size() <= kLeafTiles) return natural(nodes);
    const int tag = next_tag++;
    for (int v : nodes) in[v] = tag;
    // connected components are independent sub-trees
    std::sort(nodes.begin(), nodes.end());
    {
      std::vector<std::vector<int>> comps;
      const int st = ++stamp;
      for (int v : nodes) {
        if (seen[v] == st) continue;
        std::vector<int> comp;
        for (auto& lv : bfs_levels(adj, in, tag, v, seen, st)) comp.insert(comp.end(), lv.begin(), lv.end());
        comps.push_back(std::move(comp));
      }
      if (comps.size() > 1) {
        for (auto& c : comps) run(c);
        return;
      }
    }
    // level structure rooted at a pseudo-peripheral node
    int root = nodes.front();
    std::vector<std::vector<int>> levels;
    for (int pass = 0; pass < 3; ++pass) {
      for (int v : nodes) in[v] = tag;   // (recursion below re-tags; restore before every BFS)
      levels = bfs_levels(adj, in, tag, root, seen, ++stamp);
      const std::vector<int>& last = levels.back();
      root = *std::min_element(last.begin(), last.end());
    }
    const int m = (int)levels.size();
    if (m < 3) return natural(nodes);    // too dense to dissect
    std::vector<long> prefix(m + 1, 0);
    for (int l = 0; l < m; ++l) prefix[l + 1] = prefix[l] + (long)levels[l].size();
    int best = -1;
    long best_cost = -1;
    for (int s = 1; s + 1 < m; ++s) {
      const long left = prefix[s], right = prefix[m] - prefix[s + 1];
      const long cost = std::max(left, right) * 4 + (long)levels[s].size();
      if (best < 0 || cost < best_cost) { best = s; best_cost = cost; }
    }
    if ((long)levels[best].size() * 2 >= (long)nodes.size()) return natural(nodes);
    std::vector<int> left, right, sep = levels[best];
    for (int l = 0; l < best; ++l) left.insert(left.end(), levels[l].begin(), levels[l].end());
    for (int l = best + 1; l < m; ++l) right.insert(right.end(), levels[l].begin(), levels[l].end());
    run(left);
    run(right);
    natural(sep);
  }
};

}  // namespace

void build_tile_plan(int T, const std::vector<std::pair<int, int>>& tile_pairs, bool dense, bool reorder,
                     int border_tile, TilePlan* plan) {
  TilePlan& P = *plan;
  P = TilePlan();
  P.T = T;
  P.tile_pos.assign(std::max(T, 1), 0);
  P.pos_tile.assign(std::max(T, 1), 0);
  for (int t = 0; t < T; ++t) P.tile_pos[t] = P.pos_tile[t] = t;
  if (reorder && !dense && T > kLeafTiles) {
    Adj adj(T);
    for (auto& pr : tile_pairs)
      if (pr.first != pr.second && pr.first != border_tile && pr.second != border_tile) {
        adj[pr.first].push_back(pr.second);
        adj[pr.second].push_back(pr.first);
      }
    for (auto& a : adj) {
      std::sort(a.begin(), a.end());
      a.erase(std::unique(a.begin(), a.end()), a.end());
    }
    Dissector d(adj);
    std::vector<int> all;
    for (int t = 0; t < T; ++t)
      if (t != border_tile) all.push_back(t);
    d.run(all);
    if (border_tile >= 0 && border_tile < T) d.order.push_back(border_tile);
    for (int pos = 0; pos < T; ++pos) {
      P.pos_tile[pos] = d.order[pos];
      P.tile_pos[d.order[pos]] = pos;
    }
  }
  // ---- symbolic factorisation on the permuted tile graph
  std::vector<char> nz((size_t)T * T, 0);
  for (int i = 0; i < T; ++i) nz[(size_t)i * T + i] = 1;
  if (dense) {
    for (int i = 0; i < T; ++i)
      for (int j = 0; j <= i; ++j) nz[(size_t)i * T + j] = 1;
  } else {
    for (auto& pr : tile_pairs) {
      const int pa = P.tile_pos[pr.first], pb = P.tile_pos[pr.second];
      nz[(size_t)std::max(pa, pb) * T + std::min(pa, pb)] = 1;
    }
  }
  P.row_ptr.assign(T + 1, 0);
  std::vector<int> level(std::max(T, 1), 0);
  for (int k = 0; k < T; ++k) {
    P.row_ptr[k] = (int)P.rows.size();
    const size_t r0 = P.rows.size();
    for (int i = k + 1; i < T; ++i)
      if (nz[(size_t)i * T + k]) P.rows.push_back(i);
    for (size_t x = r0; x < P.rows.size(); ++x) {
      level[P.rows[x]] = std::max(level[P.rows[x]], level[k] + 1);   // panel rows[x] waits for panel k
      for (size_t y = r0; y <= x; ++y) nz[(size_t)P.rows[x] * T + P.rows[y]] = 1;  // fill
    }
  }
  P.row_ptr[T] = (int)P.rows.size();
  P.tile_slot.assign((size_t)T * T, -1);
  P.lrow_ptr.assign(T + 1, 0);
  for (int i = 0; i < T; ++i) {
    P.lrow_ptr[i] = (int)P.lrow_cols.size();
    for (int j = 0; j <= i; ++j)
      if (nz[(size_t)i * T + j]) {
        P.tile_slot[(size_t)i * T + j] = (int)P.nz_tiles.size();
        P.nz_tiles.push_back(make_int2(i, j));
        if (j < i) P.lrow_cols.push_back(j);
      }
  }
  P.lrow_ptr[T] = (int)P.lrow_cols.size();
  // ---- levels
  P.n_levels = 0;
  for (int k = 0; k < T; ++k) P.n_levels = std::max(P.n_levels, level[k] + 1);
  P.panel_ptr.assign(P.n_levels + 1, 0);
  for (int k = 0; k < T; ++k) P.panel_ptr[level[k] + 1]++;
  for (int l = 0; l < P.n_levels; ++l) P.panel_ptr[l + 1] += P.panel_ptr[l];
  P.panels.assign(std::max(T, 1), 0);
  {
    std::vector<int> cur(P.panel_ptr.begin(), P.panel_ptr.end() - (P.n_levels ? 1 : 0));
    for (int k = 0; k < T; ++k) P.panels[cur[level[k]]++] = k;
  }
  P.trsm_ptr.assign(P.n_levels + 1, 0);
  P.level_group_ptr.assign(P.n_levels + 1, 0);
  P.group_ptr.assign(1, 0);
  const double t3 = 96.0 * 96.0 * 96.0;
  std::vector<int> slot_level((size_t)P.nz_tiles.size(), -1);
  std::vector<uint64_t> slot_colors((size_t)P.nz_tiles.size(), 0);
  for (int l = 0; l < P.n_levels; ++l) {
    P.trsm_ptr[l] = (int)P.trsm.size();
    P.level_group_ptr[l] = (int)P.group_ptr.size() - 1;
    std::vector<std::vector<int>> by_color;   // panels per colour
    for (int q = P.panel_ptr[l]; q < P.panel_ptr[l + 1]; ++q) {
      const int k = P.panels[q];
      P.flops += 2.0 * t3 / 3.0;
      const int rb = P.row_ptr[k], re = P.row_ptr[k + 1];
      for (int x = rb; x < re; ++x) {
        P.trsm.push_back(make_int2(P.rows[x], k));
        P.flops += t3;
      }
      if (re == rb) continue;
      uint64_t used = 0;
      for (int x = rb; x < re; ++x)
        for (int y = rb; y <= x; ++y) {
          const int sl = P.tile_slot[(size_t)P.rows[x] * T + P.rows[y]];
          if (slot_level[sl] == l) used |= slot_colors[sl];
        }
      int color = 0;
      while (color < 63 && ((used >> color) & 1)) ++color;
      if (color == 63) color = (int)std::max<size_t>(by_color.size(), 63);   // own group: always conflict-free
      for (int x = rb; x < re; ++x)
        for (int y = rb; y <= x; ++y) {
          const int sl = P.tile_slot[(size_t)P.rows[x] * T + P.rows[y]];
          if (slot_level[sl] != l) { slot_level[sl] = l; slot_colors[sl] = 0; }
          if (color < 63) slot_colors[sl] |= (uint64_t)1 << color;
        }
      if ((int)by_color.size() <= color) by_color.resize(color + 1);
      by_color[color].push_back(k);
    }
    for (auto& group : by_color) {
      if (group.empty()) continue;
      for (int k : group) {
        const int rb = P.row_ptr[k], re = P.row_ptr[k + 1];
        for (int x = rb; x < re; ++x)
          for (int y = rb; y <= x; ++y) {
            P.upd.push_back(make_int4(P.rows[x], P.rows[y], k, 0));
            P.flops += 2.0 * t3;
          }
      }
      P.group_ptr.push_back((long)P.upd.size());
    }
  }
  P.trsm_ptr[P.n_levels] = (int)P.trsm.size();
  P.level_group_ptr[P.n_levels] = (int)P.group_ptr.size() - 1;
}

// ---------------------------------------------------------------- task graph of the numeric phase
void build_dag_plan(const TilePlan& P, int merge_levels, DagPlan* dag) {
  DagPlan& D = *dag;
  D = DagPlan();
  const int T = P.T;
  const int L = P.n_levels;
  if (merge_levels < 1) merge_levels = 1;
  std::vector<int> level(std::max(T, 1), 0);
  for (int l = 0; l < L; ++l)
    for (int q = P.panel_ptr[l]; q < P.panel_ptr[l + 1]; ++q) level[P.panels[q]] = l;
  auto slot = [&](int i, int j) { return P.tile_slot[(size_t)i * T + j]; };
  D.need.assign(std::max<size_t>(P.nz_tiles.size(), 1) * 4, 0);

  struct Keyed { long key[7]; DagTask t; };
  std::vector<Keyed> all;
  auto push = [&](std::initializer_list<long> key, DagTask t) {
    Keyed k{};
    int n = 0;
    for (long v : key) k.key[n++] = v;
    k.t = t;
    all.push_back(k);
  };
  // ---- factorisation: one FACTOR per panel, kTrsmParts TRSM slabs per tile below it
  for (int k = 0; k < T; ++k) {
    push({level[k], 0, k}, DagTask{kTaskFactor, k, slot(k, k), 0, 0, 0, 0, 0});
    for (int x = P.row_ptr[k]; x < P.row_ptr[k + 1]; ++x) {
      const int i = P.rows[x];
      int fwd = -1;
      for (int q = P.lrow_ptr[i]; q < P.lrow_ptr[i + 1]; ++q)
        if (P.lrow_cols[q] == k) fwd = q;
      for (int part = 0; part < kTrsmParts; ++part)
        push({level[k], 2, level[i], i, k, part}, DagTask{kTaskTrsm, i, k, part, slot(i, k), slot(k, k), fwd, 0});
    }
  }
  // ---- trailing updates, gathered per target tile; sources in (level, panel) order
  {
    std::vector<std::vector<int>> by_target(P.nz_tiles.size());   // source panels k of each target slot
    for (const int4& u : P.upd) by_target[slot(u.x, u.y)].push_back(u.z);
    for (size_t s = 0; s < by_target.size(); ++s) {
      std::vector<int>& ks = by_target[s];
      if (ks.empty()) continue;
      const int i = P.nz_tiles[s].x, j = P.nz_tiles[s].y;
      std::sort(ks.begin(), ks.end(), [&](int a, int b) { return level[a] != level[b] ? level[a] < level[b] : a < b; });
      const int lj = level[j];
      // group id: the sources right below the target's column stay on their own (they are on the critical
      // path); older ones are taken merge_levels levels at a time
      auto group_of = [&](int k) { return level[k] >= lj - 1 ? (long)1 << 30 : (long)(level[k] / merge_levels); };
      int order = 0;
      for (size_t b = 0; b < ks.size();) {
        size_t e = b;
        while (e < ks.size() && group_of(ks[e]) == group_of(ks[b])) ++e;
        const int first = (int)D.sources.size();
        int lmax = 0;
        for (size_t x = b; x < e; ++x) {
          D.sources.push_back(make_int2(slot(i, ks[x]), slot(j, ks[x])));
          lmax = std::max(lmax, level[ks[x]]);
        }
        for (int q = 0; q < 4; ++q) {
          if (i == j && q == 1) continue;          // diagonal targets: the upper-right quadrant is never read
          const DagTask t{kTaskUpdate, (int)s, q, order, first, (int)(e - b), 0, 0};
          if (lmax == lj - 1) push({lmax, 3, i != j, level[i], i, j, q}, t);
          else push({lmax + 1, 1, lj, i != j, i, j, q}, t);
          D.need[s * 4 + q] = order + 1;
        }
        ++order;
        b = e;
      }
    }
  }
  // ---- backward substitution: y_k = L_kk^-T (z_k - sum_{i > k} L_ik^T y_i), highest level first
  for (int k = 0; k < T; ++k) {
    push({L + (L - 1 - level[k]), 0, k},
         DagTask{kTaskBackFin, k, slot(k, k), P.row_ptr[k], P.row_ptr[k + 1] - P.row_ptr[k], 0, 0, 0});
    for (int x = P.row_ptr[k]; x < P.row_ptr[k + 1]; ++x) {
      const int i = P.rows[x];
      push({L + (L - 1 - level[i]), 1, -level[k], i, k}, DagTask{kTaskBackTile, i, k, slot(i, k), x, 0, 0, 0});
    }
  }
  std::stable_sort(all.begin(), all.end(), [](const Keyed& a, const Keyed& b) {
    for (int n = 0; n < 7; ++n)
      if (a.key[n] != b.key[n]) return a.key[n] < b.key[n];
    return false;
  });
  D.tasks.reserve(all.size());
  for (const Keyed& k : all) {
    D.tasks.push_back(k.t);
    if (k.t.type <= kTaskUpdate) D.n_factor_tasks = (int)D.tasks.size();
  }
}

}  // namespace rsba
