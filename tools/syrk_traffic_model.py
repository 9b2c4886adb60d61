"""Host-only model of the Schur SYRK's work at a given config, from the structure analysis alone (no GPU):
panel loads (L2 -> shared-memory fill), DMMA tiles executed and algorithmic flops under the CURRENT work lists, and
the panel loads of two candidate regroupings.  Usage: python tools/syrk_traffic_model.py [C3]"""
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rsba_b200.api as api                      # noqa: E402
from rsba_b200.scene import make_config          # noqa: E402

PANEL_BYTES = 3 * 52 * 8


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    sc = make_config(name)
    st = api.analyze_structure(sc.obs_frame, sc.obs_point, sc.num_frames, sc.num_points)
    items, n_inc = st["items"], st["n_inc"]
    diag = (items[:, 3] & 1) == 1
    ma, mb = (items[:, 3] >> 4) & 3, (items[:, 3] >> 8) & 3
    pop = np.array([0, 1, 1, 2])
    ent = items[:, 2].astype(np.int64)                         # padded entries per item
    loads = int(ent[diag].sum() + 2 * ent[~diag].sum())
    live_warps = np.where(diag, 0, pop[ma] * pop[mb])
    tiles = np.where(diag, 21, 9 * live_warps).astype(np.int64)
    dmma = int((tiles * ent * 3 // 4).sum())                   # 8x8x4 DMMAs: K = 3 per entry, 4 per instruction
    flops_exec = dmma * 512
    k = np.diff(st["pt_ptr"]).astype(np.int64)
    flops_alg = int(((12 * k) * (12 * k + 1) // 2 * 3 * 2).sum())
    m = np.diff(st["pt_inc_ptr"]).astype(np.int64)             # incidences (sub-tiles) per point
    real_entries = int((m * (m + 1) // 2).sum())
    print(f"{name}: {n_inc} incidences, {items.shape[0]} work items, {real_entries} entries (+{int(ent.sum()) - real_entries} padding)")
    print(f"  current lists : {loads / 1e6:8.2f} M panel loads = {loads * PANEL_BYTES / 1e9:6.2f} GB fill per launch; "
          f"{dmma / 1e6:8.1f} M DMMAs = {flops_exec / 1e9:6.1f} GF executed, {flops_alg / 1e9:6.1f} GF algorithmic "
          f"({flops_alg / flops_exec:.2f})")
    # (a; b, b+1): one CTA computes two row sub-tiles against one column sub-tile from three panels
    row_pairs = int(sum((np.ceil((mm - np.arange(mm)) / 2) + (mm - np.arange(mm))).sum() for mm in np.unique(m) if mm > 0
                        for _ in range(int((m == mm).sum()))) if np.unique(m).size < 64 else 0)
    if row_pairs:
        print(f"  (a; b, b+1)   : {row_pairs / 1e6:8.2f} M panel loads = {row_pairs * PANEL_BYTES / 1e9:6.2f} GB "
              f"({row_pairs / max(1, int((m * m).sum())):.2f} of the unpadded current {int((m * m).sum()) / 1e6:.2f} M)")
    # 2 x 2: one CTA per Cholesky-tile pair of a point (96 x 96), four panels (two on the diagonal)
    fr, pt_obs, ptr = np.asarray(sc.obs_frame), st["pt_obs"], st["pt_ptr"]
    tile_of_obs = fr[pt_obs] // 8
    first = np.ones(tile_of_obs.size, bool)
    first[1:] = tile_of_obs[1:] != tile_of_obs[:-1]
    first[ptr[:-1][ptr[:-1] < tile_of_obs.size]] = True
    tiles_per_point = np.add.reduceat(first.astype(np.int64), ptr[:-1][k > 0]) if (k > 0).any() else np.zeros(0)
    t = tiles_per_point
    loads22 = int((2 * t + 4 * (t * (t - 1) // 2)).sum())
    blocks22 = int((3 * t + 4 * (t * (t - 1) // 2)).sum())      # 48 x 48 blocks computed (lower part on the diagonal)
    print(f"  2 x 2 tiles   : {loads22 / 1e6:8.2f} M panel loads = {loads22 * PANEL_BYTES / 1e9:6.2f} GB, but "
          f"{blocks22 / 1e6:.2f} M 48x48 blocks instead of {real_entries / 1e6:.2f} M ({blocks22 / real_entries:.2f}x the DMMAs "
          f"unless absent sub-tiles are masked per class)")


if __name__ == "__main__":
    main()
