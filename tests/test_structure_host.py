"""CPU suite: the one-off host structure analysis of rsba_cuda_solve (rsba_b200/csrc/structure.cu), run through the
device-free C ABI `rsba_cuda_analyze_structure`.  It is the analogue of what Ceres does before its first iteration
(Program reordering, SchurEliminator block-structure detection, CHOLMOD analyse; reached through ceres::Solve,
CeresHandler.h:403,419).  The work lists the Schur SYRK kernel consumes are EXECUTED here in numpy and compared with
a direct per-point Schur complement, so a wrong pair, a dropped entry or a wrong half mask shows up without a GPU."""
import numpy as np
import pytest

import rsba_b200.api as api
from helpers import small_scene

SUB, FP, SEG = 4, 12, 512          # frames per sub-tile, parameters per frame, max entries per work item
PANEL_DOUBLES = 3 * (SUB * FP + 4)


def ragged_topology(seed, F, P, const_frac=0.1, dup_frac=0.05, max_track=9):
    """Frame-sorted (frame, point) lists with ragged tracks, unobserved points, constant points and a few
    points seen twice in one frame."""
    rng = np.random.default_rng(seed)
    fr, pt = [], []
    for p in range(P):
        if rng.random() < 0.05:
            continue                                  # never observed
        k = int(rng.integers(1, max_track + 1))
        start = int(rng.integers(0, F))
        frames = sorted(set(int(f) for f in rng.integers(start, min(F, start + 2 * max_track), size=k)))
        for f in frames:
            fr.append(f); pt.append(p)
            if rng.random() < dup_frac:
                fr.append(f); pt.append(p)            # duplicate observation in the same frame
    fr, pt = np.array(fr, np.int32), np.array(pt, np.int32)
    order = np.argsort(fr, kind="stable")
    const_point = (rng.random(P) < const_frac).astype(np.uint8)
    return fr[order], pt[order], const_point


def execute_work_lists(st, Phi):
    """The Schur work lists executed the way schur_syrk_kernel + schur_reduce do: {(a, b): 48 x 48 block}, summing
    the items of each pair; patches the kernel skips (half masks) are skipped here too.  Also checks the lists' own
    invariants.  Returns (blocks, number of non-padding entries, total entries)."""
    n_inc = st["n_inc"]
    ip, it = st["inc_point"], st["inc_tile"]
    pa, pb, pitem, items, entries = st["pair_a"], st["pair_b"], st["pair_item_ptr"], st["items"], st["entries"]
    H = 2 * st["T"]
    assert np.all(pa <= pb) and np.all(pb < H)
    keys = pa.astype(np.int64) * H + pb
    assert np.all(np.diff(keys) > 0), "pairs sorted by (a, b), no repeats"
    assert pitem.size == pa.size + 1 and pitem[-1] == st["n_items"]
    blocks = {}
    n_real_entries = 0
    pos = 0
    for q in range(pa.size):
        acc = np.zeros((SUB * FP, SUB * FP))
        diag = pa[q] == pb[q]
        for j in range(pitem[q], pitem[q + 1]):
            iq, beg, cnt, w = (int(v) for v in items[j])
            assert iq == q and beg == pos and cnt % 8 == 0 and 0 < cnt <= SEG
            pos += cnt
            assert (w & 1) == int(diag)
            ma, mb = ((w >> 4) & 3, (w >> 8) & 3) if not diag else (3, 3)
            rows_b = np.concatenate([np.arange(24) + 24 * hb for hb in range(2) if mb >> hb & 1])
            rows_a = np.concatenate([np.arange(24) + 24 * ha for ha in range(2) if ma >> ha & 1])
            for e in range(beg, beg + cnt):
                y, x = (int(v) for v in entries[e])
                if y == n_inc:
                    assert x == n_inc                  # padding: the all-zero panel
                    continue
                n_real_entries += 1
                assert it[x] == pa[q] and it[y] == pb[q] and ip[x] == ip[y]
                # the kernel only computes the patches of populated halves: what it skips must be zero
                part = np.zeros_like(acc)
                part[np.ix_(rows_b, rows_a)] = Phi[y][rows_b] @ Phi[x][rows_a].T
                acc += part
        blocks[(int(pa[q]), int(pb[q]))] = acc
    return blocks, n_real_entries, pos


def check_structure(fr, pt, F, P, const_point=None, free_cam=False, free_ratio=False, priors=(), seed=0, **kw):
    const_point = np.zeros(P, np.uint8) if const_point is None else const_point
    pf = [a for a, _ in priors]
    pp = [b for _, b in priors]
    st = api.analyze_structure(fr, pt, F, P, const_point, free_cam, free_ratio, pf, pp, **kw)
    N = fr.size
    pseudo = free_cam or free_ratio
    Fc = F + (1 if pseudo else 0)
    T = (12 * Fc + 95) // 96
    assert st["T"] == T

    # ---- point-major CSR = stable sort of the observations by point
    want_obs = np.argsort(pt, kind="stable").astype(np.int32)
    assert np.array_equal(st["pt_obs"], want_obs)
    assert np.array_equal(st["pt_ptr"], np.concatenate([[0], np.cumsum(np.bincount(pt, minlength=P))]))

    # ---- frame chunks: each frame's run of observations cut into pieces of <= 128, in order
    cf, cb, cc, fcp = st["chunk_frame"], st["chunk_beg"], st["chunk_cnt"], st["frame_chunk_ptr"]
    assert fcp.size == Fc + 1 and fcp[-1] == cf.size
    assert np.all(cc >= 1) and np.all(cc <= 128)
    covered = np.concatenate([np.arange(b, b + c) for b, c in zip(cb, cc)]) if cf.size else np.zeros(0, int)
    assert np.array_equal(covered, np.arange(N))
    assert np.array_equal(np.repeat(cf, cc), fr)
    for f in range(F):
        assert np.all(cf[fcp[f]:fcp[f + 1]] == f)

    # ---- incidences: one per (sub-tile, eliminated point) with an observation (+ the pseudo-frame's sub-tile)
    n_inc = st["n_inc"]
    ip, it, sb, sc_ = st["inc_point"], st["inc_tile"], st["slot_beg"], st["slot_cnt"]
    pip, cam_inc = st["pt_inc_ptr"], st["cam_inc"]
    assert ip.size == n_inc == it.size and sb.shape == (n_inc, SUB) and sc_.shape == (n_inc, SUB)
    cam_sub, cam_slot = F // SUB, F % SUB
    for p in range(P):
        obs = want_obs[st["pt_ptr"][p]:st["pt_ptr"][p + 1]]
        subs = sorted(set(int(fr[o]) // SUB for o in obs)) if not const_point[p] else []
        if free_cam and subs and subs[-1] != cam_sub:
            subs.append(cam_sub)
        got = it[pip[p]:pip[p + 1]]
        assert list(got) == subs, (p, got, subs)
        assert np.all(ip[pip[p]:pip[p + 1]] == p)
        if free_cam and subs:
            assert it[cam_inc[p]] == cam_sub and ip[cam_inc[p]] == p
        else:
            assert cam_inc[p] == -1
        for i in range(pip[p], pip[p + 1]):
            for fs in range(SUB):
                mine = [x for x in range(st["pt_ptr"][p], st["pt_ptr"][p + 1]) if fr[want_obs[x]] == it[i] * SUB + fs]
                assert sc_[i, fs] == len(mine)
                assert sb[i, fs] == (mine[0] if mine else -1)

    # ---- groups of whole points for the thread-per-observation back-substitution: every point in exactly one group or
    # in the long-track list, <= 256 observations and <= 128 points per group
    grp, big = st["point_groups"], st["point_big"]
    counts = np.diff(st["pt_ptr"])
    seen_pt = np.zeros(P, int)
    for lo, hi in grp:
        assert 0 <= lo < hi <= P and hi - lo <= 128 and counts[lo:hi].sum() <= 256
        seen_pt[lo:hi] += 1
    seen_pt[big] += 1
    assert np.all(seen_pt == 1) and np.all(counts[big] > 256) and np.all(counts[seen_pt == 1][counts[seen_pt == 1] > 256] > 0)
    assert set(big.tolist()) == set(np.nonzero(counts > 256)[0].tolist())

    # ---- where frame_blocks writes each observation's 12 panel rows
    dup = set(int(i) for i in st["dup_inc"])
    assert dup == set(int(i) for i in np.nonzero((sc_ > 1).any(axis=1))[0])
    off = st["obs_phi_off"]
    want_off = np.full(max(N, 1), -1, np.int64)
    for i in range(n_inc):
        if i in dup:
            continue
        for fs in range(SUB):
            if sc_[i, fs] == 1:
                want_off[want_obs[sb[i, fs]]] = i * PANEL_DOUBLES + fs * FP
    assert np.array_equal(off, want_off)

    # ---- half masks
    half = st["inc_half"]
    for i in range(n_inc):
        m = (1 if sc_[i, 0] or sc_[i, 1] else 0) | (2 if sc_[i, 2] or sc_[i, 3] else 0)
        if free_cam and cam_inc[ip[i]] == i:
            m |= 1 << (cam_slot // 2)
        assert half[i] == m

    # ---- random panels with the structural zero rows the kernels leave
    rng = np.random.default_rng(seed)
    Phi = np.zeros((n_inc + 1, SUB * FP, 3))
    for i in range(n_inc):
        for fs in range(SUB):
            if sc_[i, fs] or (free_cam and cam_inc[ip[i]] == i and fs == cam_slot):
                Phi[i, fs * FP:(fs + 1) * FP] = rng.standard_normal((FP, 3))
    H = 2 * T
    direct = {}
    for p in range(P):
        for x in range(pip[p], pip[p + 1]):
            for y in range(x, pip[p + 1]):
                key = (int(it[x]), int(it[y]))
                direct[key] = direct.get(key, 0) + Phi[y] @ Phi[x].T

    # ---- execute the work lists the way schur_syrk_kernel + schur_reduce do
    pa, pb, entries = st["pair_a"], st["pair_b"], st["entries"]
    blocks, n_real_entries, pos = execute_work_lists(st, Phi)
    for (a, b), acc in blocks.items():
        want = direct.pop((a, b), np.zeros_like(acc))
        if a == b:   # the kernel produces the lower triangle of a diagonal pair only
            acc, want = np.tril(acc), np.tril(want)
        assert np.allclose(acc, want, rtol=1e-12, atol=1e-12), (a, b)
    assert not direct, f"pairs with Schur terms but no work items: {list(direct)[:5]}"
    assert n_real_entries == sum((pip[p + 1] - pip[p]) * (pip[p + 1] - pip[p] + 1) // 2 for p in range(P))
    assert entries.shape[0] == max(pos, 1)

    # ---- every diagonal sub-tile with a frame, and every prior coupling, is a pair (even without a Schur term)
    have = set(zip(pa.tolist(), pb.tolist()))
    for t in range((Fc + SUB - 1) // SUB):
        assert (t, t) in have
    for a, b in priors:
        lo, hi = min(a, b) // SUB, max(a, b) // SUB
        assert (lo, hi) in have
        if free_ratio:
            assert (lo, cam_sub) in have and (hi, cam_sub) in have

    # ---- the tile plan covers every pair's Cholesky tile; forward-substitution slots point at the right tile
    tile_pos, slot = st["plan.tile_pos"], st["plan.tile_slot"]
    for a, b in have:
        i, j = sorted((int(tile_pos[a // 2]), int(tile_pos[b // 2])))
        assert slot[j * T + i] >= 0, (a, b)
    trsm, lptr, lcols, fwd = st["plan.trsm"], st["plan.lrow_ptr"], st["plan.lrow_cols"], st["fwd_slot"]
    for t in range(trsm.shape[0]):
        i, k = (int(v) for v in trsm[t])
        assert lptr[i] <= fwd[t] < lptr[i + 1] and lcols[fwd[t]] == k
    return st


def test_structure_of_c1():
    sc = small_scene()
    check_structure(sc.obs_frame, sc.obs_point, sc.num_frames, sc.num_points)


@pytest.mark.parametrize("F", [1, 3, 4, 5, 8, 9, 17, 37])
def test_structure_ragged_tracks_around_tile_borders(F):
    fr, pt, cp = ragged_topology(100 + F, F, 60)
    check_structure(fr, pt, F, 60, cp, seed=F)


@pytest.mark.parametrize("F", [7, 8, 12, 21])
def test_structure_with_intrinsics_pseudo_frame(F):
    fr, pt, cp = ragged_topology(200 + F, F, 50)
    check_structure(fr, pt, F, 50, cp, free_cam=True, seed=F)


@pytest.mark.parametrize("free_cam", [False, True])
def test_structure_with_motion_priors_and_free_ratio(free_cam):
    F = 19
    fr, pt, cp = ragged_topology(300, F, 40, max_track=4)
    priors = [(f, f - 1) for f in range(1, F)]
    check_structure(fr, pt, F, 40, cp, free_cam=free_cam, free_ratio=True, priors=priors)
    check_structure(fr, pt, F, 40, cp, free_cam=free_cam, free_ratio=False, priors=priors)


def test_structure_long_tracks_split_into_segments_of_512_entries():
    # 700 points all seen by the same two sub-tiles: (0, 1) needs two work items per half-mask class
    F, P = 8, 700
    fr = np.repeat(np.arange(F, dtype=np.int32), P)
    pt = np.tile(np.arange(P, dtype=np.int32), F)
    st = check_structure(fr, pt, F, P)
    assert st["n_items"] == 3 * 2 and np.array_equal(st["items"][:, 2], [512, 192] * 3)


def test_structure_sparse_key_path_is_identical():
    fr, pt, cp = ragged_topology(7, 41, 80)
    a = api.analyze_structure(fr, pt, 41, 80, cp)
    b = api.analyze_structure(fr, pt, 41, 80, cp, sparse_keys=True)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("free_cam", [False, True])
def test_structure_is_independent_of_the_host_thread_count(free_cam, monkeypatch):
    # the analysis runs on a few host threads (contiguous point ranges); entries of a (pair, class) must stay in
    # point order whatever the count -- the GPU summation order, hence bit reproducibility, depends on it
    F, P = 57, 900
    fr, pt, cp = ragged_topology(11, F, P, max_track=12)
    priors = [(f, f - 1) for f in range(1, F)]
    monkeypatch.setenv("RSBA_CUDA_HOST_THREADS", "1")
    ref = check_structure(fr, pt, F, P, cp, free_cam=free_cam, free_ratio=True, priors=priors)
    for n in (2, 3, 8, 13):
        monkeypatch.setenv("RSBA_CUDA_HOST_THREADS", str(n))
        got = api.analyze_structure(fr, pt, F, P, cp, free_cam, True, [a for a, _ in priors], [b for _, b in priors])
        for k in ref:
            assert np.array_equal(ref[k], got[k]), (n, k)
    # more threads than points, and threads whose point range is empty
    monkeypatch.setenv("RSBA_CUDA_HOST_THREADS", "16")
    fr2, pt2, cp2 = ragged_topology(12, 9, 5)
    check_structure(fr2, pt2, 9, 5, cp2, free_cam=free_cam)


@pytest.mark.parametrize("world,free_cam", [(2, False), (3, True)])
def test_structure_of_a_multi_gpu_shard(world, free_cam):
    # every rank keeps all observations of the points it owns: the ranks' work lists together must give the
    # single-GPU Schur complement, and every rank must derive the SAME tile plan (it comes from the whole scene)
    F, P = 45, 500
    fr, pt, cp = ragged_topology(21, F, P, max_track=10)
    priors = [(f, f - 1) for f in range(1, F)]
    pf, pp = [a for a, _ in priors], [b for _, b in priors]
    one = api.analyze_structure(fr, pt, F, P, cp, free_cam, False, pf, pp)
    rng = np.random.default_rng(3)
    panel = {}                                     # one random panel per (sub-tile, point), shared by all ranks

    def panels_of(st):
        Phi = np.zeros((st["n_inc"] + 1, SUB * FP, 3))
        cam_slot = F % SUB
        for i in range(st["n_inc"]):
            key = (int(st["inc_tile"][i]), int(st["inc_point"][i]))
            if key not in panel:
                panel[key] = rng.standard_normal((SUB * FP, 3))
            for fs in range(SUB):
                if st["slot_cnt"][i, fs] or (free_cam and st["cam_inc"][st["inc_point"][i]] == i and fs == cam_slot):
                    Phi[i, fs * FP:(fs + 1) * FP] = panel[key][fs * FP:(fs + 1) * FP]
        return Phi

    want, n_one, _ = execute_work_lists(one, panels_of(one))
    total, n_sum, owned, seen_obs = {}, 0, np.zeros(P, int), np.zeros(fr.size, int)
    owners = api.point_owners(type("S", (), dict(obs_frame=fr, obs_point=pt, num_frames=F, num_points=P))(), world)
    for rank in range(world):
        st = api.analyze_structure(fr, pt, F, P, cp, free_cam, False, pf, pp, rank=rank, world_size=world)
        assert np.array_equal(st["point_owned"], owners == rank)
        ids = st["local_ids"]
        assert np.array_equal(ids, np.nonzero(owners[pt] == rank)[0])     # all observations of the rank's points
        owned += st["point_owned"]
        seen_obs[ids] += 1
        for k in st:
            if k.startswith("plan.") or k in ("T", "fwd_slot"):
                assert np.array_equal(st[k], one[k]), (rank, k)
        # the rank's lists refer to ITS observations: point CSR over the local list
        assert np.array_equal(st["pt_obs"], np.argsort(pt[ids], kind="stable"))
        blocks, n_real, _ = execute_work_lists(st, panels_of(st))
        n_sum += n_real
        for key, blk in blocks.items():
            total[key] = total.get(key, 0) + blk
    assert np.all(owned == 1) and np.all(seen_obs == 1) and n_sum == n_one
    assert set(k for k, v in total.items() if np.any(v)) == set(k for k, v in want.items() if np.any(v))
    for key, blk in want.items():
        got = total.get(key, np.zeros_like(blk))
        if key[0] == key[1]:
            got, blk = np.tril(got), np.tril(blk)
        assert np.allclose(got, blk, rtol=1e-12, atol=1e-12), key


def test_structure_empty_and_all_constant():
    st = api.analyze_structure(np.zeros(0, np.int32), np.zeros(0, np.int32), 5, 3)
    assert st["n_inc"] == 0 and st["n_items"] == 0 and st["pair_a"].size == 2      # the diagonal sub-tiles
    fr, pt, _ = ragged_topology(9, 10, 20)
    st = check_structure(fr, pt, 10, 20, np.ones(20, np.uint8))
    assert st["n_inc"] == 0 and st["n_items"] == 0


def test_structure_argument_checks():
    fr = np.array([1, 0], np.int32)
    with pytest.raises(api.RsbaError):
        api.analyze_structure(fr, np.zeros(2, np.int32), 2, 1)             # not sorted by frame
    with pytest.raises(api.RsbaError):
        api.analyze_structure(np.array([0, 5], np.int32), np.zeros(2, np.int32), 2, 1)   # frame out of range
    with pytest.raises(api.RsbaError):
        api.analyze_structure(np.zeros(300, np.int32), np.zeros(300, np.int32), 1, 1)    # > 255 in one frame


@pytest.mark.parametrize("threads", ["1", "5"])
def test_scene_intake_is_a_stable_sort_by_frame(threads, monkeypatch):
    # CeresHandler::Add inserts frame-major (CeresHandler.h:208-255); whatever order the caller uses, the library
    # keeps the caller's order inside a frame
    monkeypatch.setenv("RSBA_CUDA_HOST_THREADS", threads)
    rng = np.random.default_rng(5)
    fr = rng.integers(0, 13, size=4000).astype(np.int32)
    pt = rng.integers(0, 50, size=4000).astype(np.int32)
    order = api.sort_observations(fr, pt, 13, 50)
    assert np.array_equal(order, np.argsort(fr, kind="stable"))
    srt = np.sort(fr)
    assert np.array_equal(api.sort_observations(srt, pt, 13, 50), np.arange(4000))      # sorted input: identity
    # a single inversion, at a thread border or anywhere else, is seen
    for at in (1, 799, 800, 801, 3999):
        f2 = srt.copy()
        f2[at - 1], f2[at] = 12, 0
        assert np.array_equal(api.sort_observations(f2, pt, 13, 50), np.argsort(f2, kind="stable")), at
    assert api.sort_observations(np.zeros(0, np.int32), np.zeros(0, np.int32), 3, 3).size == 0
    for bad_fr, bad_pt in ((13, 0), (-1, 0), (0, 50), (0, -1)):
        f2, p2 = fr.copy(), pt.copy()
        f2[3999], p2[3999] = bad_fr, bad_pt
        with pytest.raises(api.RsbaError):
            api.sort_observations(f2, p2, 13, 50)


def test_k2_algorithm_over_the_real_work_lists_gives_the_restated_reduced_system():
    """K2 end to end on the CPU: the panels  F_i = Jc_i^T (Jx_i s_p) L^-T  (k2_normal.cu / k2_schur.cu header), summed
    over the REAL work lists of the structure analysis, reduced as schur_reduce does and finalised as schur_finalize
    does (camera Jacobi scaling + LM diagonal after the sum, identity rows for constant parameters), must be the reduced
    camera system of the Ceres-style LM step restated in oracle/lm_oracle.py."""
    import oracle
    from oracle import lm_oracle
    sc = small_scene()
    F, P, radius = sc.num_frames, sc.num_points, 1e3
    r, J, valid = oracle.evaluate(sc)
    assert valid.all()
    want = lm_oracle.lm_step(sc, r, J, radius)
    fr, pt = sc.obs_frame, sc.obs_point
    Jc = np.concatenate([J[:, :12].reshape(-1, 2, 6), J[:, 12:24].reshape(-1, 2, 6)], axis=2)       # [N, 2, 12]
    Jx = J[:, 24:].reshape(-1, 2, 3)
    free_c = np.repeat(~np.asarray(sc.const_frames, dtype=bool), 12).reshape(F, 12)
    Jc = Jc * free_c[fr][:, None, :]                                   # constant camera parameters: zero columns
    # unscaled blocks
    B = np.zeros((F, 12, 12)); np.add.at(B, fr, np.einsum("nri,nrj->nij", Jc, Jc))
    C = np.zeros((P, 3, 3)); np.add.at(C, pt, np.einsum("nri,nrj->nij", Jx, Jx))
    gc = np.zeros((F, 12)); np.add.at(gc, fr, np.einsum("nri,nr->ni", Jc, r))
    gp = np.zeros((P, 3)); np.add.at(gp, pt, np.einsum("nri,nr->ni", Jx, r))
    diagB = np.einsum("fii->fi", B)
    s_c = np.where(free_c, 1.0 / (1.0 + np.sqrt(diagB)), 1.0)
    s_p = 1.0 / (1.0 + np.sqrt(np.einsum("pii->pi", C)))
    # point side (point_invert): damped scaled block, its Cholesky factor, t_p = Cinv g_p
    Cs = C * s_p[:, :, None] * s_p[:, None, :]
    Cs = Cs + np.einsum("pi,ij->pij", np.clip(np.einsum("pii->pi", Cs), 1e-6, 1e32) / radius, np.eye(3))
    Linv = np.linalg.inv(np.linalg.cholesky(Cs))                       # Minv
    tp = np.einsum("pi,pij,pj->pi", s_p, np.linalg.inv(Cs), s_p * gp)  # s_p C'^-1 s_p g_p
    # per-observation panel rows and the rhs correction w_f
    Fi = np.einsum("nri,nrk,nk,nlk->nil", Jc, Jx, s_p[pt], Linv[pt])   # Jc^T (Jx s_p) L^-T   [N, 12, 3]
    wf = np.zeros((F, 12)); np.add.at(wf, fr, np.einsum("nri,nrk,nk->ni", Jc, Jx, tp[pt]))
    st = api.analyze_structure(fr, pt, F, P)
    Phi = np.zeros((st["n_inc"] + 1, SUB * FP, 3))
    for i in range(st["n_inc"]):
        for fs in range(SUB):
            b, c = st["slot_beg"][i, fs], st["slot_cnt"][i, fs]
            for x in range(b, b + c):
                Phi[i, fs * FP:(fs + 1) * FP] += Fi[st["pt_obs"][x]]
    blocks, _, _ = execute_work_lists(st, Phi)
    n = 12 * F
    S = np.zeros((n + 48, n + 48))                                     # (room for the padding frames of the last sub-tile)
    for (a, b), blk in blocks.items():
        S[48 * b:48 * b + 48, 48 * a:48 * a + 48] = -blk if a != b else -np.tril(blk)
    S = S[:n, :n]
    S = np.tril(S) + np.tril(S, -1).T
    for f in range(F):
        S[12 * f:12 * f + 12, 12 * f:12 * f + 12] += B[f]
    # finalize (after the all-reduce on several GPUs): camera scaling, LM diagonal, identity rows
    sv, act = s_c.reshape(-1), free_c.reshape(-1)
    S = S * sv[:, None] * sv[None, :]
    S[np.diag_indices(n)] += np.clip(sv * diagB.reshape(-1) * sv, 1e-6, 1e32) / radius
    S[~act, :] = 0.0
    S[:, ~act] = 0.0
    S[~act, ~act] = 1.0
    rhs = np.where(act, sv * (gc - wf).reshape(-1), 0.0)
    assert np.linalg.norm(S - want["S"]) <= 1e-11 * np.linalg.norm(want["S"])
    assert np.linalg.norm(-rhs - want["rhs"]) <= 1e-11 * np.linalg.norm(want["rhs"])


def test_point_groups_with_long_tracks_and_the_point_limit():
    """A track longer than a group (300 observations of one point) goes to the long-track list and splits its
    neighbours into two groups; 400 single-observation points need four groups of <= 128 points."""
    F = 300
    fr = np.concatenate([np.arange(5), np.arange(F), np.arange(7)]).astype(np.int32)
    pt = np.concatenate([np.zeros(5), np.ones(F), np.full(7, 2)]).astype(np.int32)
    order = np.argsort(fr, kind="stable")
    st = api.analyze_structure(fr[order], pt[order], F, 3)
    assert st["point_big"].tolist() == [1] and st["point_groups"].tolist() == [[0, 1], [2, 3]]
    P = 400
    st = api.analyze_structure(np.zeros(P, np.int32), np.arange(P, dtype=np.int32), 4, P)
    assert st["point_groups"].tolist() == [[0, 128], [128, 256], [256, 384], [384, 400]] and st["point_big"].size == 0
