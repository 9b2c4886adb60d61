"""Host logic (no GPU): the symbolic analysis of the reduced camera system -- nested-dissection
ordering, fill, elimination levels, conflict-free update groups (rsba_b200/csrc/tile_plan.cu,
the stand-in for CHOLMOD's analyse phase behind Ceres' SPARSE_SCHUR, CeresHandler.h:403).
The plan is executed here in numpy, launch by launch in the order the GPU would run it, on a
random SPD matrix with the plan's block sparsity, and compared with numpy's Cholesky."""
import numpy as np
import pytest

import rsba_b200.api as api

B = 4  # numpy stand-in for the 96-row tile (the plan is size-agnostic)


def band_pairs(T, bw):
    return [(a, b) for a in range(T) for b in range(a, min(T, a + bw + 1))]


def random_pairs(T, n, seed):
    rng = np.random.default_rng(seed)
    pr = {(min(a, b), max(a, b)) for a, b in rng.integers(0, T, (n, 2))}
    return sorted(pr | {(t, t + 1) for t in range(T - 1)})   # connected


def spd_with_pattern(T, pairs, seed=0):
    rng = np.random.default_rng(seed)
    n = T * B
    A = np.zeros((n, n))
    for a, b in pairs:
        blk = rng.normal(size=(B, B))
        A[b * B:(b + 1) * B, a * B:(a + 1) * B] = blk
        A[a * B:(a + 1) * B, b * B:(b + 1) * B] = blk.T
    A = 0.5 * (A + A.T)
    A += np.eye(n) * (np.abs(A).sum(axis=1).max() + 1.0)
    return A


def run_plan(plan, A, T, rhs=None):
    """Tile right-looking Cholesky in plan order; returns L in permuted position order plus checks.
    With ``rhs`` (in position order) also runs the forward substitution the way the device does, inside the
    factorisation's own launches: the potrf launch forms z_k = L_kk^-1 (b_k - sum of the slots of row k), the trsm
    launch leaves L_ik z_k in the slot of tile (i, k); returns (L, perm, z)."""
    pos = plan["tile_pos"]
    perm = np.empty(T * B, dtype=int)
    for t in range(T):
        perm[pos[t] * B:(pos[t] + 1) * B] = np.arange(t * B, (t + 1) * B)
    M = A[np.ix_(perm, perm)].copy()
    nz = {tuple(x) for x in plan["nz_tiles"]}
    blk = lambda i, j: (slice(i * B, (i + 1) * B), slice(j * B, (j + 1) * B))  # noqa: E731
    # the original pattern must be inside the symbolic one
    for i in range(T):
        for j in range(i + 1):
            if np.any(M[blk(i, j)] != 0):
                assert (i, j) in nz
    done = set()
    z = None if rhs is None else np.array(rhs, dtype=np.float64)
    slots = {}                                                             # (i, k) -> L_ik z_k
    for l in range(plan["n_levels"]):
        pan = plan["panels"][plan["panel_ptr"][l]:plan["panel_ptr"][l + 1]]
        for k in pan:                                                      # potrf (one launch)
            M[blk(k, k)] = np.linalg.cholesky(M[blk(k, k)])
            if z is not None:                                              # ... and its forward-substitution epilogue
                row = sorted(j for (i, j) in nz if i == k and j < k)
                assert all((k, j) in slots for j in row), "a term of row k is not there yet"
                t = z[k * B:(k + 1) * B] - sum((slots[(k, j)] for j in row), np.zeros(B))
                z[k * B:(k + 1) * B] = np.linalg.solve(M[blk(k, k)], t)
        tr = plan["trsm"][plan["trsm_ptr"][l]:plan["trsm_ptr"][l + 1]]
        for i, k in tr:                                                    # trsm (one launch)
            assert k in pan and (i, k) in nz and i > k
            M[blk(i, k)] = np.linalg.solve(M[blk(k, k)], M[blk(i, k)].T).T
            if z is not None:
                assert (i, k) not in slots
                slots[(int(i), int(k))] = M[blk(i, k)] @ z[k * B:(k + 1) * B]
        for g in range(plan["level_group_ptr"][l], plan["level_group_ptr"][l + 1]):
            ups = plan["upd"][plan["group_ptr"][g]:plan["group_ptr"][g + 1]]
            targets = [(i, j) for i, j, k in ups]
            assert len(set(targets)) == len(targets), "two updates of one launch write the same tile"
            for i, j, k in ups:                                            # update (one launch per group)
                assert k in pan and (i, j) in nz and i >= j > k
                assert j not in done and j not in pan, "update of an already factorised panel"
                M[blk(i, j)] -= M[blk(i, k)] @ M[blk(j, k)].T
        done.update(int(k) for k in pan)
    assert done == set(range(T))
    L = np.tril(M)
    for i in range(T):                      # tiles outside the symbolic pattern never received data
        for j in range(i):
            if (i, j) not in nz:
                L[blk(i, j)] = 0.0
    if z is not None:
        assert set(slots) == {(i, j) for (i, j) in nz if i > j}, "every off-diagonal tile leaves exactly one term"
        return L, perm, z
    return L, perm


@pytest.mark.parametrize("T,pairs,reorder", [
    (1, [(0, 0)], True),
    (7, band_pairs(7, 2), True),
    (40, band_pairs(40, 3), True),
    (40, band_pairs(40, 3), False),
    (125, band_pairs(125, 3), True),
    (33, random_pairs(33, 40, 1), True),
    (20, [(a, b) for a in range(20) for b in range(a, 20)], True),
])
def test_plan_factorises_like_numpy(T, pairs, reorder):
    pa, pb = np.array([p[0] for p in pairs]), np.array([p[1] for p in pairs])
    plan = api.plan_reduced_system(T, pa, pb, dense=False, reorder=reorder)
    assert sorted(plan["tile_pos"]) == list(range(T))
    if not reorder:
        assert list(plan["tile_pos"]) == list(range(T))
    A = spd_with_pattern(T, pairs)
    L, perm = run_plan(plan, A, T)
    want = np.linalg.cholesky(A[np.ix_(perm, perm)])
    assert np.abs(L - want).max() <= 1e-10 * np.abs(want).max()


def test_nested_dissection_shortens_the_critical_path():
    T = 125                                   # C3: 1000 frames, points seen from 25 consecutive frames
    pairs = band_pairs(T, 3)
    pa, pb = np.array([p[0] for p in pairs]), np.array([p[1] for p in pairs])
    nat = api.plan_reduced_system(T, pa, pb, reorder=False)
    nd = api.plan_reduced_system(T, pa, pb, reorder=True)
    assert nat["n_levels"] == T               # a band in natural order is a chain
    assert nd["n_levels"] <= 25
    dense = api.plan_reduced_system(T, pa, pb, dense=True)
    assert dense["n_levels"] == T and len(dense["nz_tiles"]) == T * (T + 1) // 2
    assert nd["flops"] < 0.05 * dense["flops"]


def test_disconnected_components_are_independent():
    T = 30
    pairs = band_pairs(15, 2) + [(a + 15, b + 15) for a, b in band_pairs(15, 2)]
    pa, pb = np.array([p[0] for p in pairs]), np.array([p[1] for p in pairs])
    plan = api.plan_reduced_system(T, pa, pb)
    A = spd_with_pattern(T, pairs, seed=3)
    L, perm = run_plan(plan, A, T)
    want = np.linalg.cholesky(A[np.ix_(perm, perm)])
    assert np.abs(L - want).max() <= 1e-10 * np.abs(want).max()
    assert plan["n_levels"] <= 15


@pytest.mark.parametrize("T,pairs,reorder", [
    (7, band_pairs(7, 2), True),
    (40, band_pairs(40, 3), False),
    (125, band_pairs(125, 3), True),
    (33, random_pairs(33, 40, 1), True),
])
def test_forward_substitution_rides_in_the_factorisation(T, pairs, reorder):
    """K3 folds z = L^-1 b into the potrf / trsm launches (k3_cholesky.cu): every term L_kj z_j that panel k needs
    was produced at the level of panel j, which the elimination order puts strictly before panel k's."""
    pa, pb = np.array([p[0] for p in pairs]), np.array([p[1] for p in pairs])
    plan = api.plan_reduced_system(T, pa, pb, dense=False, reorder=reorder)
    A = spd_with_pattern(T, pairs, seed=5)
    b = np.random.default_rng(9).normal(size=T * B)
    L, perm, z = run_plan(plan, A, T, rhs=b)
    want = np.linalg.solve(np.linalg.cholesky(A[np.ix_(perm, perm)]), b)
    assert np.abs(z - want).max() <= 1e-10 * np.abs(want).max()
