"""Host logic (no GPU): the numeric phase of the reduced-system solve as the static task list that the
persistent kernel rsba_b200/csrc/k3_dag.cu executes (the stand-in for CHOLMOD's numeric factorisation +
solve behind Ceres' SPARSE_SCHUR, CeresHandler.h:403).  The list is executed here in numpy in LIST ORDER,
by one worker: every wait of a task must already be satisfied by the tasks in front of it (the list is a
topological order -- the property that makes in-order fetching by any number of CTAs deadlock-free), and
what comes out must be numpy's Cholesky factor and the solution of A x = b."""
import numpy as np
import pytest

import rsba_b200.api as api
from test_tile_plan import B, band_pairs, random_pairs, spd_with_pattern

PARTS = 3          # kTrsmParts
H = B // 2         # quadrant of the numpy stand-in tile


def run_task_graph(plan, dag, A, T, rhs):
    pos = plan["tile_pos"]
    perm = np.empty(T * B, dtype=int)
    for t in range(T):
        perm[pos[t] * B:(pos[t] + 1) * B] = np.arange(t * B, (t + 1) * B)
    M = A[np.ix_(perm, perm)].copy()
    nz = [tuple(int(v) for v in x) for x in plan["nz_tiles"]]
    slot_of = {ij: s for s, ij in enumerate(nz)}
    tiles = [M[i * B:(i + 1) * B, j * B:(j + 1) * B].copy() for (i, j) in nz]     # tile-packed S
    x = np.array(rhs, dtype=np.float64)[perm].copy()
    n_nz = len(nz)
    ready = np.zeros(n_nz, int)
    done = np.zeros((n_nz, 4), int)
    need = dag["need"]
    yready = np.zeros(T, int)
    bcnt = np.zeros(T, int)
    dinv = {}
    # lists the kernel indexes: row lists (forward terms) and column lists (backward terms)
    lrow = {i: [j for (a, j) in nz if a == i and j < i] for i in range(T)}
    col_rows = {k: [i for (i, b) in nz if b == k and i > k] for k in range(T)}
    fwd = {}
    bwd = {}
    nf = dag["n_factor_tasks"]
    seen_back = False
    for n, (typ, a, b, c, d, e, f, g) in enumerate(dag["tasks"]):
        if typ in (api.TASK_BACK_FIN, api.TASK_BACK_TILE):
            seen_back = True
            assert n >= nf
        else:
            assert not seen_back and n < nf, "factorisation tasks come first"
        if typ == api.TASK_FACTOR:
            k, s = a, b
            assert nz[s] == (k, k)
            assert np.all(done[s] >= need[s]), "FACTOR fetched before its updates"
            Lkk = np.linalg.cholesky(np.tril(tiles[s]) + np.tril(tiles[s], -1).T)
            tiles[s] = Lkk
            dinv[k] = np.linalg.inv(Lkk)
            terms = [fwd[(k, j)] for j in lrow[k]]          # KeyError = a term is not there yet
            x[k * B:(k + 1) * B] = dinv[k] @ (x[k * B:(k + 1) * B] - sum(terms, np.zeros(B)))
            ready[s] += 1
        elif typ == api.TASK_TRSM:
            i, k, part, s, skk, fslot = a, b, c, d, e, f
            assert nz[s] == (i, k) and nz[skk] == (k, k) and i > k
            assert np.all(done[s] >= need[s]), "TRSM fetched before the tile's updates"
            assert ready[skk] == 1, "TRSM fetched before its panel's FACTOR"
            if part == 0:                                  # (the 4-row numpy tile does not split in 3 slabs: slab 0 does it all)
                tiles[s] = tiles[s] @ dinv[k].T
                fwd[(i, k)] = tiles[s] @ x[k * B:(k + 1) * B]
            assert lrow[i][fslot - plan_lrow_ptr(plan, lrow, i)] == k
            ready[s] += 1
        elif typ == api.TASK_UPDATE:
            s, q, order, first, count = a, b, c, d, e
            i, j = nz[s]
            qi, qj = q >> 1, q & 1
            assert not (i == j and q == 1)
            assert done[s, q] == order, "update groups of one target run in their fixed order"
            acc = np.zeros((H, H))
            for sik, sjk in dag["sources"][first:first + count]:
                (i2, k), (j2, k2) = nz[sik], nz[sjk]
                assert (i2, j2) == (i, j) and k == k2 and k < j
                assert ready[sik] == PARTS and ready[sjk] == PARTS, "update fetched before its sources"
                acc += tiles[sik][qi * H:(qi + 1) * H] @ tiles[sjk][qj * H:(qj + 1) * H].T
            tiles[s][qi * H:(qi + 1) * H, qj * H:(qj + 1) * H] -= acc
            done[s, q] += 1
        elif typ == api.TASK_BACK_FIN:
            k, s, first, cnt = a, b, c, d
            assert nz[s] == (k, k) and ready[s] == 1 and cnt == len(col_rows[k])
            assert bcnt[k] == cnt, "BACKFIN fetched before the column's terms"
            terms = [bwd[first + q] for q in range(cnt)]
            x[k * B:(k + 1) * B] = dinv[k].T @ (x[k * B:(k + 1) * B] - sum(terms, np.zeros(B)))
            yready[k] = 1
        else:
            i, k, s, bslot = a, b, c, d
            assert nz[s] == (i, k) and ready[s] == PARTS
            assert yready[i] == 1, "BACKTILE fetched before y_i"
            assert bslot not in bwd
            bwd[bslot] = tiles[s].T @ x[i * B:(i + 1) * B]
            bcnt[k] += 1
    assert np.all(ready[[slot_of[(k, k)] for k in range(T)]] == 1) and np.all(yready == 1)
    assert np.all(done >= need)
    L = np.zeros_like(M)
    for s, (i, j) in enumerate(nz):
        L[i * B:(i + 1) * B, j * B:(j + 1) * B] = np.tril(tiles[s]) if i == j else tiles[s]
    sol = np.empty_like(x)
    sol[perm] = x
    return L, perm, sol


def plan_lrow_ptr(plan, lrow, i):
    """first index of row i in the concatenated row lists (what lrow_ptr holds on the device)"""
    return sum(len(lrow[r]) for r in range(i))


CASES = [
    (1, [(0, 0)], True, 1),
    (7, band_pairs(7, 2), True, 1),
    (40, band_pairs(40, 3), True, 4),
    (40, band_pairs(40, 3), False, 1),
    (125, band_pairs(125, 3), True, 4),
    (125, band_pairs(125, 3), True, 1),
    (33, random_pairs(33, 40, 1), True, 2),
    (20, [(a, b) for a in range(20) for b in range(a, 20)], True, 4),
    (20, [(a, b) for a in range(20) for b in range(a, 20)], True, 1),
]


@pytest.mark.parametrize("T,pairs,reorder,merge", CASES)
def test_task_graph_in_list_order_factorises_and_solves(T, pairs, reorder, merge):
    pa, pb = np.array([p[0] for p in pairs]), np.array([p[1] for p in pairs])
    plan = api.plan_reduced_system(T, pa, pb, dense=False, reorder=reorder)
    dag = api.plan_task_graph(T, pa, pb, dense=False, reorder=reorder, merge_levels=merge)
    A = spd_with_pattern(T, pairs, seed=T)
    b = np.random.default_rng(3).normal(size=T * B)
    L, perm, x = run_task_graph(plan, dag, A, T, b)
    want = np.linalg.cholesky(A[np.ix_(perm, perm)])
    assert np.abs(L - want).max() <= 1e-10 * np.abs(want).max()
    assert np.abs(x - np.linalg.solve(A, b)).max() <= 1e-9 * np.abs(x).max()


def test_task_counts_and_merging():
    T = 125
    pairs = band_pairs(T, 3)
    pa, pb = np.array([p[0] for p in pairs]), np.array([p[1] for p in pairs])
    plan = api.plan_reduced_system(T, pa, pb)
    eager = api.plan_task_graph(T, pa, pb, merge_levels=1)
    merged = api.plan_task_graph(T, pa, pb, merge_levels=8)
    for dag in (eager, merged):
        t = dag["tasks"][:, 0]
        assert (t == api.TASK_FACTOR).sum() == T == (t == api.TASK_BACK_FIN).sum()
        assert (t == api.TASK_TRSM).sum() == PARTS * len(plan["trsm"])
        assert (t == api.TASK_BACK_TILE).sum() == len(plan["trsm"])
        # every update triple of the plan is applied exactly once per quadrant of its target (3 on diagonal tiles)
        upd = dag["tasks"][t == api.TASK_UPDATE]
        quads = sum(3 if i == j else 4 for i, j, k in plan["upd"])
        assert upd[:, 5].sum() == quads
    n_e = (eager["tasks"][:, 0] == api.TASK_UPDATE).sum()
    n_m = (merged["tasks"][:, 0] == api.TASK_UPDATE).sum()
    assert n_m < n_e                          # fewer read-modify-writes of the targets
    # the sources right below a target's column stay on their own: the critical path is not lengthened
    assert len(merged["sources"]) == len(eager["sources"]) == len(plan["upd"])
