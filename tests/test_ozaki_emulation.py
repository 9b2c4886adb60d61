"""Sizing the tcgen05 route for the FP64 Schur SYRK (task 10): an Ozaki-style split of the panels
Y = E L^-T into 7-bit integer slices, contracted exactly in integers (what int8 tensor-core MMAs with int32
accumulators compute), recombined in FP64.  This file is the ERROR half of the sizing -- exact emulation in numpy on
the real panels of an RS scene; the THROUGHPUT half is tools/imma_peak.cu.  DESIGN.md section 9 item 4 has the verdict.

    S = B + D^2 - Y Y^T,      Y (12F x 3P) row-scaled to |y| < 1, y = sum_t q_t 2^(-7 (t+1)),  q_t in [-64, 64]
    Y Y^T ~ sum_{t+u <= s-1} 2^(-7 (t+u+2)) Q_t Q_u^T      s (s + 1) / 2 integer GEMMs for s slices

Measured here (C1-like scene, radius 1e4): relative error of the LM camera step against the FP64 Schur complement."""
import numpy as np
import pytest
import scipy.linalg
import scipy.sparse as sp

import oracle
from oracle import lm_oracle as lo
from helpers import small_scene


def schur_parts(scene, radius=1e4):
    r, J, _ = oracle.evaluate(scene, impl="port")
    F, P = scene.num_frames, scene.num_points
    nc = 12 * F
    act_c, act_p = lo.param_masks(scene, None, None)
    active = np.concatenate([act_c, act_p])
    Js = lo.sparse_jacobian(scene, J, active)
    scale = lo.jacobi_scale(Js, active, True)
    Jp = Js @ sp.diags(scale)
    H = (Jp.T @ Jp).tocsr()
    g = Jp.T @ r.reshape(-1)
    D2 = np.clip(H.diagonal(), 1e-6, 1e32) / radius
    D2[~active] = 1.0
    B = H[:nc, :nc].toarray() + np.diag(D2[:nc])
    E = H[:nc, nc:].toarray()
    C = np.zeros((P, 3, 3))
    Cd = H[nc:, nc:].tocoo()
    C[Cd.row // 3, Cd.row % 3, Cd.col % 3] = Cd.data
    C += np.einsum("pi,ij->pij", D2[nc:].reshape(P, 3), np.eye(3))
    Y = np.zeros((nc, 3 * P))
    for p in range(P):                                   # Y_p = E_p L_p^-T  (the panels of k2_schur.cu)
        L = np.linalg.cholesky(C[p])
        Y[:, 3 * p:3 * p + 3] = scipy.linalg.solve_triangular(L, E[:, 3 * p:3 * p + 3].T, lower=True).T
    Cinv = np.linalg.inv(C)
    w = np.einsum("cpi,pij,pj->c", E.reshape(nc, P, 3), Cinv, g[nc:].reshape(P, 3))
    return B, Y, g[:nc] - w


def ozaki_gram(Y, slices):
    """Y Y^T from `slices` 7-bit integer slices per entry, rows scaled by a power of two; integer products are exact."""
    e = np.ceil(np.log2(np.maximum(np.abs(Y).max(axis=1), 1e-300) / 0.99))
    rem = Y / 2.0 ** e[:, None]
    Q = []
    for _ in range(slices):
        rem = rem * 128.0
        q = np.rint(rem)
        assert np.abs(q).max() <= 127                    # fits int8
        Q.append(q.astype(np.int64))
        rem = rem - q
    G = np.zeros((Y.shape[0], Y.shape[0]))
    n_gemm = 0
    for t in range(slices):
        for u in range(slices - t):
            acc = Q[t] @ Q[u].T                          # exact: |acc| <= K 127^2 << 2^63 (and < 2^31 for K < 133 000)
            assert np.abs(acc).max() < 2 ** 31
            G += acc.astype(np.float64) * 2.0 ** (-7 * (t + u + 2))
            n_gemm += 1
    return G * 2.0 ** (e[:, None] + e[None, :]), n_gemm


@pytest.fixture(scope="module")
def parts():
    return schur_parts(small_scene())


def test_ozaki_slices_reach_fp64_schur_accuracy(parts):
    B, Y, rhs = parts
    S64 = B - Y @ Y.T
    y64 = scipy.linalg.cho_solve(scipy.linalg.cho_factor(S64, lower=True), rhs)
    cond = np.linalg.cond(S64)
    rows = []
    for s in (3, 4, 5, 6, 7, 8):
        G, n_gemm = ozaki_gram(Y, s)
        S = B - G
        err_S = np.abs(S - S64).max() / np.abs(S64).max()
        try:
            y = scipy.linalg.cho_solve(scipy.linalg.cho_factor(S, lower=True), rhs)
            err_y = np.linalg.norm(y - y64) / np.linalg.norm(y64)
        except np.linalg.LinAlgError:
            err_y = np.inf
        rows.append((s, n_gemm, err_S, err_y))
    print(f"\ncond(S) = {cond:.2e}")
    for s, n, eS, ey in rows:
        print(f"slices {s}: {n:2d} int8 GEMMs, max |S - S64| / max |S64| = {eS:.2e}, camera step rel. error = {ey:.2e}")
    errs = [r[3] for r in rows]
    assert all(b <= a * 1.0001 for a, b in zip(errs, errs[1:]) if np.isfinite(a))     # more slices never hurt
    by = {r[0]: r for r in rows}
    assert by[7][3] <= 1e-6                      # the 1e-6 bar of north_star needs <= 7 slices (28 GEMMs) here
    assert by[8][3] <= 1e-9
    assert by[3][3] > 1e-6                       # ... and more than three: the split is not free
