"""Host logic of the multi-GPU path, on CPU with world_size 2 over gloo (no GPU):

* the sharding rule (rsba_cuda_point_owners): every point has one owner, owners follow contiguous
  frame ranges, a rank's share is "all observations of its points";
* the algebra of the exchange step: each rank forms the UNSCALED partial reduced system
  [B - E C^-1 E^T | g_c - E C^-1 g_p | diag B] from its share alone, one all-reduce (sum) adds them,
  and only then the camera-side Jacobi scaling and LM diagonal are applied -- the result must be
  the reduced system of the whole problem (oracle/lm_oracle.py).  This is exactly what
  lm_solver.cu does around ncclAllReduce (schur_reduce -> all-reduce -> schur_finalize).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from rsba_b200.scene import Scene, make_scene  # noqa: E402


def test_point_owner_rule():
    import rsba_b200.api as api
    sc = make_scene(64, 3000, 10, name="own")
    for world in (1, 2, 4, 8):
        own = api.point_owners(sc, world)
        assert own.min() >= 0 and own.max() < world
        if world == 1:
            assert not own.any()
            continue
        counts = np.bincount(own, minlength=world)
        assert counts.min() > 0.5 * counts.mean()                   # balanced for a uniform video
        # owners follow the median observation frame: monotone in it
        med = np.array([np.median(sc.obs_frame[sc.obs_point == p]) for p in range(0, sc.num_points, 37)])
        o = own[::37]
        order = np.argsort(med, kind="stable")
        assert np.all(np.diff(o[order]) >= 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _partial_system(sc, r, J, owned_pts, radius, lo):
    """Unscaled partial reduced system of the observations whose point is in owned_pts."""
    import scipy.sparse as sp
    F, P = sc.num_frames, sc.num_points
    keep = owned_pts[sc.obs_point]
    sub = Scene(**{**sc.__dict__, "obs_xy": sc.obs_xy[keep], "obs_frame": sc.obs_frame[keep],
                   "obs_point": sc.obs_point[keep]})
    act_c, act_p = lo.param_masks(sc)
    active = np.concatenate([act_c, act_p])
    Js = lo.sparse_jacobian(sub, J[keep], active)
    rr = r[keep].reshape(-1)
    H = (Js.T @ Js).tocsr()
    g = Js.T @ rr
    nc = 12 * F
    B = H[:nc, :nc].toarray()
    E = H[:nc, nc:].tocsr()
    Cd = H[nc:, nc:].tocoo()
    C = np.zeros((P, 3, 3))
    C[Cd.row // 3, Cd.row % 3, Cd.col % 3] = Cd.data
    # point side: Jacobi scaling + LM diagonal are local to the owner
    sp_ = 1.0 / (1.0 + np.sqrt(np.einsum("pii->pi", C)))
    Cs = C * sp_[:, :, None] * sp_[:, None, :]
    D2 = np.clip(np.einsum("pii->pi", Cs), 1e-6, 1e32) / radius
    Cs = Cs + np.einsum("pi,ij->pij", D2, np.eye(3))
    Cinv = np.linalg.inv(Cs) * sp_[:, :, None] * sp_[:, None, :]
    Cinv[~owned_pts] = 0.0
    ECinv = (E @ lo._block_diag(Cinv)).tocsr()
    S = B - (ECinv @ E.T).toarray()
    rhs = g[:nc] - ECinv @ g[nc:]
    return S, rhs, np.diag(B).copy()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        import rsba_b200.api as api
        from oracle import lm_oracle as lo
        sc = make_scene(24, 700, 8, name="mgpu")
        radius = 1e3
        r, J, v = oracle.evaluate(sc, impl="port")
        own = api.point_owners(sc, world) == rank
        S, rhs, dB = _partial_system(sc, r, J, own, radius, lo)
        buf = torch.from_numpy(np.concatenate([S.reshape(-1), rhs, dB]))
        dist.all_reduce(buf)                                           # the one exchange step
        n = 12 * sc.num_frames
        S, rhs, dB = buf[:n * n].numpy().reshape(n, n), buf[n * n:n * n + n].numpy(), buf[n * n + n:].numpy()
        # finalize: camera Jacobi scaling, LM diagonal, constant rows
        act_c, _ = lo.param_masks(sc)
        s = np.where(act_c, 1.0 / (1.0 + np.sqrt(dB)), 1.0)
        S = S * s[:, None] * s[None, :]
        S[np.diag_indices(n)] += np.clip(s * dB * s, 1e-6, 1e32) / radius
        S[~act_c, :] = 0.0
        S[:, ~act_c] = 0.0
        S[~act_c, ~act_c] = 1.0
        rhs = np.where(act_c, s * rhs, 0.0)
        want = lo.lm_step(sc, r, J, radius)
        e1 = np.linalg.norm(S - want["S"]) / np.linalg.norm(want["S"])
        e2 = np.linalg.norm(-rhs - want["rhs"]) / np.linalg.norm(want["rhs"])
        q.put((rank, float(e1), float(e2), int(own.sum())))
    finally:
        dist.destroy_process_group()


def test_allreduce_of_unscaled_partials_gives_the_global_reduced_system():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(x[3] for x in res) == 700                      # every point has exactly one owner
    for rank, e1, e2, _ in res:
        assert e1 <= 1e-10 and e2 <= 1e-10, (rank, e1, e2)
