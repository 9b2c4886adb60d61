"""Multi-GPU parity on real devices (skipped on a single-GPU box; the host-side sharding logic and the algebra
of the exchange are covered on CPU by tests/test_multi_gpu_host.py): a 2-rank solve with motion priors, the
free interFrameRatio, GoodPosePrior blocks and free intrinsics must reproduce the single-GPU solve."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_gpu_solve_with_camera_only_blocks_matches_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("  OK") == 2 and "MISMATCH" not in r.stdout


def _solve_single(sc, iters):
    import rsba_b200.api as api
    with api.Problem(0) as pb:
        pb.load_scene(sc)
        s = pb.solve(api.default_options(max_num_iterations=iters, function_tolerance=0.0, gradient_tolerance=0.0,
                                         parameter_tolerance=0.0))
        po, pt = pb.get_parameters()
    return s, po, pt


@pytest.mark.gpu
def test_single_process_multi_handle_on_one_device_equals_plain_solve():
    """rsba_cuda_create_multi with ONE device (the single-GPU box of the driver's GPU tier): the forwarding path --
    borrowed per-rank handles, rsba_cuda_multi_solve, results from rank 0 -- must be the plain solve, bit for bit."""
    import numpy as np
    import rsba_b200.api as api
    from rsba_b200.scene import make_scene
    sc = make_scene(24, 1500, 12, name="multi1")
    s1, po1, pt1 = _solve_single(sc, 5)
    with api.MultiProblem([0]) as mp:
        assert len(mp.ranks) == 1
        mp.load_scene(sc)
        s = mp.solve(api.default_options(max_num_iterations=5, function_tolerance=0.0, gradient_tolerance=0.0,
                                         parameter_tolerance=0.0))
        po, pt = mp.get_parameters()
    assert s.iterations == s1.iterations and s.final_cost == s1.final_cost
    assert np.array_equal(po, po1) and np.array_equal(pt, pt1)


@pytest.mark.gpu
def test_single_process_two_devices_matches_one_gpu():
    """One host thread, two GPUs (rsba_cuda_create_multi): the reference's single-threaded driver
    (VideoSfMHandler::BA, VideoSfMHandler.cc:574-631) can use the box without becoming N processes."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import rsba_b200.api as api
    from rsba_b200.scene import make_scene
    sc = make_scene(64, 6000, 16, name="multi2")
    s1, po1, pt1 = _solve_single(sc, 6)
    with api.MultiProblem([0, 1]) as mp:
        mp.load_scene(sc)
        s = mp.solve(api.default_options(max_num_iterations=6, function_tolerance=0.0, gradient_tolerance=0.0,
                                         parameter_tolerance=0.0))
        po, pt = mp.get_parameters()
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))  # noqa: E731
    assert s.iterations == s1.iterations
    assert abs(s.final_cost - s1.final_cost) <= 1e-9 * s1.final_cost
    assert rel(po, po1) <= 1e-9 and rel(pt, pt1) <= 1e-9


def test_create_multi_rejects_bad_device_lists():
    """Host-side argument checks of rsba_cuda_create_multi (no device needed for these paths)."""
    import ctypes as C
    import numpy as np
    import rsba_b200.api as api
    lib = api.load_library()
    m = C.c_void_p()
    dup = np.array([0, 0], dtype=np.int32)
    assert lib.rsba_cuda_create_multi(C.byref(m), dup.ctypes.data_as(C.POINTER(C.c_int)), 2) == -1
    assert lib.rsba_cuda_create_multi(C.byref(m), None, 1) == -1
    assert lib.rsba_cuda_multi_size(None) == 0 and lib.rsba_cuda_multi_handle(None, 0) is None
    lib.rsba_cuda_destroy_multi(None)
