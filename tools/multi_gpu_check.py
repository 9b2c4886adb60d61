#!/usr/bin/env python
"""Multi-GPU parity of the camera-only residual blocks (run under torchrun, one rank per GPU):
motion priors with the free interFrameRatio, GoodPosePrior blocks and free intrinsics on a sharded scene
must reproduce the single-GPU solve (same handle-free scene on rank 0 without a communicator).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rsba_b200.api as api  # noqa: E402
from rsba_b200.scene import make_scene  # noqa: E402


def configure(pb, sc, free_cam):
    if free_cam:
        pb.set_intrinsics_free(True)
    pb.load_scene(sc)
    n = sc.num_frames - 1
    pb.set_motion_priors([1] * n, [6.0] * n, [1.0] * n, list(range(1, n + 1)), list(range(n)))
    pb.set_inter_frame_ratio_free(True, 1.0)
    rng = np.random.default_rng(3)
    frames = [f for f in range(1, sc.num_frames) for _ in (0, 1)]
    which = [w for _ in range(1, sc.num_frames) for w in (0, 1)]
    vals = np.array([sc.poses[f, 6 * w:6 * w + 6] + rng.normal(0, 1e-3, 6) for f, w in zip(frames, which)])
    pb.set_pose_priors(frames, which, [10.0] * len(frames), [3.0] * len(frames), vals)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api.load_library()
    sc = make_scene(64, 4000, 10, name="mgpu-priors")
    bad = 0
    for free_cam in (False, True):
        with api.Problem(local) as pb:
            uid = [api.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            pb.comm_init(rank, world, uid[0])
            configure(pb, sc, free_cam)
            s = pb.solve(api.default_options(max_num_iterations=6))
            po, pt = pb.get_parameters()
            ratio = pb.inter_frame_ratio()
            val, _ = pb.pose_priors()
        if rank == 0:
            with api.Problem(local) as one:
                configure(one, sc, free_cam)
                s1 = one.solve(api.default_options(max_num_iterations=6))
                po1, pt1 = one.get_parameters()
                ratio1 = one.inter_frame_ratio()
                val1, _ = one.pose_priors()
            rel = lambda a, b: float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))
            errs = dict(cost=abs(s.final_cost - s1.final_cost) / s1.final_cost, poses=rel(po, po1), points=rel(pt, pt1),
                        ratio=abs(ratio - ratio1), priors=rel(val, val1))
            ok = s.usable == 1 and s.iterations == s1.iterations and max(errs.values()) < 1e-9
            bad += 0 if ok else 1
            print(f"free_cam={free_cam} world={world}: cost {s.final_cost:.12e} vs {s1.final_cost:.12e}  "
                  f"iterations {s.iterations}/{s1.iterations}  " + "  ".join(f"{k} {v:.2e}" for k, v in errs.items()) +
                  ("  OK" if ok else "  MISMATCH"), flush=True)
        dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
