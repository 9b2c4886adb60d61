#!/usr/bin/env python
"""Headline benchmark: M residual+Jacobian evals/s (and LM iterations/s) on the BASELINE.json
scene (C3: 1 000 frames / 200 k points / 5 M observations), through the C ABI.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K1_BYTES_FIXED = 16 + 8 + 16 + 240  # xy, indices, residual, Jacobian


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def k1_bytes_per_obs(scene) -> float:
    return K1_BYTES_FIXED + (24.0 * scene.num_points + 96.0 * scene.num_frames) / scene.num_obs


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = device_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].strip().lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, local, world


def shard_scene(scene, rank, world):
    """Contiguous camera ranges per rank; K1 shards with no collective."""
    if world == 1:
        return scene
    from rsba_b200.scene import Scene
    bounds = np.linspace(0, scene.num_frames, world + 1).astype(int)
    keep = (scene.obs_frame >= bounds[rank]) & (scene.obs_frame < bounds[rank + 1])
    return Scene(**{**scene.__dict__, "obs_xy": scene.obs_xy[keep], "obs_frame": scene.obs_frame[keep],
                    "obs_point": scene.obs_point[keep]})


def cpu_reference_rate(scene, seconds_target=12.0, nthreads=0):
    """The reference's own functor + forward autodiff (oracle/_ref: reference headers compiled
    verbatim) on the host cores, on a bounded sample of the same scene.  Falls back to the
    plain-C port only if the prebuilt _ref library did not travel."""
    import oracle
    cores = os.cpu_count() or 1
    nthreads = nthreads or cores
    n_frames = min(scene.num_frames, 100)
    sample = scene.subscene(n_frames)
    n = sample.num_obs
    res, J, v = np.zeros((n, 2)), np.zeros((n, 30)), np.zeros(n, np.uint8)
    poses = np.ascontiguousarray(sample.poses)
    points = np.ascontiguousarray(sample.points)
    if oracle.ref_available():
        kind = "reference"
        pb = oracle.RefProblem(sample)
        run = lambda: pb.eval(poses, points, res, J, v, nthreads)  # noqa: E731
    else:
        kind = "port"
        run = lambda: oracle.evaluate(sample, jac=True, impl="port", nthreads=nthreads)  # noqa: E731
    run()  # warm (page faults, thread pool)
    t0 = time.perf_counter()
    run()
    one = time.perf_counter() - t0
    reps = int(max(3, min(200, seconds_target / max(one, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(reps):
        run()
    dt = (time.perf_counter() - t0) / reps
    return {"value": n / dt / 1e6, "unit": "M evals/s", "cores": int(nthreads), "kind": kind,
            "sample": f"first {n_frames} frames of the scene ({n} observations), residual+Jacobian by "
                      f"Jet<15> forward autodiff, {reps} passes, OpenMP static over observations",
            "ms_per_pass": dt * 1e3}, sample


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rsba_b200.scene import make_config
    scene = make_config(args.config)
    cores = os.cpu_count() or 1
    base, sample = cpu_reference_rate(scene, seconds_target=2.0)
    # K timed "steps", each one pass over the bounded sample
    import oracle
    n = sample.num_obs
    res, J, v = np.zeros((n, 2)), np.zeros((n, 30)), np.zeros(n, np.uint8)
    poses, points = np.ascontiguousarray(sample.poses), np.ascontiguousarray(sample.points)
    if oracle.ref_available():
        pb = oracle.RefProblem(sample)
        run = lambda: pb.eval(poses, points, res, J, v, cores)  # noqa: E731
    else:
        run = lambda: oracle.evaluate(sample, jac=True, impl="port", nthreads=cores)  # noqa: E731
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt / 1e6
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "M evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "step": STEP_DESC, "sample": base["sample"]},
        "cpu_baseline": {"value": value, "unit": "M evals/s", "cores": cores, "kind": base["kind"],
                         "sample": base["sample"]},
        "e2e": {"value": value, "unit": "M evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


METRIC = "M residual+Jacobian evals/sec (full LM iteration)"
WORKLOADS = {
    "C1": "C1: 10 frames / 500 points / 5k observations synthetic RS scene",
    "C2": "C2: 100 frames / 20k points / 500k observations synthetic RS scene",
    "C3": "C3: 1000 frames / 200k points / 5M observations synthetic RS scene",
    "C5": "C5: 4000 frames / 1M points / 20M observations synthetic RS scene",
}
STEP_DESC = ("one Levenberg-Marquardt iteration of rsba_cuda_solve: residual+Jacobian (K1), block normal "
             "equations + Schur complement (K2), tile Cholesky + solves (K3), back-substitution/step (K4), "
             "cost at the trial point (K1r), accept/reject")


def bench_options(api, iters):
    """Exactly `iters` LM iterations: tolerances zeroed so the loop cannot stop early."""
    return api.default_options(max_num_iterations=iters, function_tolerance=0.0, gradient_tolerance=0.0,
                               parameter_tolerance=0.0)


def run_ours(args):
    import torch
    import rsba_b200.api as api
    from rsba_b200.scene import make_config

    rank, local, world = dist_setup(args.gpus)
    t0 = time.time()
    scene = make_config(args.config)
    if rank == 0:
        log(f"[bench] scene {args.config}: {scene.num_obs} obs ready in {time.time() - t0:.1f}s")
    if world > 1:
        raise SystemExit("multi-GPU LM path: see --gpus handling in DESIGN.md (not built yet)")
    n_total = scene.num_obs
    warm = max(args.warmup, 3)

    pb = api.Problem(local)
    stream = torch.cuda.current_stream()
    pb.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    pb.load_scene(scene)
    upload_ms = (time.perf_counter() - t0) * 1e3
    poses_h = torch.from_numpy(np.ascontiguousarray(scene.poses)).pin_memory()
    points_h = torch.from_numpy(np.ascontiguousarray(scene.points)).pin_memory()
    out_poses = torch.empty_like(poses_h).pin_memory()
    out_points = torch.empty_like(points_h).pin_memory()

    def barrier():
        torch.cuda.synchronize()

    # ---------------- warm-up: W real LM iterations (also builds the structure once)
    t0 = time.perf_counter()
    s_w = pb.solve(bench_options(api, warm))
    barrier()
    warm_ms = (time.perf_counter() - t0) * 1e3
    log(f"[bench] warm-up solve ({warm} it incl. structure analysis): {warm_ms:.1f} ms, cost "
        f"{s_w.initial_cost:.4e} -> {s_w.final_cost:.4e}")

    # ---------------- device-resident: K LM iterations from the initial estimate
    pb.set_parameters(poses_h, points_h)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = pb.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    summ = pb.solve(bench_options(api, args.steps))
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = pb.launch_count() - launches0
    clocks = sampler.stop()
    assert summ.iterations == args.steps, (summ.iterations, summ.message)
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3) / 1e6
    stages = {k: getattr(summ, f"time_{k}_ms") / args.steps for k in ("jacobian", "residual", "schur", "cholesky", "update")}

    # ---------------- K1 alone (the HBM-bound kernel): CUDA events inside the library, per launch
    pb.set_parameters(poses_h, points_h)
    k1_ms = []
    for _ in range(10):
        pb.evaluate_device(True, fetch=True)
        k1_ms.append(pb.stage_ms("jacobian"))
    k1 = statistics.median(k1_ms[3:])

    # ---------------- end to end through the C ABI with host buffers
    e2e_steps = args.steps

    def e2e_run():
        pb.set_parameters(poses_h, points_h)                     # H2D from pinned memory
        s = pb.solve(bench_options(api, e2e_steps))
        pb.get_parameters(out_poses, out_points)                 # D2H result
        return s

    barrier()
    t0 = time.perf_counter()
    s_e = e2e_run()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    h2d = (poses_h.numel() + points_h.numel()) * 8 / e2e_steps
    d2h = (poses_h.numel() + points_h.numel()) * 8 / e2e_steps + 7 * 8 + 8

    peak, peak_src = load_peaks()
    bpo = k1_bytes_per_obs(scene)
    achieved = n_total * bpo / (k1 * 1e-3) / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": "M evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "step": STEP_DESC,
                   "l2": "no flush needed: every iteration streams > 3 GB through HBM (J alone is 1.2 GB > 126 MB L2)",
                   "parallelism": f"single GPU" if world == 1 else f"camera-range shards x{world}",
                   "setup_ms": {"scene_upload": upload_ms, "warmup_solve_incl_structure": warm_ms}},
        "lm_iters_per_sec": 1e3 / ms_per_step,
        "lm": {"iterations": summ.iterations, "successful_steps": summ.num_successful_steps,
               "jacobian_evals": summ.num_jacobian_evaluations, "residual_evals": summ.num_residual_evaluations,
               "initial_cost": summ.initial_cost, "final_cost": summ.final_cost},
        "stage_ms_per_step": stages,
        "roofline": {"bound": "hbm", "kernel": "k1_kernel<true>", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "bytes_per_obs": bpo, "kernel_ms": k1,
                     "k1_only_M_evals_per_s": n_total / (k1 * 1e-3) / 1e6},
        "e2e": {"value": n_total / (e2e_ms * 1e-3) / 1e6, "unit": "M evals/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                "call": "set_parameters(pinned host) + rsba_cuda_solve(K iterations) + get_parameters(pinned host)"},
        "gpu_launches": launches, "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        base, _ = cpu_reference_rate(scene)
        out["cpu_baseline"] = base
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
