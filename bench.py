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

# Inside the LM loop the residual + Jacobian evaluation is the fused point pass (k2_fused.cu): per observation it reads
# xy (16) and three point-major indices (12) and writes the compact Jacobian record (96: jr, jx of the interpolated
# pose, from which J_pose0 / J_pose1 / J_point follow with tau), tau (8) and the observation's 12 x 3 Schur panel
# rows (288); per point it reads X (24) + the CSR pointer (4) and writes C, g, Cinv, L^-1, t, D^2, (X | t) (288).
# The 240-byte Ceres layout of SURVEY 8d is what rsba_cuda_evaluate writes (K1_FULL_BYTES_FIXED; roofline_k1_full).
K1_BYTES_FIXED = 16 + 12 + 96 + 8 + 288
K1_BYTES_PER_POINT = 24 + 4 + 288
K1_FULL_BYTES_FIXED = 16 + 8 + 16 + 240 + 1


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner,
# library chatter) is sent to stderr; emit() writes the result to the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def k1_bytes_per_obs(scene, full=False) -> float:
    if full:
        return K1_FULL_BYTES_FIXED + (24.0 * scene.num_points + 96.0 * scene.num_frames) / scene.num_obs
    return K1_BYTES_FIXED + (K1_BYTES_PER_POINT * scene.num_points + 96.0 * scene.num_frames) / scene.num_obs


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
                                       "-lms", "50", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None
            return
        t0 = time.time()                       # nvidia-smi needs a moment before its first sample
        while time.time() - t0 < 5.0 and os.path.getsize(self.f.name) == 0:
            time.sleep(0.02)

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


class CpuLmLoop:
    """The reference's CPU path for one LM iteration, on all host cores:
    residual+Jacobian by the reference's own functor under Jet<15> forward autodiff (oracle/_ref:
    reference headers compiled verbatim; the plain-C port only if the prebuilt library did not
    travel), Schur elimination + back-substitution by oracle/cpu_lm.cc (C++/OpenMP restatement of
    Ceres' SchurEliminator), reduced system by LAPACK dpbsv (band Cholesky, stand-in for CHOLMOD),
    cost-only evaluation at the trial point, accept/reject -- the same step the GPU arm times."""

    def __init__(self, scene, nthreads=0):
        import oracle
        from oracle import cpu_lm
        self.oracle, self.scene = oracle, scene
        self.cores = nthreads or (os.cpu_count() or 1)
        # torchrun exports OMP_NUM_THREADS=1: the OpenMP parts get their thread count explicitly
        # (omp_set_num_threads inside the libraries), LAPACK/OpenBLAS through threadpoolctl
        try:
            import threadpoolctl
            self._blas_limit = threadpoolctl.threadpool_limits(limits=self.cores)
        except Exception:  # noqa: BLE001
            self._blas_limit = None
        self.n = scene.num_obs
        self.res, self.J = np.zeros((self.n, 2)), np.zeros((self.n, 30))
        self.res_t = np.zeros((self.n, 2))
        self.valid = np.zeros(self.n, np.uint8)
        if oracle.ref_available():
            self.kind = "reference"
            self.pb = oracle.RefProblem(scene)
        else:
            self.kind = "port"
            self.pb = None
        self.lm = cpu_lm.CpuLm(scene, nthreads=self.cores)
        self.k1_ms = []

    def _eval(self, poses, points, jac):
        if self.pb is not None:
            out = self.res if jac else self.res_t
            self.pb.eval(poses, points, out, self.J if jac else None, self.valid, self.cores)
            return out
        r, J, v = self.oracle.evaluate(self.scene, poses, points, jac=jac, impl="port", nthreads=self.cores)
        if jac:
            self.J = J
        self.valid = v
        return r

    def solve(self, iters):
        """`iters` LM iterations from the scene's initial estimate; returns (seconds, final cost)."""
        sc = self.scene
        poses, points = np.ascontiguousarray(sc.poses).copy(), np.ascontiguousarray(sc.points).copy()
        radius, decrease = 1e4, 2.0
        t_begin = time.perf_counter()
        t0 = time.perf_counter()
        r = self._eval(poses, points, True)
        self.k1_ms.append((time.perf_counter() - t0) * 1e3)
        cost = 0.5 * float(np.sum(r * r))
        first = True
        for _ in range(iters):
            st = self.lm.step(self.J, r, radius, compute_scale=first)
            first = False
            tp, tq = poses + st["delta_poses"], points + st["delta_points"]
            rt = self._eval(tp, tq, False)
            new_cost = 0.5 * float(np.sum(rt * rt)) if self.valid.all() else np.inf
            mcc = st["model_cost_change"]
            rho = (cost - new_cost) / mcc if mcc > 0 else -1.0
            if rho > 1e-3:
                poses, points = tp, tq
                radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
                decrease = 2.0
                t0 = time.perf_counter()
                r = self._eval(poses, points, True)
                self.k1_ms.append((time.perf_counter() - t0) * 1e3)
                cost = 0.5 * float(np.sum(r * r))
            else:
                radius /= decrease
                decrease *= 2.0
        return time.perf_counter() - t_begin, cost


def cpu_lm_baseline(scene, iters, warmup, nthreads=0):
    loop = CpuLmLoop(scene, nthreads)
    if warmup:
        loop.solve(warmup)
    loop.k1_ms = []
    secs, cost = loop.solve(iters)
    n = scene.num_obs
    sample = (f"{iters} LM iterations on {scene.name or 'the scene'} ({scene.num_frames} frames / {scene.num_points} points / "
              f"{n} observations) after {warmup} warm-up iterations; functor+Jet<15> autodiff (oracle/_ref), "
              f"Schur elimination in C++/OpenMP (oracle/cpu_lm.cc), LAPACK dpbsv band Cholesky, all host cores")
    return {"value": n / (secs / iters) / 1e6, "unit": "M evals/s", "cores": int(loop.cores), "kind": loop.kind,
            "sample": sample, "ms_per_iteration": secs / iters * 1e3, "final_cost": cost,
            "k1_only_M_evals_per_s": n / (statistics.median(loop.k1_ms) * 1e-3) / 1e6,
            "stage_ms_last_iteration": loop.lm.times}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the same LM iteration on the host
    cores (rank 0 only; the other ranks exit without work)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rsba_b200.scene import make_config
    scene = make_config(args.config)
    base = cpu_lm_baseline(scene, args.steps, args.warmup)
    out = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "M evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_iteration"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.config), "sample": base["sample"],
        "lm_iters_per_sec": 1e3 / base["ms_per_iteration"],
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "M evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


METRIC = "M residual+Jacobian evals/sec (full LM iteration)"
WORKLOADS = {
    "C1": "C1: 10 frames / 500 points / 5k observations synthetic RS scene",
    "C2": "C2: 100 frames / 20k points / 500k observations synthetic RS scene",
    "C3": "C3: 1000 frames / 200k points / 5M observations synthetic RS scene",
    "C5": "C5: 4000 frames / 1M points / 20M observations synthetic RS scene",
    "C2dense": "C2dense: 100 frames / 20k points / 500k observations, every frame pair covisible (orbit): dense reduced system",
    "C3dense": "C3dense: 1000 frames / 200k points / 5M observations, every frame pair covisible (orbit): dense reduced system",
}
STEP_DESC = ("one Levenberg-Marquardt iteration of rsba_cuda_solve: residual+Jacobian of every observation, block normal "
             "equations + Schur complement (K2), tile Cholesky + solves (K3), back-substitution/step (K4), "
             "cost at the trial point (K1r), accept/reject")


def bench_options(api, iters):
    """Exactly `iters` LM iterations: tolerances zeroed so the loop cannot stop early."""
    return api.default_options(max_num_iterations=iters, function_tolerance=0.0, gradient_tolerance=0.0,
                               parameter_tolerance=0.0)


FP64_PEAK_FALLBACK_TFLOPS = 37.1   # profiles/r01_fp64_peak.txt; only used if the in-run measurement fails


def fp64_peak(api, device):
    """FP64 peak of THIS device, measured in this run by the library itself (rsba_cuda_measure_fp64_peak:
    register-resident DFMA chains and mma.sync.m8n8k4.f64 chains; MEASURED_PEAKS.json has no FP64 figure)."""
    try:
        dfma, dmma = api.measure_fp64_peak(device)
        peak = max(dfma, dmma)
        if peak > 1.0:
            return peak, (f"measured in this run: DFMA {dfma:.2f}, DMMA {dmma:.2f} TFLOP/s "
                          "(rsba_cuda_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure")
    except Exception as e:  # noqa: BLE001
        log(f"[bench] FP64 peak measurement failed: {e}")
    return FP64_PEAK_FALLBACK_TFLOPS, "fallback: profiles/r01_fp64_peak.txt (in-run measurement failed)"


def dram_traffic(kernel, config, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the newest committed
    `ncu --set full` summary (profiles/traffic.json: {kernel: {config, bytes, source, commit}}); None when there is
    no capture of this kernel on this workload."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        # template instances are keyed with their arguments ("schur_syrk_kernel<8, 2, 5>"): match by prefix
        key = kernel if kernel in table else next(k for k in sorted(table) if k.startswith(kernel + "<"))
        t = table[key]
    except (OSError, KeyError, ValueError, StopIteration):
        return None, None
    if t.get("config") != config or world != 1:
        return None, None
    return float(t["bytes"]), f'{t["source"]} @ {t.get("commit", "?")}'


def schur_algorithmic_flops(scene) -> float:
    """per point with k observations  Y = E C^-1 (12k*9*2)  +  S -= Y E^T on the symmetric
    half ((12k)(12k+1)/2 * 3 * 2)."""
    k = np.bincount(scene.obs_point, minlength=scene.num_points).astype(np.float64)
    return float(np.sum(12 * k * 9 * 2 + (12 * k) * (12 * k + 1) / 2 * 3 * 2))


def band_cholesky_flops(scene) -> float:
    """n b^2 for the banded reduced system (b = 12 * (largest frame span of a point + 1))."""
    order = np.argsort(scene.obs_point, kind="stable")
    fr, pt = scene.obs_frame[order], scene.obs_point[order]
    first = np.r_[True, pt[1:] != pt[:-1]]
    last = np.r_[first[1:], True]
    span = int(np.max(fr[last] - fr[first])) if fr.size else 0
    n, b = 12.0 * scene.num_frames, 12.0 * (span + 1)
    return min(n * b * b, n ** 3 / 3.0)


def run_ours(args):
    import torch
    import rsba_b200.api as api
    from rsba_b200.scene import make_config

    rank, local, world = dist_setup(args.gpus)
    t0 = time.time()
    scene = make_config(args.config)
    if rank == 0:
        log(f"[bench] scene {args.config}: {scene.num_obs} obs ready in {time.time() - t0:.1f}s")
    n_total = scene.num_obs
    warm = max(args.warmup, 3)

    pb = api.Problem(local)
    stream = torch.cuda.Stream()            # not the legacy default stream: the library captures CUDA graphs
    torch.cuda.set_stream(stream)
    pb.set_stream(stream.cuda_stream)
    if world > 1:
        import torch.distributed as dist
        uid = [api.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        pb.comm_init(rank, world, uid[0])
    t0 = time.perf_counter()
    pb.load_scene(scene)
    upload_ms = (time.perf_counter() - t0) * 1e3
    poses_h = torch.from_numpy(np.ascontiguousarray(scene.poses)).pin_memory()
    points_h = torch.from_numpy(np.ascontiguousarray(scene.points)).pin_memory()
    out_poses = torch.empty_like(poses_h).pin_memory()
    out_points = torch.empty_like(points_h).pin_memory()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- warm-up: W real LM iterations (also builds the structure once)
    t0 = time.perf_counter()
    s_w = pb.solve(bench_options(api, warm))
    barrier()
    warm_ms = (time.perf_counter() - t0) * 1e3
    if rank == 0:
        log(f"[bench] warm-up solve ({warm} it incl. structure analysis): {warm_ms:.1f} ms, cost "
            f"{s_w.initial_cost:.4e} -> {s_w.final_cost:.4e}")

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---------------- device-resident: K LM iterations from the initial estimate
    pb.set_parameters(poses_h, points_h)
    launches0 = pb.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    summ = pb.solve(bench_options(api, args.steps))
    ev1.record(stream)
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = pb.launch_count() - launches0
    assert summ.iterations == args.steps, (summ.iterations, summ.message)
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3) / 1e6
    stages = {k: getattr(summ, f"time_{k}_ms") / args.steps
              for k in ("jacobian", "residual", "schur", "cholesky", "update", "allreduce")}
    # RSBA_CUDA_FINE_TIMERS=1 (diagnostic): the per-kernel events are also recorded inside rsba_cuda_solve
    inloop = {k: round(pb.stage_ms(k), 4) for k in api.STAGES} if os.environ.get("RSBA_CUDA_FINE_TIMERS") else None

    # ---------------- per-kernel times: CUDA events inside the library on the launching stream
    pb.set_parameters(poses_h, points_h)
    names = ("jacobian", "point_blocks", "frame_blocks", "phi_build", "schur_syrk", "schur_reduce", "cholesky", "factor",
             "tri_solve", "update")
    samples = {k: [] for k in names}
    for _ in range(7):
        pb.linearize_and_step(1e4, bench_options(api, 1), want_S=False, fetch=False)
        for k in names:
            samples[k].append(pb.stage_ms(k))
    kern = {k: statistics.median(v[2:]) for k, v in samples.items()}
    # the reference-facing evaluation (rsba_cuda_evaluate_device): residuals + the full 2 x 30 Jacobian per observation
    full = []
    for _ in range(7):
        pb.evaluate_device(with_jacobian=True, fetch=False)
        full.append(pb.stage_ms("jacobian"))
    kern["jacobian_full"] = statistics.median(full[2:])

    # ---------------- end to end through the C ABI with host buffers
    def e2e_run():
        pb.set_parameters(poses_h, points_h)                     # H2D from pinned memory
        s = pb.solve(bench_options(api, args.steps))
        pb.get_parameters(out_poses, out_points)                 # D2H result
        return s

    barrier()
    t0 = time.perf_counter()
    e2e_run()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    # ---------------- the same, cold: a fresh handle pays the scene upload (120 MB of observations at C3) and the
    # one-off structure analysis (ordering, symbolic factorisation, pair lists) -- what ONE BA call of the
    # reference's driver would see (CeresHandler::Add for every frame + ceres::Solve's preprocessing)
    # Twice: the first fresh handle also grows the device-memory pool (the bench's own handle is still alive and holds
    # its 3 GB); the second finds the blocks the first one freed -- the steady state of a driver that builds one handle
    # per BA call, as the reference does (VideoSfMHandler.cc:585).
    cold_ms = cold_first_ms = None
    if world == 1:
        for rep in range(2):
            barrier()
            t0 = time.perf_counter()
            with api.Problem(local) as pc:
                pc.set_stream(stream.cuda_stream)
                pc.load_scene(scene)
                pc.solve(bench_options(api, args.steps))
                pc.get_parameters(out_poses, out_points)
            barrier()
            cold_ms = (time.perf_counter() - t0) * 1e3
            if rep == 0:
                cold_first_ms = cold_ms
    clocks = sampler.stop() if rank == 0 else None
    # ---------------- parity of the sharded solve with the single-GPU one (outside every timed region): rank 0
    # re-solves the same K iterations from the same initial estimate on its own GPU alone and compares
    parity = None
    if world > 1:
        pb.set_parameters(poses_h, points_h)
        s_multi = pb.solve(bench_options(api, args.steps))
        pb.get_parameters(out_poses, out_points)
        if rank == 0:
            po_m, pt_m = out_poses.numpy().copy(), out_points.numpy().copy()
            with api.Problem(local) as p1:
                p1.set_stream(stream.cuda_stream)
                p1.load_scene(scene)
                s_one = p1.solve(bench_options(api, args.steps))
                po_1, pt_1 = p1.get_parameters()
            rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))  # noqa: E731
            parity = {"final_cost_rel": abs(s_multi.final_cost - s_one.final_cost) / s_one.final_cost,
                      "poses_rel": rel(po_m, po_1), "points_rel": rel(pt_m, pt_1),
                      "successful_steps": [int(s_multi.num_successful_steps), int(s_one.num_successful_steps)]}
        barrier()
    h2d = (poses_h.numel() + points_h.numel()) * 8 / args.steps
    d2h = (poses_h.numel() + points_h.numel()) * 8 / args.steps + 7 * 8 + 8
    if world > 1:
        import torch.distributed as dist
        pb.close()
        dist.destroy_process_group()
    if rank != 0:
        return

    hbm_peak, hbm_src = load_peaks()
    fp64_tf, fp64_src = fp64_peak(api, local)
    bpo = k1_bytes_per_obs(scene)
    k1 = kern["jacobian"]
    n_local = int(summ.num_residual_blocks)     # this rank's share of the observations (== n_total on one GPU)
    share = n_local / n_total
    tr_k1, tr_k1_src = dram_traffic("point_pass_kernel", args.config, world)
    roof_k1 = {"bound": "hbm", "kernel": "point_pass_kernel (residual + Jacobian of every observation, point blocks + inverse, "
                                          "Schur panel rows, point-major compact Jacobian records: the solver's linearisation)",
               "achieved": n_local * bpo / (k1 * 1e-3) / 1e9,
               "peak": hbm_peak, "unit": "GB/s", "traffic": tr_k1, "traffic_source": tr_k1_src, "peak_source": hbm_src,
               "algorithmic_bytes": n_local * bpo, "bytes_per_obs": bpo, "kernel_ms": k1,
               "k1_only_M_evals_per_s": n_local / (k1 * 1e-3) / 1e6, "observations_on_this_rank": n_local}
    roof_k1["frac"] = roof_k1["achieved"] / hbm_peak
    bpo_f, k1f = k1_bytes_per_obs(scene, full=True), kern["jacobian_full"]
    roof_k1_full = {"bound": "hbm", "kernel": "k1_kernel<true> (residual + full 2x30 Jacobian, rsba_cuda_evaluate)",
                    "achieved": n_local * bpo_f / (k1f * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "traffic": dram_traffic("k1_kernel<1, 0, 0>", args.config, world)[0], "peak_source": hbm_src, "algorithmic_bytes": n_local * bpo_f, "bytes_per_obs": bpo_f, "kernel_ms": k1f,
                    "k1_only_M_evals_per_s": n_local / (k1f * 1e-3) / 1e6}
    roof_k1_full["frac"] = roof_k1_full["achieved"] / hbm_peak
    fl = schur_algorithmic_flops(scene) * share       # rank 0's share of the points
    tr_sy, tr_sy_src = dram_traffic("schur_syrk_kernel", args.config, world)
    roof_syrk = {"bound": "tensor", "kernel": "schur_syrk_kernel (Schur complement, FP64 mma.sync m8n8k4)",
                 "achieved": fl / (kern["schur_syrk"] * 1e-3) / 1e12, "peak": fp64_tf, "unit": "TFLOP/s",
                 "traffic": tr_sy, "traffic_source": tr_sy_src, "peak_source": fp64_src, "algorithmic_flops": fl,
                 "kernel_ms": kern["schur_syrk"]}
    roof_syrk["frac"] = roof_syrk["achieved"] / fp64_tf
    # K3: factorisation + both substitutions are ONE persistent task-graph kernel (k3_dag.cu).  Algorithmic flops =
    # what the tile plan of this scene executes (sum over non-zero tiles after fill: potrf n^3/3, trsm n^3, update 2 n^3),
    # which is n b^2 for a banded (video) scene and n^3 / 3 for a fully covisible one.
    k3_ms = kern["cholesky"]
    cf = float(summ.tile_flops) if getattr(summ, "tile_flops", 0) else band_cholesky_flops(scene)
    tr_k3, tr_k3_src = dram_traffic("k3_dag_kernel", args.config, world)
    roof_chol = {"bound": "tensor", "kernel": "k3_dag_kernel (tile Cholesky + forward/backward substitution, one persistent kernel)",
                 "achieved": cf / (k3_ms * 1e-3) / 1e12, "peak": fp64_tf, "unit": "TFLOP/s", "traffic": tr_k3,
                 "traffic_source": tr_k3_src, "peak_source": fp64_src, "algorithmic_flops": cf, "kernel_ms": k3_ms,
                 "note": "a banded (video) system is a latency chain of dependent 96x96 panels, not a throughput problem; "
                         "a fully covisible one (--config C3dense) is the dense blocked algorithm"}
    roof_chol["frac"] = roof_chol["achieved"] / fp64_tf
    roofs = {"k1": roof_k1, "schur_syrk": roof_syrk, "cholesky": roof_chol}
    dominant = max(roofs, key=lambda k: roofs[k]["kernel_ms"])
    out = {
        "metric": METRIC, "value": value, "unit": "M evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step,
        # (kept early in the line: the driver's record of a scaling run is a truncated tail)
        "stage_ms_per_step": {k: round(v, 4) for k, v in stages.items()},
        "parity_vs_1gpu": parity,
        "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.config),
        "l2": "no flush needed: every iteration streams > 8 GB through HBM (ncu, C3: point pass 2.1 GB, SYRK 5.5 GB, back-substitution 0.6 GB; the Schur panels alone are 1.8 GB against 126 MB of L2)",
        "parallelism": "single GPU" if world == 1 else f"observations sharded by point owner x{world}, one NCCL allreduce of the reduced system per linear solve",
        "setup_ms": {"scene_upload": upload_ms, "warmup_solve_incl_structure": warm_ms},
        "lm_iters_per_sec": 1e3 / ms_per_step,
        "lm": {"iterations": summ.iterations, "successful_steps": summ.num_successful_steps,
               "jacobian_evals": summ.num_jacobian_evaluations, "residual_evals": summ.num_residual_evaluations,
               "initial_cost": summ.initial_cost, "final_cost": summ.final_cost},
        "kernel_ms": kern, "inloop_stage_ms_last_iteration": inloop,
        "roofline": dict(roofs[dominant], dominant=dominant),
        "roofline_k1": roof_k1, "roofline_k1_full": roof_k1_full, "roofline_schur": roof_syrk, "roofline_cholesky": roof_chol,
        "e2e": {"value": n_total / (e2e_ms * 1e-3) / 1e6, "unit": "M evals/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                "call": "rsba_cuda_set_parameters(pinned host) + rsba_cuda_solve(K iterations) + rsba_cuda_get_parameters(pinned host)",
                "cold_call_ms_total": cold_ms, "cold_call_ms_first_in_process": cold_first_ms,
                "cold_call": "fresh handle: rsba_cuda_create + set_camera + set_scene (observations H2D) + set_parameters + "
                             "solve(K iterations, incl. structure analysis) + get_parameters + destroy"},
        "gpu_launches": launches, "clocks": clocks,
        "denominator_note": "the reference arm / cpu_baseline is the reference's own functor under Jet<15> (oracle/_ref) PLUS "
                            "our CPU restatement of Ceres' Schur elimination / LM loop (oracle/cpu_lm.cc, LAPACK dpbsv); "
                            "it is not Ceres itself, which cannot be built here",
    }
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_lm_baseline(scene, args.cpu_iters, 1)
    emit(out)


def bench_config(name):
    """The `config` object: identical in both arms (ours / --impl reference)."""
    return {"workload": WORKLOADS[name], "step": STEP_DESC}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-iters", type=int, default=5, help="LM iterations of the bounded CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
