#!/usr/bin/env python3
"""bench.py -- correspondence + linearisation + reduction throughput of the small_gicp hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--points P]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): synthetic room pair, P = 1,000,000 target points and 1,000,000 source
points PER GPU, GICP factor, DistanceRejector(1 m^2), poses = the Gauss-Newton trajectory from identity
(<= 20 iterations).  One "step" = one Reduction::linearize over the rank's source points at the next pose of that
trajectory (+ for N > 1 the all-reduce of H|b|e).  Metric: Mpoints/s per iteration = source points of all ranks /
max-over-ranks device time of a step.

`value`      inputs resident in HBM, CUDA events on the launching stream, L2 flushed between steps.
`e2e`        the same step through the C-ABI with HOST (pinned) buffers: source upload (H2D) + linearize + 344 B D2H.
`roofline`   algorithmic bytes (GICP: 100 B / source point) / kernel time vs the measured HBM copy peak.
`cpu_baseline` / --impl reference: the CPU oracle's OpenMP restatement of reduction_omp.hpp on this box's cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "correspondence+reduction Mpoints/sec per GICP iter"
UNIT = "Mpoints/s"
BYTES_PER_POINT = {"ICP": 36, "PLANE_ICP": 52, "GICP": 100, "VGICP": 116}
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--covs", default="knn", choices=["knn", "analytic"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--leaf", type=int, default=0, help="max leaf size of the device kd-tree (0 = library default)")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# inputs
# ------------------------------------------------------------------------------------------------
def make_inputs(points, rank, covs_mode, estimator=None):
    """target (shared by all ranks) and this rank's source shard; covariances as 4x4 zero-padded doubles.
    covs_mode "knn": the reference recipe (k = 20 neighbours, eigenvalues -> (1e-3, 1, 1), normal_estimation.hpp:28-92)
    evaluated by `estimator(points4) -> covs` (our arm: the device estimator; reference arm: the oracle's)."""
    from small_gicp_b200 import synthetic as syn

    world = syn.make_world(points, 42)
    Tgt = syn.gt_transform()
    tgt, tn = syn.sample_cloud(world, points, 43, return_normals=True)
    src_w, sn_w = syn.sample_cloud(world, points, 44 + rank, return_normals=True)
    Ti = np.linalg.inv(Tgt)
    src = src_w @ Ti[:3, :3].T + Ti[:3, 3]
    sn = sn_w @ Ti[:3, :3].T
    tgt = tgt.astype(np.float32).astype(np.float64)
    src = src.astype(np.float32).astype(np.float64)
    tgt4 = np.concatenate([tgt, np.ones((len(tgt), 1))], axis=1)
    src4 = np.concatenate([src, np.ones((len(src), 1))], axis=1)
    if covs_mode == "knn" and estimator is not None:
        tcov, scov = estimator(tgt4), estimator(src4)
        desc = "k=20 nearest-neighbour covariances, eigenvalues regularised to (1e-3,1,1) (reference recipe)"
    else:
        tcov, scov = syn.plane_covariances(tn), syn.plane_covariances(sn)
        desc = "analytic plane covariances I-(1-1e-3)nn^T from the generator's face normals"
    return {"target": tgt4, "source": src4, "target_covs": tcov, "source_covs": scov, "T_gt": Tgt, "covs": desc}


def se3_exp(a):
    """util/lie.hpp:73-96 (host side of the optimizer; numpy)"""
    w, t = a[:3], a[3:]
    th2 = w @ w
    th = np.sqrt(th2)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    T = np.eye(4)
    if th < 1e-10:
        R = np.eye(3) + K
        V = np.eye(3)
    else:
        R = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th2 * (K @ K)
        V = np.eye(3) + (1 - np.cos(th)) / th2 * K + (th - np.sin(th)) / (th2 * th) * (K @ K)
    T[:3, :3] = R
    T[:3, 3] = V @ t
    return T


def gn_trajectory(linearize, max_iter=20, lam=1e-6):
    """GaussNewtonOptimizer::optimize (optimizer.hpp:24-63): returns the poses at which linearize was called."""
    T = np.eye(4)
    poses = []
    for _ in range(max_iter):
        poses.append(T.copy())
        H, b, e = linearize(T)
        d = np.linalg.solve(H + lam * np.eye(6), -b)
        T = T @ se3_exp(d)
        if np.linalg.norm(d[:3]) <= 0.1 * np.pi / 180 and np.linalg.norm(d[3:]) <= 1e-3:
            break
    return poses, T


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_first_sample(self, timeout_s):
        """block until nvidia-smi has written its first line (it needs ~0.1 s to start)"""
        t0 = time.perf_counter()
        while self.proc and time.perf_counter() - t0 < timeout_s:
            try:
                if os.path.getsize(self.path) > 0:
                    return True
            except OSError:
                pass
            time.sleep(0.02)
        return False

    def stop(self):
        res = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return res
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            res = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}
        return res


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle's OpenMP reduction on this box's host cores
# ------------------------------------------------------------------------------------------------
def cpu_kind():
    """("reference", lib) when oracle/_ref exists -- the reference's OWN headers compiled in place against the Eigen API shim
    (oracle/ref_build; built in the dev container, shipped as a .so) -- else ("port", None): the oracle's restatement."""
    import oracle as O

    ref = O.reference_lib()
    return ("reference", ref) if ref is not None else ("port", None)


CPU_NOTE = {
    "reference": "small_gicp's own headers (ParallelReductionOMP<GICPFactor>, schedule(guided,8)) compiled against an Eigen API shim: the reference's code, but without Eigen's SIMD kernels; -O3, no -march=native",
    "port": "oracle restatement of reduction_omp.hpp (schedule(guided,8)); not the reference binary, no Eigen SIMD; -O3, no -march=native",
}


def cpu_setup(inp, threads):
    import oracle as O

    kind, lib = cpu_kind()
    tc = O.Cloud(inp["target"], _lib=lib)
    sc = O.Cloud(inp["source"], _lib=lib)
    t0 = time.perf_counter()
    tt = O.KdTree(tc)
    build_s = time.perf_counter() - t0
    tc.set_features(None, inp["target_covs"])
    sc.set_features(None, inp["source_covs"])
    reg = O.Registration(factor=O.FACTOR_GICP, rejector=O.REJECT_DISTANCE, max_dist_sq=1.0, num_threads=threads, _lib=lib)
    return O, tc, tt, sc, reg, build_s


def host_threads():
    """All the host threads the CPU path can use.  torchrun exports OMP_NUM_THREADS=1 and torch pins OpenMP to the physical
    core count, so the count is passed explicitly (the oracle's parallel regions carry a num_threads() clause, like
    reduction_omp.hpp:36)."""
    return max(1, os.cpu_count() or 1)


def best_cpu_registration(inp):
    """Time one linearize with all logical CPUs and with half of them (= physical cores on an SMT-2 host) and keep the faster."""
    best = None
    n = host_threads()
    for threads in sorted({n, max(1, n // 2)}, reverse=True):
        O, tc, tt, sc, reg, build_s = cpu_setup(inp, threads)
        reg.linearize(tc, tt, sc, np.eye(4))
        t0 = time.perf_counter()
        reg.linearize(tc, tt, sc, np.eye(4))
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, threads, (O, tc, tt, sc, reg, build_s))
    return best[1], best[2]


def run_reference(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return 0
    os.environ["OMP_NUM_THREADS"] = str(host_threads())  # before libgomp is loaded (torchrun sets it to 1)
    import oracle as O

    threads = host_threads()

    def oracle_covs(p4):  # the reference arm prepares its inputs with the CPU oracle only (none of our kernels on this arm)
        c = O.Cloud(p4, _lib=cpu_kind()[1])
        t = O.KdTree(c)
        t.estimate(20, O.FEAT_COV, threads)
        return c.covs

    inp = make_inputs(args.points, 0, args.covs, oracle_covs)
    threads, (O, tc, tt, sc, reg, build_s) = best_cpu_registration(inp)
    poses, _ = gn_trajectory(lambda T: reg.linearize(tc, tt, sc, T))
    for i in range(args.warmup):
        reg.linearize(tc, tt, sc, poses[i % len(poses)])
    t0 = time.perf_counter()
    for i in range(args.steps):
        reg.linearize(tc, tt, sc, poses[i % len(poses)])
    dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    value = args.points / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, inp),
        "cpu_baseline": {
            "value": value,
            "unit": UNIT,
            "cores": threads,
            "kind": cpu_kind()[0],
            "sample": f"{args.steps} x Reduction::linearize over all {args.points} source points; " + CPU_NOTE[cpu_kind()[0]],
            "nproc": os.cpu_count(),
            "kdtree_build_s": build_s,
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, inp):
    return {
        "workload": f"BASELINE configs[1]: {args.points}-pt synthetic room pair per GPU, GICP factor, DistanceRejector(1.0), GaussNewton trajectory from identity (<=20 iters), one step = one linearize",
        "points_target": args.points,
        "points_source_per_gpu": args.points,
        "covariances": inp["covs"],
        "l2": "flushed between timed steps (256 MiB write)",
        "parallelism": f"source sharded over {args.gpus} GPU(s), target + kd-tree replicated, all-reduce of H|b|e (44 doubles) per step (see allreduce)" if args.gpus > 1 else "single GPU",
    }


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    import small_gicp_b200 as sg

    rank, local_rank, world = dist_env()
    if world != args.gpus and world != 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    ctx = sg.Context(local_rank)
    inp = make_inputs(args.points, rank, args.covs, lambda p4: ctx.estimate_features(p4, 20, normals=False)[1])
    n_src = inp["source"].shape[0]
    # one explicit (non-default) stream for everything: the context's kernels, torch's fills / events and NCCL all
    # run on it, so CUDA events recorded on it bracket exactly the work being timed
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    # N > 1: the all-reduce of H|b|e is fused into the reduction kernel (peer mailboxes over NVLink); SGB_FUSED_ALLREDUCE=0
    # selects the NCCL all_reduce of the 44 doubles instead (A/B)
    fused = False
    if use_dist and os.environ.get("SGB_FUSED_ALLREDUCE", "1") != "0":
        from small_gicp_b200.distributed import connect_fused

        fused = connect_fused(ctx)
    nccl = use_dist and not fused
    ctx.set_target(inp["target"], None, inp["target_covs"])
    torch.cuda.synchronize()
    t_build = time.perf_counter()
    ctx.build_target_kdtree(args.leaf)  # device kd-tree + block lists of the grid front end
    ctx.synchronize()
    target_build_ms = (time.perf_counter() - t_build) * 1e3
    # pinned host copies of the step's inputs for the e2e leg
    src_pin = torch.from_numpy(inp["source"]).pin_memory()
    cov_pin = torch.from_numpy(inp["source_covs"]).pin_memory()
    ctx.set_source(src_pin.numpy(), cov_pin.numpy())
    out = torch.zeros(64, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    if use_dist:
        dist.barrier()  # the ranks prepared their shards at different speeds: enter the first collective linearize together

    def step_device(T):
        ctx.linearize_device(T, out.data_ptr(), factor=sg.FACTOR_GICP, rejector=sg.REJECT_DISTANCE, max_dist_sq=1.0)
        if nccl:
            dist.all_reduce(out[:44])

    def linearize_host(T):
        step_device(T)
        h = out[:44].cpu().numpy()
        return h[:36].reshape(6, 6), h[36:42], h[42]

    poses, T_final = gn_trajectory(linearize_host)
    err = np.linalg.inv(inp["T_gt"]) @ T_final
    rot_err = float(np.arccos(np.clip((np.trace(err[:3, :3]) - 1) / 2, -1, 1)))
    trans_err = float(np.linalg.norm(err[:3, 3]))

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs, per-step CUDA events, L2 flushed between steps ----
    for i in range(args.warmup):
        flush.zero_()
        step_device(poses[i % len(poses)])
    barrier()
    # Clocks: nvidia-smi needs ~0.1 s to start and samples every 0.1 s, the K timed steps take ~10 ms.  The sampler therefore runs
    # over the timed steps PLUS the same step repeated untimed before and after them (identical load, ~0.5 s each side), so that
    # the reported clocks / throttle reasons are those of the GPU under exactly this load.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample(1.5)
    barrier()  # rank 0 waited for the sampler: every rank enters the first collective step of the roll together
    roll_steps = int(min(2000, max(20, 2_000_000_000 // max(1, n_src))))  # the same count on every rank (the step holds a collective)
    if os.environ.get("SGB_BENCH_ROLL"):  # profiler runs: a short roll keeps the launch numbering simple (ncu -s / -c)
        roll_steps = max(1, int(os.environ["SGB_BENCH_ROLL"]))

    def load_roll():
        for i in range(roll_steps):
            flush.zero_()
            step_device(poses[i % len(poses)])
            if i % 64 == 63:
                stream.synchronize()  # keep the launch queue short
        barrier()

    load_roll()
    launches0 = ctx.kernel_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.zero_()
        T = poses[i % len(poses)]
        ev[i][0].record(stream)
        kev[i][0].record(stream)
        ctx.linearize_device(T, out.data_ptr(), factor=sg.FACTOR_GICP, rejector=sg.REJECT_DISTANCE, max_dist_sq=1.0)
        kev[i][1].record(stream)
        if nccl:
            dist.all_reduce(out[:44])
        ev[i][1].record(stream)
    barrier()
    launches = ctx.kernel_launches - launches0
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    kern_ms = np.array([a.elapsed_time(b) for a, b in kev])
    total_ms = float(step_ms.sum())
    # the same steps grouped by the pose of the trajectory they ran at (misaligned first iterations walk the tree, converged ones settle in the grid)
    per_pose_ms = [float(np.mean([kern_ms[i] for i in range(args.steps) if i % len(poses) == k])) if k < args.steps else None for k in range(len(poses))]
    # ---- warm-L2 variant (what consecutive optimiser iterations actually see), informational ----
    barrier()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record(stream)
    for i in range(args.steps):
        step_device(poses[i % len(poses)])
    w1.record(stream)
    barrier()
    warm_ms = w0.elapsed_time(w1) / args.steps
    load_roll()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = f"the {args.steps} timed steps + {roll_steps} identical untimed steps before and after them"

    # ---- e2e: host buffers through the C-ABI, H2D of the step's inputs + D2H of H|b|e inside the timed region ----
    h_out = torch.zeros(64, dtype=torch.float64).pin_memory()

    def step_e2e(T):
        ctx.set_source(src_pin.numpy(), cov_pin.numpy())  # H2D 160 B / point (reference layout: Vector4d + Matrix4d)
        step_device(T)
        h_out[:44].copy_(out[:44], non_blocking=True)
        stream.synchronize()

    for i in range(2):
        step_e2e(poses[i % len(poses)])
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(e2e_steps):
        step_e2e(poses[i % len(poses)])
    e1.record(stream)
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    e2e_dev_ms = e0.elapsed_time(e1) / e2e_steps
    e2e_ms = max(e2e_wall_ms, e2e_dev_ms)

    # ---- informational: one whole align() through the host-facing calls -- source uploaded ONCE from pinned host memory, then every
    # Gauss-Newton iteration = pose in (128 B), linearize, H|b|e out (352 B) + host solve: what Registration<>::align does with the
    # ParallelReductionCUDA glue, whose device mirror of the source is reused by all iterations of an align() ----
    def align_e2e():
        ctx.set_source(src_pin.numpy(), cov_pin.numpy())
        p, _ = gn_trajectory(linearize_host)
        return len(p)

    align_e2e()
    barrier()
    t0 = time.perf_counter()
    align_reps, align_iters = 3, 0
    for _ in range(align_reps):
        align_iters += align_e2e()
    torch.cuda.synchronize()
    align_ms = (time.perf_counter() - t0) * 1e3 / align_reps
    align_iters //= align_reps

    # ---- max over ranks ----
    stats = torch.tensor([total_ms, float(kern_ms.mean()), warm_ms, e2e_ms, align_ms], dtype=torch.float64, device=dev)
    if use_dist:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    total_ms, kern_ms_mean, warm_ms, e2e_ms, align_ms = [float(x) for x in stats.cpu()]
    total_points = n_src * world
    ms_per_step = total_ms / args.steps
    value = total_points / (ms_per_step * 1e-3) / 1e6

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as O

        threads, (_, tc, tt, sc, reg, build_s) = best_cpu_registration(inp)
        reg.linearize(tc, tt, sc, poses[0])
        reps = 0
        t0 = time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < len(poses)):
            H0, b0, e0c = reg.linearize(tc, tt, sc, poses[reps % len(poses)])
            reps += 1
        cpu_ms = (time.perf_counter() - t0) / reps * 1e3
        # parity of the sums at the last CPU pose, on the same inputs
        Hg, bg, eg = linearize_host(poses[(reps - 1) % len(poses)])
        cpu_baseline = {
            "value": args.points / (cpu_ms * 1e-3) / 1e6,
            "unit": UNIT,
            "cores": threads,
            "kind": cpu_kind()[0],
            "sample": f"{reps} x linearize over all {args.points} source points at the GN poses; " + CPU_NOTE[cpu_kind()[0]],
            "nproc": os.cpu_count(),
            "ms_per_step": cpu_ms,
            "kdtree_build_s": build_s,
            "parity_rel_H": float(np.linalg.norm(Hg - H0) / np.linalg.norm(H0)),
            "parity_rel_e": float(abs(eg - e0c) / e0c),
        }

    if rank == 0:
        peak, peak_src = hbm_peak()
        alg_bytes = BYTES_PER_POINT["GICP"] * n_src
        achieved = alg_bytes / (kern_ms_mean * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "linearize_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32 search / f64 factor algebra + sums",
            "data": "synthetic",
            "config": dict(workload_config(args, inp), allreduce=("none (single GPU)" if not use_dist else "fused into factor_reduce_kernel: peer mailboxes over NVLink (sgb_comm_*)" if fused else "NCCL all_reduce of 44 doubles after the kernel")),
            "clocks": clocks,
            "e2e": {
                "value": total_points / (e2e_ms * 1e-3) / 1e6,
                "unit": UNIT,
                "h2d_bytes_per_step": int(n_src * 160 + 128),
                "d2h_bytes_per_step": 44 * 8,
                "ms_per_step": e2e_ms,
                "what": "sgb_source_set_points (pinned host Vector4d+Matrix4d layout, device conversion + Hilbert sort) + sgb_linearize_device + D2H of H|b|e, per step",
            },
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "hbm",
                "kernel": "one sgb_linearize = sgb::grid_probe_blocks_kernel + pending_search_kernel | packet_search_kernel (picked on the device) + factor_reduce_kernel<2,0>: four launches, timed together",
                "achieved": achieved,
                "peak": peak,
                "peak_source": peak_src,
                "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": traffic,
                "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms": kern_ms_mean,
                "launches_per_step": launches / max(1, args.steps),
            },
            "cpu_baseline": cpu_baseline,
            "value_l2_warm": total_points / (warm_ms * 1e-3) / 1e6,
            "per_pose_ms": per_pose_ms,
            "e2e_align": {
                "value": total_points * align_iters / (align_ms * 1e-3) / 1e6,
                "unit": UNIT,
                "ms_per_align": align_ms,
                "gn_iterations": align_iters,
                "h2d_bytes_per_align": int(n_src * 160 + 128 * align_iters),
                "d2h_bytes_per_align": 44 * 8 * align_iters,
                "what": "informational: whole GaussNewton align() through the host-facing calls, wall clock -- source uploaded once (pinned host, reference layout), then per iteration pose in / linearize / H|b|e out / host 6x6 solve; Mpoints/s = points x iterations / time",
            },
            "setup": {"target_build_ms": target_build_ms, "what": "sgb_target_build_kdtree: device kd-tree construction + block lists + hash table, first call (includes allocations)"},
            "pose_error_vs_gt": {"rot_rad": rot_err, "trans_m": trans_err, "gn_iterations": len(poses)},
        }
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.barrier()  # nobody unmaps a mailbox a peer may still write to
    ctx.close()
    if use_dist:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
