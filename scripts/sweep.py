#!/usr/bin/env python3
"""BASELINE.json configs[4]: kNN + reduction throughput over cloud sizes on ONE GPU.

  python scripts/sweep.py [--sizes 100000,300000,1000000,3000000,10000000] [--reps 10] [--out gpurun_out/sweep.json]

Per size N (N target points, N source points, density kept constant: the room grows with sqrt(N)): one `linearize` at the
iteration-0 pose (identity: misaligned, the tree walk dominates) and one at the converged pose of the Gauss-Newton
trajectory (the grid front end settles almost every query), for the GICP and the ICP factor, DistanceRejector(1.0).
Timing: CUDA events on the launching stream, 256 MiB L2 flush before every timed call, mean of --reps calls after 3
warm-ups.  Reported next to it: the algorithmic-byte roofline fraction (100 B / 36 B per source point over the measured HBM peak).
Writes one JSON document and prints a markdown table."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench as B  # noqa: E402  (inputs, Gauss-Newton trajectory, HBM peak: the same definitions as the bench line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="100000,300000,1000000,3000000,10000000")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    args = ap.parse_args()
    import torch

    import small_gicp_b200 as sg

    if not torch.cuda.is_available():
        raise SystemExit("sweep.py: no CUDA device; the hot path has no CPU fallback")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = torch.zeros(64, dtype=torch.float64, device=dev)
    peak, peak_src = B.hbm_peak()
    rows = []
    for n in [int(x) for x in args.sizes.split(",")]:
        t_gen = time.perf_counter()
        ctx = sg.Context(0)
        ctx.set_stream(stream.cuda_stream)
        inp = B.make_inputs(n, 0, "knn", lambda p4: ctx.estimate_features(p4, 20, normals=False)[1])
        t_gen = time.perf_counter() - t_gen
        ctx.set_target(inp["target"], None, inp["target_covs"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.build_target_kdtree(0)
        ctx.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        ctx.set_source(inp["source"], inp["source_covs"])

        def lin_host(T):
            ctx.linearize_device(T, out.data_ptr(), factor=sg.FACTOR_GICP, rejector=sg.REJECT_DISTANCE, max_dist_sq=1.0)
            h = out[:44].cpu().numpy()
            return h[:36].reshape(6, 6), h[36:42], h[42]

        poses, T_final = B.gn_trajectory(lin_host)
        err = np.linalg.inv(inp["T_gt"]) @ T_final
        rot = float(np.arccos(np.clip((np.trace(err[:3, :3]) - 1) / 2, -1, 1)))
        for fname, factor, bpp in (("GICP", sg.FACTOR_GICP, 100), ("ICP", sg.FACTOR_ICP, 36)):
            for pname, T in (("iteration-0 (identity)", poses[0]), ("converged", poses[-1])):
                for _ in range(3):
                    flush.zero_()
                    ctx.linearize_device(T, out.data_ptr(), factor=factor, rejector=sg.REJECT_DISTANCE, max_dist_sq=1.0)
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
                torch.cuda.synchronize()
                for a, b in ev:
                    flush.zero_()
                    a.record(stream)
                    ctx.linearize_device(T, out.data_ptr(), factor=factor, rejector=sg.REJECT_DISTANCE, max_dist_sq=1.0)
                    b.record(stream)
                torch.cuda.synchronize()
                ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
                inl = int(out[43].item())
                rows.append({
                    "points": n, "factor": fname, "pose": pname, "ms_per_linearize": ms, "mpoints_per_s": n / (ms * 1e-3) / 1e6,
                    "algorithmic_gbs": bpp * n / (ms * 1e-3) / 1e9, "roofline_frac": bpp * n / (ms * 1e-3) / 1e9 / peak, "inliers": inl,
                    "target_build_ms": build_ms, "gn_iterations": len(poses), "rot_err_vs_gt_rad": rot, "input_prep_s": t_gen,
                })
                print(json.dumps(rows[-1]), flush=True)
        ctx.close()
        del inp
        # written after every size: a run cut short by its time limit still leaves the sizes it finished
        doc = {"what": "BASELINE configs[4] size sweep, one B200, L2 flushed before every timed linearize, CUDA events", "hbm_peak_gbs": peak, "peak_source": peak_src,
               "gpu": torch.cuda.get_device_name(0), "rows": rows}
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(doc, f, indent=1)
    print("\n| points | factor | pose | ms / linearize | Mpoints/s | algorithmic GB/s | of HBM peak |\n|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['points']:,} | {r['factor']} | {r['pose']} | {r['ms_per_linearize']:.4f} | {r['mpoints_per_s']:.0f} | {r['algorithmic_gbs']:.0f} | {100 * r['roofline_frac']:.1f} % |")


if __name__ == "__main__":
    main()
