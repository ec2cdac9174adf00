#!/usr/bin/env python
"""Turn one `scripts/gpu_ncu.sh` (or `gpu_round.sh`) visit (gpurun_out/<tag>/) into the tracked artefacts under profiles/:
   profiles/<round>/<letter>_linearize.md  -- per-kernel ncu summary of the four launches of sgb_linearize over the bench's poses
   profiles/linearize_traffic.json         -- mean DRAM bytes per linearize (feeds bench.py's roofline.traffic)
   copies of the bench lines and launch lists.
usage: python scripts/ncu_summary.py gpurun_out/r01ae profiles/r01 ae
"""
import csv
import json
import os
import shutil
import subprocess
import sys


def main():
    src, dst, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    rep = os.path.join(src, "prof_linearize.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]

    def g(r, w):
        return r[hdr.index(w)]

    recs = []
    for r in rows[2:]:
        recs.append(
            dict(
                name=g(r, "Kernel Name").split("(")[0].replace("void ", "").replace("sgb::", ""),
                dur=float(g(r, "gpu__time_duration.sum")),
                mb=float(g(r, "dram__bytes_read.sum")) + float(g(r, "dram__bytes_write.sum")),
                inst=float(g(r, "smsp__inst_executed.sum")),
                issue=float(g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")),
                occ=float(g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")),
                regs=g(r, "launch__registers_per_thread"),
                grid=g(r, "launch__grid_size"),
                thr=g(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                l2=float(g(r, "lts__t_sector_hit_rate.pct")),
                fp64=float(g(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")),
            )
        )
    assert rows[1][hdr.index("dram__bytes_read.sum")] == "Mbyte" and rows[1][hdr.index("gpu__time_duration.sum")] == "us"
    K = 3  # launches per linearize: probe, packet_search (few pending: warp-per-query ring search; many: packet walk), factor_reduce
    groups = [recs[k : k + K] for k in range(0, len(recs) - len(recs) % K, K)]
    conv = [x for x in groups if x[1]["dur"] < 40.0]  # few queries pending: the finishing kernel runs its warp-per-query regime
    mis = [x for x in groups if x[1]["dur"] >= 40.0]

    def avg(gs, i, key):
        return sum(x[i][key] for x in gs) / len(gs)

    b = json.load(open(os.path.join(src, "bench.json")))
    r = json.load(open(os.path.join(src, "bench_ref.json")))
    L = []
    L.append(f"# {os.path.basename(dst)}/{tag} -- the three launches of `sgb_linearize` (1M x 1M synthetic GICP), ncu --set full\n")
    L.append('Command: `ncu --set full --clock-control none --import-source on -k regex:"grid_probe|packet_search|factor_reduce" -s 45 -c 15 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras` with `SGB_BENCH_ROLL=5` (`scripts/gpu_ncu.sh`)')
    L.append(f"(k=20 covariances, L2 flushed between steps; the captured launches are one pass over the {len(groups)} poses of the Gauss-Newton trajectory: {len(conv)} converged, {len(mis)} misaligned).")
    L.append("Which regime of the finishing search kernel runs is decided on the device from the probe's pending counter.\n")
    L.append(f"## Converged poses -- mean of {len(conv)} linearizes\n")
    L.append("| metric | grid_probe_blocks | packet_search (warp-per-query regime) | factor_reduce<GICP,none> |")
    L.append("|---|---|---|---|")
    for label, key, fmt in (
        ("gpu__time_duration (under ncu, us)", "dur", "{:.1f}"),
        ("dram read + write (MB)", "mb", "{:.1f}"),
        ("sm__warps_active (%)", "occ", "{:.1f}"),
        ("smsp__issue_active (%)", "issue", "{:.1f}"),
        ("L2 hit (%)", "l2", "{:.0f}"),
        ("FP64 pipe active (%)", "fp64", "{:.1f}"),
    ):
        L.append(f"| {label} | " + " | ".join(fmt.format(avg(conv, i, key)) for i in range(K)) + " |")
    L.append("| warp instructions (M) | " + " | ".join(f"{avg(conv, i, 'inst') / 1e6:.2f}" for i in range(K)) + " |")
    L.append("| threads per instruction | " + " | ".join(conv[0][i]["thr"] for i in range(K)) + " |")
    L.append("| registers, grid | " + " | ".join(f"{conv[0][i]['regs']}, {conv[0][i]['grid']}" for i in range(K)) + " |")
    L.append(f"\nSum: **{sum(avg(conv, i, 'dur') for i in range(K)):.0f} us**, {sum(avg(conv, i, 'mb') for i in range(K)):.0f} MB of DRAM traffic per linearize.\n")
    if mis:
        L.append("## Misaligned first iterations -- more than N/16 queries pending: the packet walk takes them\n")
        L.append("| pose | probe us | packet us | factor us | sum us | DRAM MB |")
        L.append("|---|---|---|---|---|---|")
        for n, x in enumerate(mis):
            L.append(f"| T{n} | {x[0]['dur']:.1f} | {x[1]['dur']:.1f} | {x[2]['dur']:.1f} | {sum(y['dur'] for y in x):.0f} | {sum(y['mb'] for y in x):.0f} |")
    traffic = sum(sum(y["mb"] for y in x) for x in groups) / len(groups)
    kernel_us = sum(sum(y["dur"] for y in x) for x in groups) / len(groups)
    L.append(f"\nMean over the poses: {kernel_us:.0f} us of kernel time (cold, serialised by the profiler), **{traffic:.1f} MB** of DRAM traffic per linearize")
    L.append("(`profiles/linearize_traffic.json` feeds `roofline.traffic`); algorithmic bytes: 100 MB (GICP 100 B/point x 1M).\n")
    L.append(f"Timed without the profiler (`{tag}_bench_n1.json`: CUDA events on the launching stream, 20 steps, L2 flushed between steps):")
    L.append(f"**{b['ms_per_step'] * 1e3:.1f} us per linearize = {b['value']:.0f} Mpoints/s** = {b['roofline']['achieved']:.0f} GB/s algorithmic = **{b['roofline']['frac'] * 100:.1f} % of the measured {b['roofline']['peak']} GB/s HBM peak**")
    L.append(f"({traffic / b['ms_per_step'] / 1e3:.2f} TB/s of actual DRAM traffic); {b['value_l2_warm']:.0f} Mpoints/s with a warm L2; e2e (160 MB H2D of the reference-layout source per step) {b['e2e']['value']:.0f} Mpoints/s;")
    L.append(f"clocks {b['clocks']['sm_mhz']:.0f}/{b['clocks']['sm_max_mhz']:.0f} MHz, throttle reasons {b['clocks']['reasons']}; target kd-tree + block lists built in {b['setup']['target_build_ms']:.1f} ms (first call).")
    L.append(f"Reference arm (`{tag}_bench_reference_arm.json`, kind \"{r['cpu_baseline']['kind']}\", {r['cpu_baseline']['cores']} threads): {r['value']:.2f} Mpoints/s ({r['ms_per_step']:.0f} ms per linearize);")
    if b.get("cpu_baseline"):
        L.append(f"sums agree with it to {b['cpu_baseline']['parity_rel_H']:.1e} (H) / {b['cpu_baseline']['parity_rel_e']:.1e} (e).")
    os.makedirs(dst, exist_ok=True)
    open(os.path.join(dst, f"{tag}_linearize.md"), "w").write("\n".join(L) + "\n")
    json.dump(
        {
            "dram_bytes_per_launch": traffic * 1e6,
            "source": f"{dst}/{tag}_linearize.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum summed over the three launches of one sgb_linearize, mean over the poses of the bench trajectory, 1M x 1M GICP)",
        },
        open(os.path.join(os.path.dirname(dst.rstrip("/")), "linearize_traffic.json"), "w"),
        indent=1,
    )
    for a, c in (("bench.json", "bench_n1.json"), ("bench_ref.json", "bench_reference_arm.json"), ("launches_timed.csv", "launches_timed_region.csv"), ("launches_setup.csv", "launches_setup_first400.csv"), ("gpu.txt", "gpu.txt"), ("nproc.txt", "nproc.txt")):
        if os.path.exists(os.path.join(src, a)):
            shutil.copy(os.path.join(src, a), os.path.join(dst, f"{tag}_{c}"))
    print("\n".join(L))


if __name__ == "__main__":
    main()
