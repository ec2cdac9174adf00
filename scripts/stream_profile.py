#!/usr/bin/env python3
"""One LiDAR-like frame pair through the per-frame pipeline of bench.py's c3 leg (voxel grid -> source upload + covariances -> LM align ->
target tree + grid + covariances), a few times: run under `ncu --metrics gpu__time_duration.sum` to see which kernels a frame spends its
GPU time in (scripts/gpu_stream.sh), or alone for the wall-clock split."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as B
import small_gicp_b200 as sg
from small_gicp_b200 import synthetic as syn

dev = torch.device("cuda", 0)
world = syn.make_world(16_000_000, 42)
frames = [syn.lidar_frame_torch(world, syn.lidar_pose(f, 400.0), 45 + f, dev).cpu().numpy() for f in range(4)]
ctx = sg.Context(0)
prev = None
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    for f in range(4):
        t = [time.perf_counter()]
        pts = ctx.voxelgrid_sampling(frames[f], 0.25); ctx.synchronize(); t.append(time.perf_counter())
        if prev is not None:
            ctx.set_source(pts); ctx.synchronize(); t.append(time.perf_counter())
            ctx.estimate_source_features(20); ctx.synchronize(); t.append(time.perf_counter())
            B.lm_align(lambda T: ctx.linearize(T, factor=sg.FACTOR_GICP), ctx.error); t.append(time.perf_counter())
        ctx.set_target(pts); ctx.synchronize(); t.append(time.perf_counter())
        ctx.build_target_kdtree(0); ctx.synchronize(); t.append(time.perf_counter())
        ctx.estimate_target_features(20); ctx.synchronize(); t.append(time.perf_counter())
        if prev is not None and rep > 0:
            names = ["voxelgrid", "set_source", "source_features", "lm_align", "set_target", "build_tree+grid", "target_features"]
            print(len(pts), " ".join("%s %.3f" % (n, (b - a) * 1e3) for n, a, b in zip(names, t[:-1], t[1:])))
        prev = pts
