#!/usr/bin/env python3
"""One device-resident cloud pair of N points (bench.py's c5 recipe) and a few GICP linearize calls at the identity and at the converged pose:
the thing to put under ncu when the question is what bounds the kernels at 10M - 100M points (scripts/gpu_ncu_size.sh).
usage: python scripts/linearize_at_size.py N [calls per pose]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import small_gicp_b200 as sg
from small_gicp_b200 import synthetic as syn

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
world = syn.make_world(n, 42)
Tgt = syn.gt_transform_scaled(syn.world_side(n))
ctx = sg.Context(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)  # the events below must sit on the stream the library launches on
tgt = syn.sample_cloud_torch(world, n, 43, dev)
ctx.set_target(tgt); del tgt
ctx.build_target_kdtree(0)
ctx.estimate_target_features(20)
src = syn.sample_cloud_torch(world, n, 44, dev, transform=np.linalg.inv(Tgt))
ctx.set_source(src); del src
ctx.estimate_source_features(20)
out = torch.zeros(64, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for name, T in (("identity", np.eye(4)), ("converged", Tgt)):
    ms = []
    for _ in range(calls):
        ctx.drop_seeds()
        flush.fill_(1)
        torch.cuda.synchronize()
        ev[0].record(torch.cuda.current_stream())
        ctx.linearize_device(T, out.data_ptr(), factor=sg.FACTOR_GICP, rejector=sg.REJECT_DISTANCE, max_dist_sq=1.0)
        ev[1].record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        ms.append(ev[0].elapsed_time(ev[1]))
    print(name, "ms per linearize:", " ".join("%.4f" % x for x in ms), "inliers", int(out[43].item()))
ctx.close()
