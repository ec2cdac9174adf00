#!/usr/bin/env python3
"""CPU-side analysis behind profiles/r01/ag_nn_distance_analysis.md: distribution of the exact nearest-neighbour distance of the
1M x 1M bench workload at every pose of its Gauss-Newton trajectory, and how far the queries move between consecutive poses.
Runs the CPU oracle (test infrastructure) -- analysis only, nothing the product uses.   python scripts/nn_distance_histogram.py"""
import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench as B, oracle as O
inp = B.make_inputs(1_000_000, 0, "analytic")
tc, sc = O.Cloud(inp["target"]), O.Cloud(inp["source"])
tt = O.KdTree(tc)
tc.set_features(None, inp["target_covs"]); sc.set_features(None, inp["source_covs"])
reg = O.Registration(factor=O.FACTOR_GICP, rejector=O.REJECT_NONE, num_threads=8)
regd = O.Registration(factor=O.FACTOR_GICP, rejector=O.REJECT_DISTANCE, max_dist_sq=1.0, num_threads=8)
poses, Tf = B.gn_trajectory(lambda T: regd.linearize(tc, tt, sc, T))
print("iterations", len(poses))
P, S = inp["target"][:, :3], inp["source"][:, :3]
for k, T in enumerate(poses):
    t0 = time.time()
    reg.linearize(tc, tt, sc, T)
    corr = reg.correspondences(len(sc)).astype(np.int64)
    q = S @ T[:3, :3].T + T[:3, 3]
    d = np.linalg.norm(P[corr] - q, axis=1)
    print(k, "frac d>0.2: %.4f  d>0.4: %.4f  d>0.6: %.4f d>1.0: %.4f  median %.3f" % ((d > 0.2).mean(), (d > 0.4).mean(), (d > 0.6).mean(), (d > 1.0).mean(), np.median(d)), "step", round(time.time() - t0, 1), "s")
    if k + 1 < len(poses):
        dT = np.linalg.inv(T) @ poses[k + 1]
        disp = np.linalg.norm((S @ poses[k+1][:3,:3].T + poses[k+1][:3,3]) - q, axis=1)
        print("   displacement to next pose: median %.4f max %.4f" % (np.median(disp), disp.max()))
