"""Host-side logic of bench.py that the GPU legs rely on, checked on the CPU: the LevenbergMarquardt loop of the c3 / c4 legs against
the oracle's restatement of optimizer.hpp:83-149, the size-independent ground truth of the c4 / c5 pairs, the CPU-arm thread policy."""
import numpy as np

import bench as B
import oracle as O
from conftest import pose_error


def test_lm_loop_matches_oracle(golden_prepared):
    g = golden_prepared
    reg = O.Registration(factor=O.FACTOR_GICP, num_threads=0)
    ref = reg.align(g["target"], g["target_tree"], g["source"], np.eye(4))
    T, its, conv, n_lin, n_err = B.lm_align(lambda T: reg.linearize(g["target"], g["target_tree"], g["source"], T), lambda T: reg.error(g["target"], g["source"], T))
    rot, trans = pose_error(ref.T_target_source, T)
    assert rot < 1e-9 and trans < 1e-9
    assert its == ref.iterations and conv == ref.converged and n_lin == its + 1 and n_err >= n_lin


def test_gn_loop_matches_oracle(golden_prepared):
    g = golden_prepared
    reg = O.Registration(factor=O.FACTOR_GICP, num_threads=0)
    reg.set_optimizer(type=O.OPT_GN)
    ref = reg.align(g["target"], g["target_tree"], g["source"], np.eye(4))
    poses, T = B.gn_trajectory(lambda T: reg.linearize(g["target"], g["target_tree"], g["source"], T))
    rot, trans = pose_error(ref.T_target_source, T)
    assert rot < 1e-9 and trans < 1e-9 and len(poses) == ref.iterations + 1


def test_scaled_ground_truth_keeps_the_local_misalignment():
    from small_gicp_b200 import synthetic as syn

    base = None
    for n in (100_000, 1_000_000, 10_000_000, 100_000_000):
        side = syn.world_side(n)
        T = syn.gt_transform_scaled(side)
        assert abs(np.linalg.det(T[:3, :3]) - 1.0) < 1e-12
        corners = np.array([[0, 0, 0, 1], [side, 0, 0, 1], [0, side, 10, 1], [side, side, 10, 1], [side / 2, side / 2, 5, 1]], dtype=float)
        disp = np.linalg.norm((corners @ T.T - corners)[:, :3], axis=1)
        base = disp.max() if n == 1_000_000 else base
        assert disp.max() < 2.6, (n, disp)  # the 1M pair's farthest corner moves ~2.3 m; no size moves farther
        assert 0.3 < disp[-1] < 0.5  # the centre moves by the translation only
    assert base is not None


def test_thread_candidates_are_sane():
    c = B.thread_candidates()
    assert c == sorted(c, reverse=True) and c[-1] >= 1 and len(set(c)) == len(c)
    info = B.host_cpu_info()
    assert info["nproc"] >= 1 and "cgroup_cpu_max" in info
