"""The C++ host mirror of the reference's template surface (small_gicp_b200/host/include): CPU-side API checks,
and on the GPU the full Registration<Factor, ParallelReductionCUDA, ...>::align against the oracle's
Registration<Factor, SerialReduction, ...>::align on identical inputs (1e-4 rad / 1e-3 m, BASELINE north_star)."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

import oracle as O
from conftest import noise_poses, pose_error


def test_host_kdtree_knn_matches_brute_force(golden_prepared):
    """kdtree_test.cpp:81-105 against the mirror's KdTree<PointCloud>::knn_search (exact indices, d2)."""
    from small_gicp_b200 import host_api

    P = golden_prepared["target"].points
    q = np.concatenate([P[:60], golden_prepared["source"].points[:60]])
    for k in (1, 5, 20):
        idx, d2 = host_api.kdtree_knn(P, q, k)
        dd, ii = cKDTree(P[:, :3]).query(q[:, :3], k=k)
        np.testing.assert_allclose(d2, dd.reshape(len(q), -1) ** 2, atol=1e-9)
        assert np.all(idx == ii.reshape(len(q), -1).astype(np.uint64))
    # fewer points than k: min(k, N) found, the rest stays at the sentinel (kdtree_synthetic_test.cpp:177-193)
    idx, d2 = host_api.kdtree_knn(P[:5], q[:3], 20)
    assert np.all(idx[:, 5:] == np.uint64(0xFFFFFFFFFFFFFFFF)) and np.all(idx[:, :5] < 5)


VARIANTS = [
    ("ICP", 0, 0),
    ("PLANE_ICP", 1, 0),
    ("GICP", 2, 0),
    ("HUBER_GICP", 2, 1),
    ("CAUCHY_GICP", 2, 2),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,factor,robust", VARIANTS)
@pytest.mark.parametrize("optimizer", [0, 1])
def test_align_matches_oracle(golden_prepared, name, factor, robust, optimizer):
    from small_gicp_b200 import host_api

    g = golden_prepared
    tc, sc = g["target"], g["source"]
    reg = O.Registration(factor=factor, robust=robust, num_threads=0)
    reg.set_optimizer(type=optimizer)
    for j, Tn in enumerate(noise_poses()[:3]):
        ref = reg.align(tc, g["target_tree"], sc, Tn)
        for tree in (host_api.TREE_HOST_KDTREE, host_api.TREE_DEVICE_KDTREE):
            r = host_api.align(tc.points, sc.points, tc.normals, tc.covs, sc.covs, init_T=Tn, factor=factor, robust=robust, optimizer=optimizer, tree=tree)
            rot, trans = pose_error(ref.T_target_source, r.T_target_source)
            assert rot < 1e-4 and trans < 1e-3, (name, optimizer, j, tree, rot, trans)
            assert r.converged == ref.converged and r.iterations == ref.iterations
            assert abs(r.num_inliers - ref.num_inliers) <= 2
            assert abs(r.error - ref.error) <= 1e-4 * max(ref.error, 1e-9)
            rot, trans = pose_error(g["T"], r.T_target_source)
            assert rot < np.deg2rad(2.5) and trans < 0.2  # registration_test.cpp:139-151


@pytest.mark.gpu
def test_align_vgicp_matches_oracle(golden_prepared):
    from small_gicp_b200 import host_api

    g = golden_prepared
    tc, sc = g["target"], g["source"]
    for offsets in (1, 7):
        vm = O.GaussianVoxelMap(tc, 1.0, offsets)
        reg = O.Registration(factor=O.FACTOR_GICP, num_threads=0)
        for Tn in noise_poses()[:2]:
            ref = reg.align(vm, None, sc, Tn)
            r = host_api.align(tc.points, sc.points, None, tc.covs, sc.covs, init_T=Tn, tree=host_api.TREE_VOXELMAP, voxel_resolution=1.0, voxel_search_offsets=offsets)
            rot, trans = pose_error(ref.T_target_source, r.T_target_source)
            assert rot < 1e-4 and trans < 1e-3, (offsets, rot, trans)
            assert r.iterations == ref.iterations and abs(r.num_inliers - ref.num_inliers) <= 2


@pytest.mark.gpu
def test_restrict_dof_and_null_rejector(golden_prepared):
    """general_factor.hpp:41-75: masked degrees of freedom stay frozen; NullRejector keeps every point."""
    from small_gicp_b200 import host_api

    g = golden_prepared
    tc, sc = g["target"], g["source"]
    r = host_api.align(tc.points, sc.points, tc.normals, tc.covs, sc.covs, dof_mask=[0, 0, 1, 1, 1, 0], rejector=0)
    T = r.T_target_source
    # soft constraint (lambda = 1e9 on the masked diagonal, general_factor.hpp:66): frozen up to ~1e-4
    assert abs(T[2, 3]) < 1e-3  # z translation frozen
    assert abs(T[2, 2] - 1.0) < 1e-6  # only yaw
    assert r.num_inliers == len(sc)


@pytest.mark.gpu
def test_align_empty_and_tiny(golden_prepared):
    from small_gicp_b200 import host_api

    g = golden_prepared
    tc, sc = g["target"], g["source"]
    r = host_api.align(tc.points, np.zeros((0, 4)), tc.normals, tc.covs, np.zeros((0, 4, 4)))
    assert r.num_inliers == 0 and np.allclose(r.T_target_source, np.eye(4))
    r = host_api.align(np.zeros((0, 4)), sc.points, None, np.zeros((0, 4, 4)), sc.covs, tree=host_api.TREE_DEVICE_KDTREE)
    assert r.num_inliers == 0
