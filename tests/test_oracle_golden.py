"""Pins the CPU oracle (oracle/sgicp_oracle.cpp) against the reference's own golden data and test bars.

Reference tests mirrored (all paths relative to /root/reference):
  src/test/registration_test.cpp:99-151,284-292   pose within 2.5 deg / 0.2 m for 6 factor variants x reductions
  src/test/registration_test.cpp:217-224          final H symmetric (1e-3) with min eigenvalue > 10
  src/test/kdtree_test.cpp:81-105                 kNN == brute force (exact indices, d2 within 1e-3)
  src/test/kdtree_synthetic_test.cpp:177-193      k in {1,2,3,5,10,20}, min(k, |target|) results, distances only
  src/test/python_test.py:143-166                 sum of per-point H == reduced H
  src/test/helper_test.cpp:61-76                  preprocess_points sizes
  src/test/vector_test.cpp                        fast_floor == floor (through voxel coordinates)
"""
import numpy as np
import pytest
from scipy.spatial import cKDTree

import oracle as O
from conftest import noise_poses, pose_error

ROT_TOL = np.deg2rad(2.5)
TRANS_TOL = 0.2


def test_downsampling_sizes(golden):
    tgt, src, _ = golden
    assert (len(tgt), len(src)) == (69088, 69792)
    # probed sizes recorded in SURVEY.md §4 / §8 for the bundled clouds
    assert len(O.Cloud(tgt).voxelgrid_sampling(0.3)) == 5004
    assert len(O.Cloud(src).voxelgrid_sampling(0.3)) == 4950
    assert len(O.Cloud(tgt).voxelgrid_sampling(0.25)) == 6147
    assert len(O.Cloud(src).voxelgrid_sampling(0.25)) == 6167


def test_downsampling_is_voxel_mean(golden):
    tgt, _, _ = golden
    leaf = 0.5
    ds = O.Cloud(tgt).voxelgrid_sampling(leaf).points
    keys = np.floor(tgt / leaf).astype(np.int64)
    _, inv, cnt = np.unique(keys, axis=0, return_inverse=True, return_counts=True)
    means = np.zeros((len(cnt), 3))
    np.add.at(means, inv.ravel(), tgt)
    means /= cnt[:, None]
    assert len(ds) == len(means)
    a = ds[np.lexsort(ds[:, :3].T)][:, :3]
    b = means[np.lexsort(means.T)]
    np.testing.assert_allclose(a, b, atol=1e-9)
    assert np.all(ds[:, 3] == 1.0)


def test_empty_inputs():
    c = O.Cloud(np.zeros((0, 3)))
    assert len(c.voxelgrid_sampling(0.25)) == 0
    t = O.KdTree(c)
    idx, d2, cnt = t.knn(np.zeros((3, 4)), 5)
    assert np.all(cnt == 0) and np.all(idx == O.NO_INDEX)


@pytest.mark.parametrize("k", [1, 2, 3, 5, 10, 20])
def test_knn_matches_brute_force(golden_prepared, k):
    tc, sc, tt = golden_prepared["target"], golden_prepared["source"], golden_prepared["target_tree"]
    P = tc.points
    rng = np.random.default_rng(3)
    q = np.concatenate([P[:50], P[50:100] + rng.normal(0, 0.05, (50, 4)) * [1, 1, 1, 0], sc.points[:50]])
    idx, d2, cnt = tt.knn(q, k)
    dd, ii = cKDTree(P[:, :3]).query(q[:, :3], k=k)
    dd, ii = dd.reshape(len(q), -1), ii.reshape(len(q), -1)
    assert np.all(cnt == k)
    np.testing.assert_allclose(d2, dd**2, atol=1e-9)
    assert np.all(idx == ii.astype(np.uint64))


def test_knn_small_and_tied_sets():
    # kdtree_synthetic_test.cpp: 5-point / 10-point / integer-grid sets, distances compared only
    rng = np.random.default_rng(11)
    sets = [rng.uniform(-1, 1, (5, 3)), rng.uniform(-1e6, 1e6, (10, 3)), rng.integers(-3, 4, (256, 3)).astype(float)]
    for pts in sets:
        tree = O.KdTree(O.Cloud(pts))
        q = rng.uniform(-1, 1, (20, 3)) * np.abs(pts).max()
        for k in (1, 3, 20):
            idx, d2, cnt = tree.knn(q, k)
            kk = min(k, len(pts))
            assert np.all(cnt == kk)
            brute = np.sort(((q[:, None, :] - pts[None, :, :]) ** 2).sum(-1), axis=1)[:, :kk]
            np.testing.assert_allclose(d2[:, :kk], brute, rtol=1e-12, atol=1e-9)


def test_covariances_and_normals(golden_prepared):
    tc = golden_prepared["target"]
    P, N, Cv = tc.points, tc.normals, tc.covs
    # normal_estimation_test.cpp:39-77: unit normals with w = 0, symmetric covariances with zero padding
    np.testing.assert_allclose(np.linalg.norm(N[:, :3], axis=1), 1.0, atol=1e-9)
    assert np.all(N[:, 3] == 0.0)
    assert np.all(Cv[:, 3, :] == 0.0) and np.all(Cv[:, :, 3] == 0.0)
    np.testing.assert_allclose(Cv, np.swapaxes(Cv, 1, 2), atol=1e-12)
    # regularised eigenvalues (1e-3, 1, 1) and normal = smallest-eigenvalue direction (normal_estimation.hpp:40-45)
    w, V = np.linalg.eigh(Cv[:200, :3, :3])
    np.testing.assert_allclose(w, np.tile([1e-3, 1.0, 1.0], (200, 1)), atol=1e-7)
    dots = np.abs(np.einsum("ij,ij->i", V[:, :, 0], N[:200, :3]))
    np.testing.assert_allclose(dots, 1.0, atol=1e-6)
    # flipped toward the origin (normal_estimation.hpp:19-23)
    assert np.all(np.einsum("ij,ij->i", P[:, :3], N[:, :3]) <= 1e-12)
    # independent recomputation with numpy for a few points
    tree = cKDTree(P[:, :3])
    for i in (0, 17, 1234):
        _, nb = tree.query(P[i, :3], k=20)
        C = np.cov(P[nb, :3].T, bias=True)
        w_, V_ = np.linalg.eigh(C)
        assert abs(abs(V_[:, 0] @ N[i, :3]) - 1.0) < 1e-6


def test_algebra_helpers():
    rng = np.random.default_rng(5)
    from scipy.linalg import expm

    for _ in range(5):
        a = rng.normal(0, 0.3, 6)
        tw = np.zeros((4, 4))
        tw[:3, :3] = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        tw[:3, 3] = a[3:]
        np.testing.assert_allclose(O.se3_exp(a), expm(tw), atol=1e-12)
        A = rng.normal(size=(6, 6))
        A = A @ A.T + 1e-3 * np.eye(6)
        b = rng.normal(size=6)
        np.testing.assert_allclose(O.ldlt_solve6(A, b), np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)
        S = rng.normal(size=(3, 3))
        S = S @ S.T
        w, V = O.eigen_sym3(S)
        np.testing.assert_allclose(w, np.linalg.eigvalsh(S), rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(V @ np.diag(w) @ V.T, S, atol=1e-7)
    np.testing.assert_allclose(O.se3_exp(np.zeros(6)), np.eye(4), atol=0)


VARIANTS = {
    "ICP": (O.FACTOR_ICP, O.ROBUST_NONE),
    "PLANE_ICP": (O.FACTOR_PLANE, O.ROBUST_NONE),
    "GICP": (O.FACTOR_GICP, O.ROBUST_NONE),
    "HUBER_GICP": (O.FACTOR_GICP, O.ROBUST_HUBER),
    "CAUCHY_GICP": (O.FACTOR_GICP, O.ROBUST_CAUCHY),
}


@pytest.mark.parametrize("name", list(VARIANTS))
@pytest.mark.parametrize("threads", [0, 2])  # SerialReduction / ParallelReductionOMP
def test_registration_golden(golden_prepared, name, threads):
    g = golden_prepared
    factor, robust = VARIANTS[name]
    reg = O.Registration(factor=factor, robust=robust, num_threads=threads)
    Tinv = np.linalg.inv(g["T"])
    for Tn in noise_poses():
        r = reg.align(g["target"], g["target_tree"], g["source"], Tn)
        rot, trans = pose_error(g["T"], r.T_target_source)
        assert rot < ROT_TOL and trans < TRANS_TOL, (name, "forward", rot, trans)
        r = reg.align(g["source"], g["source_tree"], g["target"], Tn)
        rot, trans = pose_error(Tinv, r.T_target_source)
        assert rot < ROT_TOL and trans < TRANS_TOL, (name, "inverse", rot, trans)


def test_registration_vgicp_golden(golden_prepared):
    g = golden_prepared
    tv = O.GaussianVoxelMap(g["target"], 1.0)
    sv = O.GaussianVoxelMap(g["source"], 1.0)
    reg = O.Registration(factor=O.FACTOR_GICP, num_threads=2)
    for Tn in noise_poses():
        r = reg.align(tv, None, g["source"], Tn)
        rot, trans = pose_error(g["T"], r.T_target_source)
        assert rot < ROT_TOL and trans < TRANS_TOL
        r = reg.align(sv, None, g["target"], Tn)
        rot, trans = pose_error(np.linalg.inv(g["T"]), r.T_target_source)
        assert rot < ROT_TOL and trans < TRANS_TOL


def test_gn_optimizer_and_hessian(golden_prepared):
    g = golden_prepared
    reg = O.Registration(factor=O.FACTOR_GICP, num_threads=0)
    reg.set_optimizer(type=O.OPT_GN)
    r = reg.align(g["target"], g["target_tree"], g["source"], np.eye(4), trace=True)
    rot, trans = pose_error(g["T"], r.T_target_source)
    assert r.converged and rot < ROT_TOL and trans < TRANS_TOL
    # registration_test.cpp:217-224
    np.testing.assert_allclose(r.H, r.H.T, atol=1e-3)
    assert np.linalg.eigvalsh(r.H).min() > 10.0
    assert r.trace.shape == (r.iterations + 1, 59)
    # serial == OMP up to summation order (BENCHMARK.md:122-124)
    reg2 = O.Registration(factor=O.FACTOR_GICP, num_threads=2)
    reg2.set_optimizer(type=O.OPT_GN)
    r2 = reg2.align(g["target"], g["target_tree"], g["source"], np.eye(4))
    np.testing.assert_allclose(r2.T_target_source, r.T_target_source, atol=1e-9)


def test_sum_of_factors_equals_reduction(golden_prepared):
    """python_test.py:143-166 analogue, done exactly: per-point numpy evaluation of the GICP factor with the
    oracle's correspondences reproduces the reduced H, b, e."""
    g = golden_prepared
    T = g["T"]
    reg = O.Registration(factor=O.FACTOR_GICP, num_threads=0)
    H, b, e = reg.linearize(g["target"], g["target_tree"], g["source"], T)
    corr = reg.correspondences(len(g["source"]))
    P, Cs = g["source"].points, g["source"].covs
    Q, Ct = g["target"].points, g["target"].covs
    H2, b2, e2 = np.zeros((6, 6)), np.zeros(6), 0.0
    R = T[:3, :3]
    for i in np.nonzero(corr != O.NO_INDEX)[0]:
        k = int(corr[i])
        r = (Q[k] - T @ P[i])[:3]
        M = np.linalg.inv(Ct[k, :3, :3] + R @ Cs[i, :3, :3] @ R.T)
        p = P[i, :3]
        S = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
        J = np.hstack([R @ S, -R])
        H2 += J.T @ M @ J
        b2 += J.T @ M @ r
        e2 += 0.5 * r @ M @ r
    np.testing.assert_allclose(H, H2, rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(b, b2, rtol=1e-9, atol=1e-6)
    assert abs(e - e2) < 1e-9 * max(1.0, e2)
    # LM error() at the linearisation pose equals e (frozen mahalanobis, gicp_factor.hpp:81-89)
    assert abs(reg.error(g["target"], g["source"], T) - e) < 1e-9 * max(1.0, e)


def test_voxelmap_lookup(golden_prepared):
    g = golden_prepared
    vm = O.GaussianVoxelMap(g["target"], 1.0)
    coords, means, covs, cnt = vm.export()
    P = g["target"].points
    keys = np.floor(P[:, :3] / 1.0).astype(np.int32)
    uniq = np.unique(keys, axis=0)
    assert len(vm) == len(uniq) and cnt.sum() == len(P)
    # voxel order == first appearance order (incremental_voxelmap.hpp:61-67)
    _, first = np.unique(keys, axis=0, return_index=True)
    np.testing.assert_array_equal(coords, keys[np.sort(first)])
    idx, d2, found = vm.nn(P[:100])
    assert np.all(found == 1)
    vid = (idx >> np.uint64(32)).astype(np.int64)
    np.testing.assert_array_equal(coords[vid], keys[:100])
    np.testing.assert_allclose(d2, ((means[vid] - P[:100]) ** 2).sum(1), atol=1e-12)
