"""Device-side per-cloud preparation against the oracle's restatement of
util/downsampling.hpp:22-78 and util/normal_estimation.hpp:12-140."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import small_gicp_b200 as sg

    c = sg.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("leaf", [0.1, 0.25, 0.5, 1.0])
def test_voxelgrid_matches_oracle(golden, ctx, leaf):
    tgt, src, _ = golden
    for xyz in (tgt, src):
        ref = O.Cloud(xyz).voxelgrid_sampling(leaf).points
        got = ctx.voxelgrid_sampling(xyz, leaf)
        assert got.shape == ref.shape  # downsampling_test.cpp:90-98 compares sizes; we require equality
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9)  # same voxel order (ascending key), same means
        assert np.all(got[:, 3] == 1.0)


def test_voxelgrid_edge_cases(ctx):
    assert ctx.voxelgrid_sampling(np.zeros((0, 3)), 0.25).shape == (0, 4)
    one = ctx.voxelgrid_sampling(np.array([[1.0, 2.0, 3.0]]), 0.25)
    np.testing.assert_allclose(one, [[1.0, 2.0, 3.0, 1.0]])
    # negative coordinates use floor, not truncation (fast_floor.hpp:12-15)
    pts = np.array([[-0.1, -0.1, -0.1], [-0.2, -0.2, -0.2], [0.1, 0.1, 0.1]])
    got = ctx.voxelgrid_sampling(pts, 1.0)
    ref = O.Cloud(pts).voxelgrid_sampling(1.0).points
    np.testing.assert_allclose(got, ref, atol=1e-12)
    # out-of-range voxel coordinates are dropped
    far = np.array([[0.0, 0.0, 0.0], [1e9, 0.0, 0.0]])
    assert len(ctx.voxelgrid_sampling(far, 0.25)) == 1


@pytest.mark.parametrize("k", [10, 20])
def test_features_match_oracle(golden, ctx, k):
    tgt, _, _ = golden
    cloud = O.Cloud(tgt).voxelgrid_sampling(0.25)
    tree = O.KdTree(cloud)
    tree.estimate(k, O.FEAT_NORMAL_COV, max(1, O.max_threads()))
    P, N0, C0 = cloud.points, cloud.normals, cloud.covs
    N, C = ctx.estimate_features(P, k)
    # structural properties (normal_estimation_test.cpp:39-77)
    np.testing.assert_allclose(np.linalg.norm(N[:, :3], axis=1), 1.0, atol=1e-9)
    assert np.all(N[:, 3] == 0) and np.all(C[:, 3, :] == 0) and np.all(C[:, :, 3] == 0)
    np.testing.assert_allclose(C, np.swapaxes(C, 1, 2), atol=1e-15)
    assert np.all(np.einsum("ij,ij->i", P[:, :3], N[:, :3]) <= 1e-9)
    # values: identical neighbour sets give the same eigenvector up to rounding; FP32 near-ties at the k-th neighbour
    # may swap one neighbour for a few points
    # the orientation test (p . n > 0 -> flip) is a coin toss when p . n ~ 0: compare normals up to sign, and require
    # every sign disagreement to sit on that boundary
    flipped = np.einsum("ij,ij->i", N[:, :3], N0[:, :3]) < 0
    pn = np.abs(np.einsum("ij,ij->i", P[:, :3], N0[:, :3])) / np.linalg.norm(P[:, :3], axis=1)
    if flipped.any():
        j = np.nonzero(flipped)[0][:3]
        print("sign coin tosses:", int(flipped.sum()), "pn", pn[flipped][:8], "\nP", P[j, :3], "\nN", N[j, :3], "\nN0", N0[j, :3])
    assert flipped.sum() <= max(5, 0.005 * len(P)) and np.all((pn[flipped] < 1e-6) | ~np.isfinite(pn[flipped])), (flipped.sum(), pn[flipped])
    assert np.isfinite(N0).all(axis=1).mean() > 0.999  # the oracle's closed-form solver may emit NaN on degenerate neighbourhoods
    N0 = np.nan_to_num(N0)
    C0 = np.nan_to_num(C0)
    dn = np.minimum(np.abs(N - N0).max(axis=1), np.abs(N + N0).max(axis=1))
    dc = np.abs(C - C0).reshape(len(P), -1).max(axis=1)
    # The search runs in FP32, but the k nearest are picked -- and the covariance summed -- on the exact coordinates: the neighbour SETS are
    # the oracle's (no near-tie swaps), what is left is the conditioning of the smallest eigenvector of nearly isotropic neighbourhoods.
    ok = (dn < 1e-6) & (dc < 1e-6)
    print("features k=%d: within 1e-6: %.5f, median dn %.2e dc %.2e, max dn %.2e dc %.2e" % (k, ok.mean(), np.median(dn), np.median(dc), dn.max(), dc.max()))
    assert ok.mean() > 0.999, ok.mean()
    assert np.median(dn) < 1e-9 and np.median(dc) < 1e-9


def test_features_few_points(ctx):
    pts = np.random.default_rng(0).normal(size=(4, 3))
    N, C = ctx.estimate_features(pts, 10)
    assert not N.any()  # < 5 neighbours: zero normal, identity covariance
    ident = np.eye(4)
    ident[3, 3] = 0
    np.testing.assert_allclose(C, np.tile(ident, (4, 1, 1)))


def test_device_resident_pipeline(golden):
    """points only -> tree + features computed on the device -> GICP linearize; equals the path where the same
    features come back to the host and are uploaded again."""
    import small_gicp_b200 as sg

    tgt, src, T = golden
    a, b = sg.Context(0), sg.Context(0)
    tp = a.voxelgrid_sampling(tgt, 0.25)
    sp = a.voxelgrid_sampling(src, 0.25)
    # resident
    a.set_target(tp)
    a.build_target_kdtree()
    a.estimate_target_features(10)
    a.set_source(sp)
    a.estimate_source_features(10)
    # round trip
    tn, tc = b.estimate_features(tp, 10)
    _, sc = b.estimate_features(sp, 10, normals=False)
    b.set_target(tp, tn, tc)
    b.build_target_kdtree()
    b.set_source(sp, sc)
    for factor in (sg.FACTOR_GICP, sg.FACTOR_PLANE_ICP):
        Ha, ba, ea = a.linearize(T, factor=factor)
        Hb, bb, eb = b.linearize(T, factor=factor)
        assert np.linalg.norm(Ha - Hb) <= 1e-4 * np.linalg.norm(Hb) and abs(ea - eb) <= 1e-4 * eb
        assert np.array_equal(a.correspondences(), b.correspondences())
    a.close()
    b.close()


@pytest.mark.parametrize("k", [1, 5, 20, 32])
def test_batch_knn_matches_oracle(golden, k):
    """sgb_target_batch_knn against the oracle's restatement of UnsafeKdTree::knn_search (ann/kdtree.hpp:165-233): the same
    neighbours in the same order (FP32 near-ties verified by distance), squared distances to FP32 storage accuracy --
    kdtree_test.cpp:81-105 semantics with arbitrary (off-cloud) queries."""
    import small_gicp_b200 as sg

    tgt, src, T = golden
    cloud = O.Cloud(tgt).voxelgrid_sampling(0.25)
    tree = O.KdTree(cloud)
    rng = np.random.default_rng(3)
    queries = np.concatenate([src[:4000] @ T[:3, :3].T + T[:3, 3], cloud.points[:500, :3], rng.uniform(-60, 60, (300, 3))])
    ridx, rd2, rcnt = tree.knn(queries, k, max(1, O.max_threads()))
    for own in (True, False):
        c = sg.Context(0)
        c.set_target(cloud.points)
        if own:
            c.build_target_kdtree(0)
        else:
            c.set_target_kdtree(*tree.export())
        idx, d2 = c.target_batch_knn(queries, k)
        c.close()
        assert idx.shape == (len(queries), k) and np.all(np.diff(d2, axis=1) >= 0)
        # target coordinates are stored in FP32 relative to the cloud's centre: |delta p| <= ~8e-6 m at this extent, so
        # |delta d^2| <= 2 d |delta p|
        tol = 2.0 * np.sqrt(rd2) * 8e-6 + 1e-9
        assert np.all(np.abs(d2 - rd2) <= tol), float(np.max(np.abs(d2 - rd2) / tol))
        differ = idx != ridx
        assert differ.mean() < 2e-3
        # where the order differs the distances are tied to that accuracy
        assert np.all(np.abs(d2[differ] - rd2[differ]) <= tol[differ])
    # fewer target points than k: padded like KnnResult's initial state
    c = sg.Context(0)
    c.set_target(cloud.points[:3])
    c.build_target_kdtree(0)
    idx, d2 = c.target_batch_knn(queries[:10], 5)
    c.close()
    assert np.all(idx[:, 3:] == np.uint64(0xFFFFFFFFFFFFFFFF)) and np.all(idx[:, :3] < 3)


def test_source_features_survive_a_voxelmap_build(golden):
    """The VGICP order of calls: set_source -> build_target_voxelmap -> estimate_source_features.  The voxel-map build used to
    scribble its integer voxel coordinates over the buffer that holds the source in original order; the source covariances
    then came out of garbage.  They must equal those estimated before the map was built."""
    import small_gicp_b200 as sg

    tgt, src, T = golden
    a, b = sg.Context(0), sg.Context(0)
    tp = a.voxelgrid_sampling(tgt, 0.25)
    sp = a.voxelgrid_sampling(src, 0.25)
    _, tcov = a.estimate_features(tp, 10, normals=False)
    a.set_source(sp)
    a.build_target_voxelmap(tp, tcov, 1.0)
    a.estimate_source_features(10)
    b.set_source(sp)
    b.estimate_source_features(10)
    b.build_target_voxelmap(tp, tcov, 1.0)
    Ha, ba, ea = a.linearize(T, factor=sg.FACTOR_GICP)
    Hb, bb, eb = b.linearize(T, factor=sg.FACTOR_GICP)
    assert np.array_equal(Ha, Hb) and ea == eb and ea > 0
    # and against covariances that made the round trip through the host
    _, scov = b.estimate_features(sp, 10, normals=False)
    b.set_source(sp, scov)
    Hc, bc, ec = b.linearize(T, factor=sg.FACTOR_GICP)
    assert np.linalg.norm(Ha - Hc) <= 1e-4 * np.linalg.norm(Hc)
    c = sg.Context(0)
    c.estimate_source_features(10)  # no source at all: nothing to do (and nothing computed from whatever a scratch buffer holds)
    assert c.source_size == 0
    c.close()
    a.close()
    b.close()


def test_device_features_follow_a_tree_rebuild(golden):
    """Features estimated on the device live in leaf order; re-building (or adopting another) tree re-permutes the points.  The
    features must follow their points: same sums before and after the rebuild, for either tree."""
    import small_gicp_b200 as sg

    tgt, src, T = golden
    a = sg.Context(0)
    tp = a.voxelgrid_sampling(tgt, 0.25)
    sp = a.voxelgrid_sampling(src, 0.25)
    a.set_target(tp)
    a.build_target_kdtree()
    a.estimate_target_features(10)
    a.set_source(sp)
    a.estimate_source_features(10)
    ref = {f: a.linearize(T, factor=f) for f in (sg.FACTOR_GICP, sg.FACTOR_PLANE_ICP)}
    tree = O.KdTree(O.Cloud(tp))
    a.set_target_kdtree(*tree.export())  # a different permutation of the same points
    for f, (H0, b0, e0) in ref.items():
        H, b, e = a.linearize(T, factor=f)
        assert np.linalg.norm(H - H0) <= 1e-6 * np.linalg.norm(H0) and abs(e - e0) <= 1e-6 * e0, f
    a.build_target_kdtree()
    for f, (H0, b0, e0) in ref.items():
        H, b, e = a.linearize(T, factor=f)
        assert np.linalg.norm(H - H0) <= 1e-6 * np.linalg.norm(H0) and abs(e - e0) <= 1e-6 * e0, f
    a.close()


def test_adopting_the_source_equals_setting_the_target(golden):
    """sgb_target_adopt_source (frame streams): the source -- points, the tree its covariance estimation built, the covariances --
    becomes the target without a rebuild.  The next frame's sums must equal those against an explicitly set + built + estimated target,
    both for device-estimated and for host-supplied source covariances."""
    import small_gicp_b200 as sg

    tgt, src, T = golden
    a, b = sg.Context(0), sg.Context(0)
    f0 = a.voxelgrid_sampling(tgt, 0.25)
    f1 = a.voxelgrid_sampling(src, 0.25)
    # explicit
    b.set_target(f0)
    b.build_target_kdtree()
    b.estimate_target_features(20)
    b.set_source(f1)
    b.estimate_source_features(20)
    ref = b.linearize(T, factor=sg.FACTOR_GICP)
    ref_c = b.correspondences()
    # adopted: device-estimated covariances + the tree that estimation built
    a.set_source(f0)
    a.estimate_source_features(20)
    a.adopt_source_as_target()
    assert a.source_size == 0 and a.target_size == len(f0)
    a.set_source(f1)
    a.estimate_source_features(20)
    H, b_, e = a.linearize(T, factor=sg.FACTOR_GICP)
    assert np.linalg.norm(H - ref[0]) <= 1e-6 * np.linalg.norm(ref[0]) and abs(e - ref[2]) <= 1e-6 * ref[2]
    assert (a.correspondences() != ref_c).sum() <= 2
    with pytest.raises(sg.SgbError):  # no normals were handed over: point-to-plane says so instead of using stale ones
        a.linearize(T, factor=sg.FACTOR_PLANE_ICP)
    a.estimate_target_features(20)
    a.linearize(T, factor=sg.FACTOR_PLANE_ICP)
    # adopted: covariances that came from the host, no tree yet
    _, c0 = b.estimate_features(f0, 20, normals=False)
    a.set_source(f0, c0)
    a.adopt_source_as_target()
    a.set_source(f1)
    a.estimate_source_features(20)
    H2, _, e2 = a.linearize(T, factor=sg.FACTOR_GICP)
    assert np.linalg.norm(H2 - ref[0]) <= 1e-5 * np.linalg.norm(ref[0]) and abs(e2 - ref[2]) <= 1e-5 * ref[2]
    with pytest.raises(sg.SgbError):
        a.adopt_source_as_target()
        a.adopt_source_as_target()  # nothing left to adopt
    a.close()
    b.close()
