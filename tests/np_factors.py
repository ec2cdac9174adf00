"""Independent numpy evaluation of the reduced system for a GIVEN correspondence set (third leg of the parity
triangle: oracle <-> numpy <-> GPU).  Follows the reference formulas literally:
  icp_factor.hpp:43-51, plane_icp_factor.hpp:43-55, gicp_factor.hpp:59-70, robust_kernel.hpp:24-27,47,84-89.
Used to make the GPU parity checks immune to legitimate FP32 near-ties / rejection-boundary flips: the GPU sums are
compared against numpy sums over the GPU's own correspondences, and the correspondences themselves are compared
against the oracle's with an explicit near-tie rule."""
import numpy as np

NO = np.uint64(0xFFFFFFFFFFFFFFFF)


def _skew_batch(p):
    S = np.zeros((len(p), 3, 3))
    S[:, 0, 1], S[:, 0, 2] = -p[:, 2], p[:, 1]
    S[:, 1, 0], S[:, 1, 2] = p[:, 2], -p[:, 0]
    S[:, 2, 0], S[:, 2, 1] = -p[:, 1], p[:, 0]
    return S


def robust_weight(kind, c, x):
    if kind == 1:
        ax = np.abs(x)
        return np.where(ax < c, 1.0, c / np.maximum(ax, 1e-300))
    if kind == 2:
        return c / (c + x * x)
    return np.ones_like(x)


def weights(T, factor, src_pts, src_covs, tgt_pts, tgt_normals, tgt_covs, idx, k):
    """per-point 3x3 weight matrix M (identity / diag(n^2) / fused GICP precision at pose T)"""
    n = len(idx)
    if factor == 0:
        return np.tile(np.eye(3), (n, 1, 1))
    if factor == 1:
        nn = tgt_normals[k, :3]
        M = np.zeros((n, 3, 3))
        for a in range(3):
            M[:, a, a] = nn[:, a] ** 2
        return M
    R = T[:3, :3]
    RCR = tgt_covs[k, :3, :3] + R @ src_covs[idx, :3, :3] @ R.T
    return np.linalg.inv(RCR)


def linearize(T, corr, factor, robust, robust_c, src_pts, src_covs, tgt_pts, tgt_normals, tgt_covs):
    """sum over accepted correspondences of J^T M J, J^T M r, 1/2 r^T M r (robust-weighted)"""
    idx = np.nonzero(corr != NO)[0]
    if len(idx) == 0:
        return np.zeros((6, 6)), np.zeros(6), 0.0
    k = corr[idx].astype(np.int64)
    R = T[:3, :3]
    p = src_pts[idx, :3]
    q = p @ R.T + T[:3, 3]
    r = tgt_pts[k, :3] - q
    M = weights(T, factor, src_pts, src_covs, tgt_pts, tgt_normals, tgt_covs, idx, k)
    J = np.concatenate([R @ _skew_batch(p), np.tile(-R, (len(idx), 1, 1))], axis=2)  # (n,3,6)
    MJ = M @ J
    Mr = np.einsum("nij,nj->ni", M, r)
    e = 0.5 * np.einsum("ni,ni->n", r, Mr)
    w = robust_weight(robust, robust_c, np.sqrt(e))
    H = np.einsum("n,nki,nkj->ij", w, J, MJ)
    b = np.einsum("n,nki,nk->i", w, J, Mr)
    return H, b, float((w * e).sum())


def error(T_trial, T_lin, corr, factor, robust, robust_c, src_pts, src_covs, tgt_pts, tgt_normals, tgt_covs):
    """Reduction::error: residuals at T_trial, GICP precision frozen at T_lin (gicp_factor.hpp:81-89)"""
    idx = np.nonzero(corr != NO)[0]
    if len(idx) == 0:
        return 0.0
    k = corr[idx].astype(np.int64)
    q = src_pts[idx, :3] @ T_trial[:3, :3].T + T_trial[:3, 3]
    r = tgt_pts[k, :3] - q
    M = weights(T_lin, factor, src_pts, src_covs, tgt_pts, tgt_normals, tgt_covs, idx, k)
    e = 0.5 * np.einsum("ni,nij,nj->n", r, M, r)
    return float((robust_weight(robust, robust_c, np.sqrt(e)) * e).sum())


def compare_correspondences(gpu, cpu, tgt_pts, src_pts, T, max_dist_sq=1.0, max_mismatch_frac=2e-3):
    """GPU (FP32 search on centred FP32 coordinates) vs oracle (FP64): identical except
       * near-ties: both candidates' squared distances agree to 1e-5 relative + 1e-7
       * rejection boundary: one side rejected, the other's squared distance within 1e-4 of max_dist_sq.
    Returns the number of differing entries."""
    mism = np.nonzero(gpu != cpu)[0]
    assert len(mism) <= max(3, max_mismatch_frac * len(cpu)), (len(mism), len(cpu))
    q = src_pts[:, :3] @ T[:3, :3].T + T[:3, 3]
    for i in mism:
        dg = np.inf if gpu[i] == NO else ((tgt_pts[int(gpu[i]), :3] - q[i]) ** 2).sum()
        dc = np.inf if cpu[i] == NO else ((tgt_pts[int(cpu[i]), :3] - q[i]) ** 2).sum()
        if np.isinf(dg) or np.isinf(dc):
            fin = dc if np.isinf(dg) else dg
            assert max_dist_sq is not None and abs(fin - max_dist_sq) < 1e-4 * max(1.0, max_dist_sq), (i, dg, dc)
        else:
            assert abs(dg - dc) <= 1e-5 * dc + 1e-7, (i, dg, dc)
    return len(mism)
