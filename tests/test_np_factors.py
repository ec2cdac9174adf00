"""Pins the numpy leg of the parity triangle (tests/np_factors.py) to the oracle for every factor / robust kernel,
linearize and error -- CPU only, so the GPU tests can rely on it."""
import numpy as np
import pytest

import np_factors as NF
import oracle as O
from conftest import noise_poses

CASES = [(f, r) for f in (0, 1, 2) for r in (0, 1, 2)]


@pytest.mark.parametrize("factor,robust", CASES)
def test_numpy_leg_matches_oracle(golden_prepared, factor, robust):
    g = golden_prepared
    tc, sc = g["target"], g["source"]
    tp, tn, tcv, sp, scv = tc.points, tc.normals, tc.covs, sc.points, sc.covs
    reg = O.Registration(factor=factor, robust=robust, robust_c=0.7, num_threads=0)
    for T in (np.eye(4), g["T"], noise_poses()[1]):
        H, b, e = reg.linearize(tc, g["target_tree"], sc, T)
        corr = reg.correspondences(len(sc))
        H2, b2, e2 = NF.linearize(T, corr, factor, robust, 0.7, sp, scv, tp, tn, tcv)
        assert np.linalg.norm(H - H2) <= 1e-10 * np.linalg.norm(H)
        assert np.abs(b - b2).max() <= 1e-9 * np.sqrt(2 * e * np.diag(H)).max()
        assert abs(e - e2) <= 1e-10 * e
        T2 = T @ O.se3_exp(np.array([0.01, -0.02, 0.005, 0.05, -0.03, 0.02]))
        e_trial = reg.error(tc, sc, T2)
        assert abs(e_trial - NF.error(T2, T, corr, factor, robust, 0.7, sp, scv, tp, tn, tcv)) <= 1e-10 * e_trial
        assert NF.compare_correspondences(corr, corr, tp, sp, T) == 0
