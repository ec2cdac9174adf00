"""N > 1 host logic on CPU (gloo, world_size 2): sharding + all-reduce of H|b|e gives every rank the single-process
result (<= 1e-12 relative), and the Gauss-Newton loop driven by it converges to the same pose on all
ranks.  The per-rank reduction here is the CPU oracle (the GPU path is exercised by bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import oracle as O
    from conftest import load_golden_xyz
    from small_gicp_b200.distributed import ShardedReduction, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tc = O.Cloud(load_golden_xyz("target")).voxelgrid_sampling(0.5)
    sc = O.Cloud(load_golden_xyz("source")).voxelgrid_sampling(0.5)
    tt, st = O.KdTree(tc), O.KdTree(sc)
    tt.estimate(10, O.FEAT_NORMAL_COV, 1)
    st.estimate(10, O.FEAT_COV, 1)
    lo, hi = shard_range(len(sc), rank, world)
    shard = O.Cloud(sc.points[lo:hi])
    shard.set_features(None, sc.covs[lo:hi])

    class OracleLocal:
        def __init__(self):
            self.reg = O.Registration(factor=O.FACTOR_GICP, num_threads=0)

        def linearize_into(self, T, buf):
            H, b, e = self.reg.linearize(tc, tt, shard, T)
            n_in = int((self.reg.correspondences(len(shard)) != O.NO_INDEX).sum())
            buf[:44] = torch.tensor(np.concatenate([H.ravel(), b, [e, n_in]]))

        def error_into(self, T, buf1):
            buf1[0] = self.reg.error(tc, shard, T)

    red = ShardedReduction(OracleLocal())
    T = np.eye(4)
    for it in range(20):
        H, b, e, n_in = red.linearize(T)
        if it == 0:
            first = (H.copy(), b.copy(), e, n_in, red.error(T))
        d = np.linalg.solve(H + 1e-6 * np.eye(6), -b)
        T = T @ O.se3_exp(d)
        if np.linalg.norm(d[:3]) <= 0.1 * np.pi / 180 and np.linalg.norm(d[3:]) <= 1e-3:
            break
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), H=first[0], b=first[1], e=first[2], n=first[3], err=first[4], T=T, it=it)
    dist.destroy_process_group()


def test_shard_range():
    from small_gicp_b200.distributed import shard_range

    for n in (0, 1, 7, 8, 9, 1000003):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_reduction_world2(tmp_path):
    import torch.multiprocessing as mp

    import oracle as O
    from conftest import load_golden_xyz

    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in ("H", "b", "e", "n", "err", "T", "it"):
        assert np.array_equal(r0[k], r1[k]), k  # every rank holds the same reduced system and pose
    # single-process reference
    tc = O.Cloud(load_golden_xyz("target")).voxelgrid_sampling(0.5)
    sc = O.Cloud(load_golden_xyz("source")).voxelgrid_sampling(0.5)
    tt, st = O.KdTree(tc), O.KdTree(sc)
    tt.estimate(10, O.FEAT_NORMAL_COV, 1)
    st.estimate(10, O.FEAT_COV, 1)
    reg = O.Registration(factor=O.FACTOR_GICP, num_threads=0)
    H, b, e = reg.linearize(tc, tt, sc, np.eye(4))
    assert np.linalg.norm(r0["H"] - H) <= 1e-12 * np.linalg.norm(H)
    assert np.abs(r0["b"] - b).max() <= 1e-12 * np.sqrt(2 * e * np.diag(H)).max()
    assert abs(r0["e"] - e) <= 1e-12 * e
    assert int(r0["n"]) == int((reg.correspondences(len(sc)) != O.NO_INDEX).sum())
    assert abs(float(r0["err"]) - reg.error(tc, sc, np.eye(4))) <= 1e-12 * e
    reg.set_optimizer(type=O.OPT_GN)
    ref = reg.align(tc, tt, sc, np.eye(4))
    np.testing.assert_allclose(r0["T"], ref.T_target_source, atol=1e-9)
    assert int(r0["it"]) == ref.iterations
