"""Worker of tests/test_gpu_multiprocess.py: one process per GPU (torchrun), source sharded, target replicated.  Every rank checks
  (1) the sums of the exchange fused into the reduction kernel (peer mailboxes over CUDA IPC / NVLink) against NCCL's all_reduce of the
      per-rank LOCAL sums (a second, unconnected context on the same shard): <= 1e-12 relative;
  (2) the same sums against ONE context holding the whole source on rank 0's GPU: <= 1e-5 (shards re-centre their own boxes);
  (3) a whole Gauss-Newton align driven through the fused exchange: bit-identical pose on every rank; against the un-sharded context the
      pose agrees to 1e-6 rad / 2e-6 m (every shard stores its FP32 coordinates relative to its OWN box centre, so the roundings differ:
      measured 0 rad / 2.4e-7 m with four shards of 30k points; BASELINE's bound is 1e-4 rad / 1e-3 m);
  (4) Reduction::error through the exchange.
Exit code 0 = all good (assertion failures raise)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import bench as B
    import small_gicp_b200 as sg
    from small_gicp_b200.distributed import connect_fused, shard_range
    from small_gicp_b200.synthetic import make_pair

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    n = 120_000
    tgt, src, Tgt = make_pair(n)
    scratch = sg.Context(local)
    tcov = scratch.estimate_features(tgt, 20, normals=False)[1]
    scov = scratch.estimate_features(src, 20, normals=False)[1]
    scratch.close()
    lo, hi = shard_range(n, rank, world)

    def context(a, b):
        c = sg.Context(local)
        c.set_stream(stream.cuda_stream)
        c.set_target(tgt, None, tcov)
        c.build_target_kdtree(0)
        c.set_source(src[a:b], scov[a:b])
        return c

    fused, local_ctx, whole = context(lo, hi), context(lo, hi), context(0, n)
    assert connect_fused(fused)
    out = torch.zeros(64, dtype=torch.float64, device=dev)
    out2 = torch.zeros(64, dtype=torch.float64, device=dev)
    for T in (np.eye(4), Tgt):
        fused.linearize_device(T, out.data_ptr())
        local_ctx.linearize_device(T, out2.data_ptr())
        dist.all_reduce(out2[:44])
        a, b = out[:44].cpu().numpy(), out2[:44].cpu().numpy()
        assert np.isfinite(a).all()
        assert np.linalg.norm(a[:36] - b[:36]) <= 1e-12 * np.linalg.norm(b[:36]) and abs(a[42] - b[42]) <= 1e-12 * abs(b[42]) and a[43] == b[43], (rank, "fused vs NCCL")
        H0, b0, e0 = whole.linearize(T)
        assert np.linalg.norm(a[:36].reshape(6, 6) - H0) <= 1e-5 * np.linalg.norm(H0) and abs(a[42] - e0) <= 1e-5 * e0, (rank, "sharded vs whole")
        # every rank holds the bit-identical sum
        g = [torch.zeros(44, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(g, out[:44].clone())
        for t in g:
            assert torch.equal(t, g[0])
        # error() through the exchange
        T2 = T @ B.se3_exp(np.array([1e-3, -2e-3, 1e-3, 0.01, 0.02, -0.01]))
        e_f = fused.error(T2)
        local_ctx.error_device(T2, out2.data_ptr())
        dist.all_reduce(out2[:1])
        assert abs(e_f - float(out2[0].cpu())) <= 1e-12 * abs(e_f)
    # a whole align through the collective calls
    poses_f, T_f = B.gn_trajectory(lambda T: fused.linearize(T))
    poses_w, T_w = B.gn_trajectory(lambda T: whole.linearize(T))
    rot, trans = B.pose_error(T_w, T_f)
    assert len(poses_f) == len(poses_w) and rot < 1e-6 and trans < 2e-6, (rot, trans)
    tt = torch.from_numpy(T_f.copy()).to(dev)
    g = [torch.zeros_like(tt) for _ in range(world)]
    dist.all_gather(g, tt)
    for t in g:
        assert torch.equal(t, g[0])
    assert fused.comm_status() == 0
    dist.barrier()
    for c in (fused, local_ctx, whole):
        c.close()
    dist.destroy_process_group()
    if rank == 0:
        print(f"mp_fused_check ok: world {world}, pose vs un-sharded {rot:.1e} rad / {trans:.1e} m, {len(poses_f)} GN iterations")


if __name__ == "__main__":
    main()
