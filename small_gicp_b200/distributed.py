"""Multi-GPU plumbing of the hot path: one process per GPU, source points sharded in contiguous
ranges, target + search structure replicated, ONE all-reduce(sum) of H|b|e|num_inliers (44 doubles) per linearize and
of e (1 double) per error.  All ranks then solve the same 6x6 system, so no broadcast of the update is needed.

Two transports:
  * fused (GPUs, default of bench.py): `connect_fused(ctx)` wires the contexts' mailboxes through CUDA IPC; the CTA that
    finishes the reduction kernel then writes its sums into the peers' mailboxes over NVLink and adds theirs -- the
    all-reduce is part of the kernel, torch.distributed only carried the 64-byte handles once;
  * torch.distributed all_reduce (NCCL on GPUs, gloo in the CPU tests) on the 44 doubles the local reduction left in
    device memory.
The per-rank reduction is any object with linearize_into(T, out44) / error_into(T, out1) writing into a torch tensor
(GPU: small_gicp_b200.Context writing device memory on the current stream; tests: the CPU oracle)."""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous, balanced split of range(n): the first n % world ranks get one extra element."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def connect_fused(ctx, group=None):
    """Wire the mailboxes of every rank's Context (one process per GPU, same node) through CUDA IPC.  Afterwards
    ctx.linearize*/error* return the sum over all ranks with no further collective call.  Collective: every rank calls it."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return False
    handles = [None] * world
    dist.all_gather_object(handles, ctx.comm_handle(), group=group)
    ctx.comm_connect(rank, world, handles)
    dist.barrier(group=group)  # nobody starts exchanging before every mailbox is cleared and mapped
    return True


class ShardedReduction:
    """Reduction::linearize / ::error over a source sharded across the ranks of a process group.
    fused=True: `local` already returns global sums (connect_fused was called on its context): no collective here."""

    def __init__(self, local, device=None, group=None, fused=False):
        import torch
        import torch.distributed as dist

        self.local = local
        self.dist = dist
        self.group = group
        self.fused = fused
        self.buf = torch.zeros(64, dtype=torch.float64, device=device)

    def _collective(self):
        return (not self.fused) and self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1

    def linearize(self, T, **kw):
        """-> (H 6x6, b 6, e, num_inliers) identical on every rank"""
        self.local.linearize_into(T, self.buf, **kw)
        if self._collective():
            self.dist.all_reduce(self.buf[:44], op=self.dist.ReduceOp.SUM, group=self.group)
        h = self.buf[:44].cpu().numpy()
        return h[:36].reshape(6, 6).copy(), h[36:42].copy(), float(h[42]), int(round(h[43]))

    def error(self, T):
        self.local.error_into(T, self.buf[48:49])
        if self._collective():
            self.dist.all_reduce(self.buf[48:49], op=self.dist.ReduceOp.SUM, group=self.group)
        return float(self.buf[48].cpu())


class ContextReduction:
    """Adapter: a small_gicp_b200.Context whose source holds this rank's shard; writes straight into device memory."""

    def __init__(self, ctx, factor, robust=0, robust_c=1.0, rejector=1, max_dist_sq=1.0, bind_stream=True):
        """bind_stream: run the context on torch's CURRENT stream (the one ShardedReduction's all_reduce / .cpu() are ordered on).
        A context keeps its own non-blocking stream otherwise, and nothing would order the collective after the kernel."""
        import torch

        self.ctx, self.kw = ctx, dict(factor=factor, robust=robust, robust_c=robust_c, rejector=rejector, max_dist_sq=max_dist_sq)
        self.bound = bool(bind_stream) and torch.cuda.is_available()
        if self.bound:
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    def _order(self):
        if not self.bound:  # the context runs on a stream torch knows nothing about: finish its work before torch touches the buffer
            self.ctx.synchronize()

    def linearize_into(self, T, buf, **kw):
        self.ctx.linearize_device(T, buf.data_ptr(), **dict(self.kw, **kw))
        self._order()

    def error_into(self, T, buf1):
        self.ctx.error_device(T, buf1.data_ptr())
        self._order()
