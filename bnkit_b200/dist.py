"""Multi-GPU plumbing for the column-sharded path (one process per GPU, torch.distributed).

Alignment columns are independent (the reference runs one thread-pool job per column,
src/asr/ThreadedDecorators.java:28-72), so ranks own contiguous column blocks and there is no
data-path collective inside a sweep.  The only exchanges are
  * an all-reduce (sum) of the per-rank log-likelihood sums, once per evaluation of an
    optimisation loop (the reduction IndelPeeler.calcProbAlnGivenTree does over columns,
    src/asr/IndelPeeler.java:125-150), and
  * a gather of reconstructed states to rank 0.
Works with the NCCL backend (device tensors) and with gloo (CPU tensors, used by the tests).
"""
import numpy as np


def shard_columns(n_cols, world, rank):
    """Contiguous block [lo, hi) of columns owned by `rank`: sizes differ by at most one."""
    base, rem = divmod(n_cols, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum(t):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def gather_columns(local, n_cols, dst=0):
    """local: tensor [rows][cols_of_this_rank] -> on rank dst, the full [rows][n_cols]; None elsewhere."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    widths = [shard_columns(n_cols, world, r) for r in range(world)]
    wmax = max(hi - lo for lo, hi in widths)
    pad = torch.zeros((local.shape[0], wmax), dtype=local.dtype, device=local.device)
    pad[:, : local.shape[1]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:, : hi - lo] for b, (lo, hi) in zip(bufs, widths)], dim=1)
