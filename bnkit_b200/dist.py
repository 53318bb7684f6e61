"""One process per GPU (torch.distributed): the column-sharded run bench.py times at N = 1, 2, 4, 8.

Alignment columns are independent (the reference runs one thread-pool job per column,
src/asr/ThreadedDecorators.java:28-72), so the FIXED alignment is cut into contiguous column blocks, one per
rank, and no sweep contains a data-path collective.  The only exchanges are
  * an all-reduce (sum) of the per-rank log-likelihood sums, once per evaluation of an optimisation loop (the
    reduction IndelPeeler.calcProbAlnGivenTree does over columns, src/asr/IndelPeeler.java:125-150); the value is
    one double that bnk_loglik_dev leaves on the device, reduced by NCCL on the workspace's stream, and
  * a gather of the reconstructed states (uint8 [nodes][columns of the rank]).
`ShardedRun` is that logic, written against the small workspace interface of bnkit_b200.lib.Workspace
(set_rates / set_distances / joint_dev / loglik_dev ...): bench.py drives it with real workspaces over NCCL, the
CPU test (tests/test_multi_gloo.py) drives the very same class over gloo with a stand-in workspace.
The in-process multi-device counterpart behind the C ABI is bnk_mgpu_* (bnkit_b200/csrc/multi.cu); both use
the same block rule.
"""
import numpy as np


def shard_columns(n_cols, world, rank):
    """Contiguous block [lo, hi) of columns owned by `rank`: sizes differ by at most one (bnk_mgpu_shard's rule)."""
    base, rem = divmod(n_cols, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def allreduce_sum(t):
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return t


def gather_columns(local, n_cols, dst=0):
    """local: tensor [rows][cols_of_this_rank] -> on rank dst, the full [rows][n_cols]; None elsewhere."""
    import torch
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return local
    world, rank = d.get_world_size(), d.get_rank()
    widths = [shard_columns(n_cols, world, r) for r in range(world)]
    wmax = max(hi - lo for lo, hi in widths)
    pad = torch.zeros((local.shape[0], wmax), dtype=local.dtype, device=local.device)
    pad[:, : local.shape[1]] = local
    if d.get_backend() == "nccl":   # NCCL has no gather-to-one: all-gather, rank dst keeps the result
        out = torch.empty((world,) + tuple(pad.shape), dtype=pad.dtype, device=pad.device)
        d.all_gather_into_tensor(out, pad)
        bufs = list(out) if rank == dst else None
    else:
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        d.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:, : hi - lo] for b, (lo, hi) in zip(bufs, widths)], dim=1)


class ShardedRun:
    """This rank's share of a fixed alignment: its column block, the per-evaluation all-reduce of the
    log-likelihood sum, the gather of states.  `ws` needs set_rates / set_distances / joint_dev / loglik_dev."""

    def __init__(self, ws, n_cols_total, rates, obs_local, present_local, state_out, logl_out, sum_out):
        d = _dist()
        self.world = d.get_world_size() if d is not None else 1
        self.rank = d.get_rank() if d is not None else 0
        self.ws = ws
        self.n_cols_total = int(n_cols_total)
        self.lo, self.hi = shard_columns(self.n_cols_total, self.world, self.rank)
        self.ncols = self.hi - self.lo
        self.rates = None if rates is None else np.ascontiguousarray(np.asarray(rates, dtype=np.float64)[self.lo:self.hi])
        self.obs, self.present = obs_local, present_local
        self.state, self.logl, self.sum = state_out, logl_out, sum_out   # device (or host, in the CPU test) tensors
        if self.ncols > 0:
            ws.set_rates(self.ncols, self.rates)

    def evaluate(self, rate_scale=1.0, dist=None):
        """one log-likelihood evaluation of an optimisation loop: every rank sweeps its block, then ONE all-reduce
        of the per-rank sums; returns the tensor holding sum_c logL_c of the whole alignment (on every rank)"""
        if dist is not None:
            self.ws.set_distances(dist)
        if self.ncols > 0:
            if rate_scale != 1.0 or self.rates is not None:
                base = np.ones(self.ncols) if self.rates is None else self.rates
                self.ws.set_rates(self.ncols, base * rate_scale)
            self.ws.loglik_dev(self.ncols, self.obs, self.present, self.logl, self.sum)
        else:
            self.sum.zero_()
        return allreduce_sum(self.sum)

    def joint_and_gather(self, dst=0):
        """joint reconstruction of the block, then the gather of states to rank dst ([nodes][n_cols_total])"""
        if self.ncols > 0:
            self.ws.joint_dev(self.ncols, self.obs, self.present, self.state)
        return gather_columns(self.state, self.n_cols_total, dst=dst)
