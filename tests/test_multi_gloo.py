"""The N > 1 path on CPU: two gloo ranks run bnkit_b200.dist.ShardedRun -- the class bench.py drives over NCCL --
on their column blocks.  There is no GPU here, so the workspace it is given is a stand-in with the same small
interface (set_rates / set_distances / joint_dev / loglik_dev) that answers from the CPU oracle; everything
else -- the block rule, the per-evaluation all-reduce of the log-likelihood sum, the gather of states -- is the
shipped code path, and must reproduce the unsharded result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class StandInWorkspace:
    """lib.Workspace's device-resident interface on host tensors (CPU oracle underneath): test infrastructure"""

    def __init__(self, O, model, parent, dist_):
        self.O, self.model, self.parent, self.dist = O, model, parent, np.array(dist_, dtype=np.float64)
        self.rates = None

    def set_rates(self, n_cols, rates=None):
        self.rates = None if rates is None else np.array(rates, dtype=np.float64)
        assert self.rates is None or len(self.rates) == n_cols

    def set_distances(self, d):
        self.dist = np.array(d, dtype=np.float64)

    def joint_dev(self, n_cols, obs, present, out_state):
        st = self.O.fast_joint(self.model, self.parent, self.dist, obs.numpy(), None if present is None else present.numpy(), self.rates)
        out_state.copy_(torch.from_numpy(st))

    def loglik_dev(self, n_cols, obs, present, out_logl, out_sum):
        _, logl = self.O.fast_marginal(self.model, self.parent, self.dist, obs.numpy(),
                                       None if present is None else present.numpy(), self.rates, want_post=False)
        out_logl.copy_(torch.from_numpy(logl))
        out_sum[0] = float(logl.sum())


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bnkit_b200 import dist as bd
        from helpers import random_obs, random_tree
        from oracle import oracle as O
        rng = np.random.default_rng(21)                     # same inputs on every rank
        parent, dst = random_tree(rng, 30)
        n_cols = 101
        obs = random_obs(rng, parent, n_cols, 20)
        rates = rng.choice([0.5, 1.0, 2.0], size=n_cols)
        m = O.Model("LG")
        lo, hi = bd.shard_columns(n_cols, world, rank)
        o = torch.from_numpy(np.ascontiguousarray(obs[:, lo:hi]))
        run = bd.ShardedRun(StandInWorkspace(O, m, parent, dst), n_cols, rates, o, None,
                            torch.empty((len(parent), hi - lo), dtype=torch.uint8), torch.empty(hi - lo, dtype=torch.float64),
                            torch.zeros(1, dtype=torch.float64))
        assert (run.lo, run.hi) == (lo, hi)
        totals = []
        d2 = dst.copy()
        d2[3] *= 2.0
        for scale, dd in ((1.0, None), (1.3, None), (1.0, d2)):
            totals.append(float(run.evaluate(rate_scale=scale, dist=dd)[0]))
        full = run.joint_and_gather(dst=0)
        if rank == 0:
            want_tot = []
            for scale, dd in ((1.0, dst), (1.3, dst), (1.0, d2)):
                _, wl = O.fast_marginal(m, parent, dd, obs, None, rates * scale, want_post=False)
                want_tot.append(float(wl.sum()))
            want = O.fast_joint(m, parent, d2, obs, None, rates)
            q.put((bool(np.array_equal(full.numpy(), want)), totals, want_tot))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def test_block_rule_matches_the_c_abi(bnk):
    """bnkit_b200.dist.shard_columns == bnk_mgpu_shard's rule: contiguous blocks, sizes differ by at most one
    (the C side is compared with it on the GPU box, tests/test_gpu_multi.py)"""
    from bnkit_b200 import dist as bd
    for n_cols, world in ((5000, 8), (2000, 8), (101, 2), (7, 8), (1, 4)):
        blocks = [bd.shard_columns(n_cols, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n_cols
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_column_sharding_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    same, totals, want = q.get(timeout=10)
    assert same and np.allclose(totals, want, rtol=1e-12)
