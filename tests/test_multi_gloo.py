"""The N > 1 path on CPU: two gloo ranks shard the columns, each evaluates its block (here with
the CPU oracle standing in for the per-rank GPU workspace), and the only collectives -- the
log-likelihood all-reduce and the state gather -- reproduce the unsharded result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
        o = np.ascontiguousarray(obs[:, lo:hi])
        st = O.fast_joint(m, parent, dst, o, None, rates[lo:hi])
        _, logl = O.fast_marginal(m, parent, dst, o, None, rates[lo:hi], want_post=False)
        total = bd.allreduce_sum(torch.tensor([logl.sum()], dtype=torch.float64))
        full = bd.gather_columns(torch.from_numpy(st), n_cols, dst=0)
        if rank == 0:
            want = O.fast_joint(m, parent, dst, obs, None, rates)
            _, wl = O.fast_marginal(m, parent, dst, obs, None, rates, want_post=False)
            q.put((bool(np.array_equal(full.numpy(), want)), float(total[0]), float(wl.sum())))
    finally:
        dist.destroy_process_group()


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
    same, total, want = q.get(timeout=10)
    assert same and np.isclose(total, want, rtol=1e-12)
