"""bnk_mgpu_*: several devices behind the C ABI (missing 3).  On a one-GPU box the same code runs
with the device listed twice (two workspaces, two host threads, two column blocks on one device); with two or
more visible devices it runs across them and the optimisation-loop reduction goes through ncclAllReduce."""
import numpy as np
import pytest

from helpers import NONE, leaves_of

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _problem(bnk, n_leaves, n_cols, seed, gaps=0.0):
    from bnkit_b200 import synth
    rng = np.random.default_rng(seed)
    parent, dist = synth.random_tree(n_leaves, seed=seed)
    m = bnk.Model("LG")
    rates = synth.gamma_category_rates(0.5, 4)[rng.integers(0, 4, n_cols)]
    obs = synth.simulate_alignment(parent, dist, m.parts(), n_cols, rates, seed=seed + 1, gap_fraction=gaps)
    present = None
    if gaps > 0:
        leaf = leaves_of(parent)
        present = synth.path_presence_mask(parent, (obs != NONE) & leaf[:, None])
    return m, parent, dist, rates, obs, present


def _device_lists(bnk):
    n = bnk.device_count()
    lists = [[0, 0], [0, 0, 0]]            # one device, several blocks
    if n >= 2:
        lists.append(list(range(min(n, 4))))
    return lists


def test_mgpu_matches_one_workspace_and_the_oracle(bnk, oracle):
    from bnkit_b200 import dist as bd
    m, parent, dist, rates, obs, present = _problem(bnk, 150, 517, 3, gaps=0.25)
    om = oracle.Model("LG")
    tree = bnk.Tree(parent, dist)
    n_cols = obs.shape[1]
    one = bnk.Workspace(m, tree, n_cols)
    one.set_rates(n_cols, rates)
    st1 = one.joint(obs, present)
    p1, l1 = one.marginal(obs, present)
    want = oracle.fast_joint(om, parent, dist, obs, present, rates, nthreads=8)
    for devs in _device_lists(bnk):
        g = bnk.MultiGPU(m, tree, n_cols, devices=devs)
        assert g.n_devices == len(devs)
        assert [g.shard(n_cols, i) for i in range(len(devs))] == [
            (lo, hi - lo) for lo, hi in (bd.shard_columns(n_cols, len(devs), r) for r in range(len(devs)))]
        g.set_rates(n_cols, rates)
        st = g.joint(obs, present)
        assert np.array_equal(st, st1) and np.array_equal(st, want)
        p, l = g.marginal(obs, present)
        assert np.array_equal(np.isnan(p), np.isnan(p1))
        assert np.array_equal(np.nan_to_num(p), np.nan_to_num(p1)) and np.array_equal(l, l1)
        ll, total = g.loglik(obs, present)
        assert np.array_equal(ll, l1) and np.isclose(total, l1.sum(), rtol=1e-13)
        q = tree.internal_nodes()[[0, 5, 2]].astype(np.int32)
        pq, _ = g.marginal(obs, present, query=q)
        ents = {int(v): i for i, v in enumerate(tree.internal_nodes())}
        for i, v in enumerate(q):
            assert np.array_equal(np.nan_to_num(pq[i]), np.nan_to_num(p1[ents[int(v)]]))
        assert g.launch_count() > 0
        g.close()
    one.close()


def test_mgpu_resident_loop_and_reduction(bnk, oracle):
    """the optimisation loop behind the ABI: upload once, per evaluation new rates / distances and ONE reduction of
    the per-device sums; the value equals the oracle's sum of column log-likelihoods"""
    m, parent, dist, rates, obs, _ = _problem(bnk, 200, 400, 11)
    om = oracle.Model("LG")
    tree = bnk.Tree(parent, dist)
    n_cols = obs.shape[1]
    for devs in _device_lists(bnk):
        g = bnk.MultiGPU(m, tree, n_cols, devices=devs)
        g.set_rates(n_cols, rates)
        g.upload(obs)
        d = dist.copy()
        for it in range(4):
            r = rates * (1.0 + 0.05 * it)
            g.set_rates(n_cols, r)
            if it == 2:
                d = d.copy()
                d[5] *= 1.7
                g.set_distances(d)
            total = g.loglik_resident()
            _, wl = oracle.fast_marginal(om, parent, d, obs, None, r, nthreads=8, want_post=False)
            assert np.isclose(total, wl.sum(), rtol=TOL), (devs, it)
        how = g.reduction()
        if len(set(devs)) >= 2:
            assert "ncclAllReduce" in how, how     # distinct devices: the reduction runs over NCCL
        else:
            assert "host add" in how or "single device" in how, how
        g.close()


def test_mgpu_rejects_bad_input(bnk):
    m, parent, dist, rates, obs, _ = _problem(bnk, 20, 40, 5)
    tree = bnk.Tree(parent, dist)
    g = bnk.MultiGPU(m, tree, 40, devices=[0, 0])
    with pytest.raises(bnk.BnkError):
        g.set_rates(41, np.ones(41))
    bad = obs.copy()
    bad[np.nonzero(leaves_of(parent))[0][0], 30] = 77
    with pytest.raises(bnk.BnkError) as e:
        g.joint(bad)
    assert e.value.code == -1 and "device" in str(e.value)
    with pytest.raises(bnk.BnkError):
        g.upload(bad)
    g.close()
    with pytest.raises(bnk.BnkError):
        bnk.MultiGPU(m, tree, 40, n_devices=64)
