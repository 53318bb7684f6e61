"""GPU parity on the reference's own data files (committed fixtures), through the Java-API mirror,
and size-independent properties at the full BASELINE.json size."""
import json
import os

import numpy as np
import pytest

from helpers import NONE, leaves_of, random_obs, random_tree

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
AA = "ARNDCQEGHILKMFPSTWYV"


def _dataset(name):
    from bnkit_b200 import synth
    d = json.load(open(os.path.join(GOLD, name)))
    parent = np.array(d["tree"]["Parents"], dtype=np.int32)
    dist = np.array(d["tree"]["Distances"])
    labels = d["tree"]["Labels"]
    n, C = len(parent), len(d["seqs"][0])
    obs = np.full((n, C), NONE, np.uint8)
    for name_, seq in zip(d["names"], d["seqs"]):
        v = labels.index(name_)
        obs[v] = [NONE if ch == "-" else AA.index(ch) for ch in seq]
    leaf = leaves_of(parent)
    present = synth.path_presence_mask(parent, (obs != NONE) & leaf[:, None])
    return d, parent, dist, obs, present


def test_c1_documented_result_and_oracle_parity(bnk, oracle):
    """BASELINE configs[0]: data/66.*; docs/json-api.md:113-122 pins N0..N2 = A, H at columns 2, 4 (JTT)."""
    d, parent, dist, obs, present = _dataset("c1_66.json")
    for name in ("JTT", "LG"):
        m, om = bnk.Model(name), oracle.Model(name)
        ws = bnk.Workspace(m, bnk.Tree(parent, dist), obs.shape[1])
        st = ws.joint(obs, present)
        assert np.array_equal(st, oracle.fast_joint(om, parent, dist, obs, present))
        if name == "JTT":
            for anc in (0, 1, 2):
                cols = [c for c in range(obs.shape[1]) if present[anc, c]]
                assert cols == d["doc_result"]["indices"]
                assert [AA[st[anc, c]] for c in cols] == d["doc_result"]["values"]
        ws.close()


def test_c2_marginal_all_ancestors(bnk, oracle):
    """BASELINE configs[1]: data/3_2_1_1_filt.*, JTT, marginal at all 37 ancestors x 1307 columns, on the mask the
    reference's own pipeline produces: bi-directional edge parsimony (device BEP, checked against the oracle's
    restatement) then orphan pruning -- not the stand-in rule."""
    from oracle import bep as obep
    d, parent, dist, obs, standin = _dataset("c2_3_2_1_1_filt.json")
    m, om = bnk.Model("JTT"), oracle.Model("JTT")
    tree = bnk.Tree(parent, dist)
    raw = tree.bep_presence(obs)
    leaf = leaves_of(parent)
    rows = [(obs[v] != NONE).tolist() if leaf[v] else None for v in range(len(parent))]
    want_mask, _ = obep.presence_mask(parent.tolist(), rows, obs.shape[1])
    assert np.array_equal(raw, np.asarray(want_mask, np.uint8))
    present = tree.prune_orphans(raw)
    assert (present != standin).any(), "BEP and the stand-in rule agree everywhere: the switch tests nothing new"
    assert tree.n == 76 and tree.n_internal == 37 and obs.shape[1] == 1307
    ws = bnk.Workspace(m, tree, obs.shape[1])
    post, logl = ws.marginal(obs, present)
    wpost, wlogl = oracle.fast_marginal(om, parent, dist, obs, present, nthreads=8)
    wp = wpost[tree.internal_nodes()]
    assert np.array_equal(np.isnan(post), np.isnan(wp))
    ok = ~np.isnan(wp) & (np.abs(wp) >= 1e-300)
    assert (np.abs(post[ok] - wp[ok]) <= 1e-9 * np.abs(wp[ok])).all()   # relative, no floor
    assert np.allclose(logl, wlogl, rtol=1e-9, atol=1e-12)
    assert np.array_equal(ws.joint(obs, present), oracle.fast_joint(om, parent, dist, obs, present, nthreads=8))
    # gap-free variant (no mask): gaps replaced by drawn residues
    rng = np.random.default_rng(6)
    leaf = leaves_of(parent)
    obs2 = obs.copy()
    fill = (obs2 == NONE) & leaf[:, None]
    obs2[fill] = rng.integers(0, 20, fill.sum())
    post2, logl2 = ws.marginal(obs2)
    wpost2, wlogl2 = oracle.fast_marginal(om, parent, dist, obs2, nthreads=8)
    assert np.allclose(post2, wpost2[tree.internal_nodes()], rtol=1e-9, atol=1e-12)
    assert np.allclose(logl2, wlogl2, rtol=1e-9)
    ws.close()


def test_java_api_mirror_against_bruteforce(bnk, oracle):
    """The reference's own test method (MaxLhoodJointTest.joint1, MaxLhoodMarginalTest.marginal1) driven
    through the mirrored classes: TreeInstance -> MaxLhoodJoint / MaxLhoodMarginal.decorate -> getDecoration."""
    from bnkit_b200.phylo import ASRRuntimeException, IdxTree, MaxLhoodJoint, MaxLhoodMarginal, SubstModel, TreeInstance
    model = SubstModel.createModel("JC")
    om = oracle.Model("JC")
    dom = model.getDomain()
    for seed in range(12):
        rng = np.random.default_rng(300 + seed)
        parent, dist = random_tree(rng, int(rng.integers(3, 8)))
        tree = IdxTree(parent, dist)
        assign = [dom[int(rng.integers(0, 4))] if tree.isLeaf(i) else None for i in range(tree.getSize())]
        ti = TreeInstance(tree, assign)
        obs = np.array([NONE if a is None else dom.index(a) for a in assign], np.uint8)
        b = oracle.brute(om, parent, dist, obs)
        mlj = MaxLhoodJoint(tree, model)
        mlj.decorate(ti)
        got = [mlj.getDecoration(i) for i in range(tree.getSize())]
        want = [dom[s] for s in oracle.fast_joint(om, parent, dist, obs[:, None])[:, 0]]
        assert got == want
        mlm = MaxLhoodMarginal(0, tree, model)
        mlm.decorate(ti)
        assert np.allclose(mlm.getDecoration(0).distrib, b["post"][0], atol=1e-9)
        with pytest.raises(ASRRuntimeException):
            mlm.getDecoration(1)


def test_prediction_batch_entry_points(bnk, oracle):
    from bnkit_b200.phylo import ASRRuntimeException, IdxTree, Prediction, SubstModel
    d, parent, dist, obs, present = _dataset("c1_66.json")
    tree = IdxTree.fromJSON(d["tree"])
    states = [[None] * obs.shape[1] for _ in range(tree.getSize())]
    for name_, seq in zip(d["names"], d["seqs"]):
        states[tree.getIndex(name_)] = [None if ch == "-" else ch for ch in seq]
    model = SubstModel.createModel("LG")
    pred = Prediction(tree, states, present)
    joint = pred.getJoint(model)
    want = oracle.fast_joint(oracle.Model("LG"), parent, dist, obs, present)
    leaf = leaves_of(parent)
    # ancestors' rows are filled, extants' rows stay null as in the reference (Prediction.java:636-641)
    assert [[None if (s == NONE or leaf[v]) else AA[s] for s in row] for v, row in enumerate(want)] == joint
    marg = pred.getMarginal("N0", model)
    assert [x is None for x in marg] == [not present[0, c] for c in range(obs.shape[1])]
    with pytest.raises(ASRRuntimeException):
        pred.getMarginal("N999", model)
    rates = [0.5, 1.0, 2.0, 1.0, 0.25]
    j2 = pred.getJoint(model, rates)
    w2 = oracle.fast_joint(oracle.Model("LG"), parent, dist, obs, present, rates)
    assert [[None if (s == NONE or leaf[v]) else AA[s] for s in row] for v, row in enumerate(w2)] == j2
    pred.close()


def test_documented_recon_through_the_java_api_mirror(bnk):
    """docs/json-api.md:97-122 end to end through the reference-shaped calls: Prediction.PredictByBidirEdgeParsimony
    (device BEP) then getJoint under JTT gives ancestors 0, 1, 2 the states A, H at columns 2, 4 and null elsewhere."""
    from bnkit_b200.phylo import IdxTree, Prediction, SubstModel
    d, parent, dist, obs, _ = _dataset("c1_66.json")
    tree = IdxTree.fromJSON(d["tree"])
    states = [[None] * obs.shape[1] for _ in range(tree.getSize())]
    for name_, seq in zip(d["names"], d["seqs"]):
        states[tree.getIndex(name_)] = [None if ch == "-" else ch for ch in seq]
    pred = Prediction.PredictByBidirEdgeParsimony(tree, states)
    joint = pred.getJoint(SubstModel.createModel("JTT"))
    doc = d["doc_result"]
    for anc in (0, 1, 2):
        assert tree.getLabel(anc) in (str(anc), "N%d" % anc)   # branch points 0, 1, 2 are ancestors "0", "1", "2"
        row = joint[anc]
        assert [c for c, x in enumerate(row) if x is not None] == doc["indices"]
        assert [x for x in row if x is not None] == doc["values"]
    pred.close()


def test_full_size_properties_c3(bnk):
    """BASELINE configs[2] at full size (10,000 leaves x 5,000 columns, LG + Gamma4): properties that
    need no oracle -- tile-geometry invariance (bit-identical states), column-permutation
    equivariance, posteriors sum to 1, sum of column log-likelihoods == the reduced sum."""
    from bnkit_b200 import synth
    n_cols = 5000
    parent, dist = synth.balanced_tree(10000, seed=1)
    m = bnk.Model("LG")
    r4 = synth.gamma_category_rates(0.5, 4)
    rng = np.random.default_rng(4)
    rates = r4[rng.integers(0, 4, n_cols)]
    leaf = leaves_of(parent)
    obs = np.full((len(parent), n_cols), NONE, np.uint8)
    obs[leaf] = rng.integers(0, 20, (leaf.sum(), 1)) ^ (rng.random((leaf.sum(), n_cols)) < 0.15).astype(np.uint8)
    tree = bnk.Tree(parent, dist)
    ws = bnk.Workspace(m, tree, n_cols)
    ws.set_rates(n_cols, rates)
    st = ws.joint(obs)
    logl, total = ws.loglik(obs)
    assert np.isfinite(logl).all() and np.isclose(total, logl.sum(), rtol=1e-12)
    assert (st[~leaf] < 20).all() and np.array_equal(st[leaf], obs[leaf])
    # other tile geometry and the general kernels give bit-identical states
    for tile, cpt in ((34, 2), (11, -1), (-17, 1)):
        ws2 = bnk.Workspace(m, tree, n_cols)
        ws2.configure(tile, cpt)
        ws2.set_rates(n_cols, rates)
        assert np.array_equal(ws2.joint(obs), st)
        l2, _ = ws2.loglik(obs)
        assert np.allclose(l2, logl, rtol=1e-12)
        ws2.close()
    # permuting columns permutes results
    perm = rng.permutation(n_cols)
    ws.set_rates(n_cols, rates[perm])
    assert np.array_equal(ws.joint(np.ascontiguousarray(obs[:, perm])), st[:, perm])
    # posteriors of a subset of ancestors are distributions
    q = tree.internal_nodes()[::997]
    post, _ = ws.marginal(np.ascontiguousarray(obs[:, perm]), query=q)
    assert np.allclose(post.sum(-1), 1.0, atol=1e-12) and (post >= 0).all()
    ws.close()


def test_c4_full_size_properties(bnk):
    """BASELINE configs[3]: 100,000-leaf random tree x 2,000 columns, WAG (one GPU's share and more):
    hybrid schedule vs walker-only vs general kernels give bit-identical states and equal logL."""
    from bnkit_b200 import synth
    n_cols = 2000
    parent, dist = synth.random_tree(100000, seed=5)
    m = bnk.Model("WAG")
    rng = np.random.default_rng(8)
    leaf = leaves_of(parent)
    obs = np.full((len(parent), n_cols), NONE, np.uint8)
    base = rng.integers(0, 20, (int(leaf.sum()), 1)).astype(np.uint8)
    noise = rng.integers(0, 20, (int(leaf.sum()), n_cols)).astype(np.uint8)
    obs[leaf] = np.where(rng.random((int(leaf.sum()), n_cols)) < 0.2, noise, base)
    tree = bnk.Tree(parent, dist)
    assert tree.n == 199999 and tree.n_internal == 99999
    ws = bnk.Workspace(m, tree, n_cols)
    st = ws.joint(obs)
    logl, total = ws.loglik(obs)
    assert (st[~leaf] < 20).all() and np.array_equal(st[leaf], obs[leaf])
    assert np.isfinite(logl).all() and np.isclose(total, logl.sum(), rtol=1e-12)
    ws.close()
    ws2 = bnk.Workspace(m, tree, n_cols)
    ws2.configure(-24, 2)   # walker only
    assert np.array_equal(ws2.joint(obs), st)
    l2, _ = ws2.loglik(obs)
    assert np.allclose(l2, logl, rtol=1e-11)
    ws2.close()
    # a column subset through the general kernels (masks all-present => same answer)
    sub = np.ascontiguousarray(obs[:, :64])
    ws3 = bnk.Workspace(m, tree, 64)
    st3 = ws3.joint(sub, np.ones_like(sub))
    assert np.array_equal(st3, st[:, :64])
    ws3.close()


def test_c5_likelihood_loop_matches_oracle(bnk, oracle):
    """BASELINE configs[4]: repeated log-likelihood evaluations with perturbed per-column rates and a
    perturbed branch length; the per-evaluation sum (the value the NCCL all-reduce carries) vs the oracle."""
    from bnkit_b200 import synth
    rng = np.random.default_rng(19)
    parent, dist = synth.balanced_tree(300, seed=2)
    m, om = bnk.Model("LG"), oracle.Model("LG")
    n_cols = 160
    obs = synth.simulate_alignment(parent, dist, m.parts(), n_cols, None, seed=4)
    ws = bnk.Workspace(m, bnk.Tree(parent, dist), n_cols)
    r4 = synth.gamma_category_rates(0.5, 4)
    cat = rng.integers(0, 4, n_cols)
    d = dist.copy()
    for it in range(5):
        rates = (r4 * (1.0 + 0.1 * it))[cat]
        if it >= 3:
            d = d.copy()
            d[7] *= 1.5
            ws.set_distances(d)
        ws.set_rates(n_cols, rates)
        logl, total = ws.loglik(obs)
        _, wl = oracle.fast_marginal(om, parent, d, obs, None, rates, nthreads=8, want_post=False)
        assert np.allclose(logl, wl, rtol=1e-9) and np.isclose(total, wl.sum(), rtol=1e-9)
    ws.close()
