"""Gap-augmented peeling: bn.ctmc.GapSubstModel + asr.IndelPeeler.

CPU: the oracle (oracle/bnk_indel.c, the reference's log-space recursion restated literally) against the
reference's own known answers -- test/bn/ctmc/GapSubstModelTest.java (closed-form extended JC69, no-indel
equivalences) -- an independent matrix exponential, the plain Felsenstein likelihood when mu = lambda = 0, and an
independent linear-space restatement of the recursion written here.
GPU: bnk_indel_* (C ABI) against the oracle, 1e-9 relative on every column's log-likelihood."""
import numpy as np
import pytest

from helpers import NONE, leaves_of, random_tree
from test_bep import fixture_rows, gold

TOL = 1e-9


def _jc_irm():
    irm = np.zeros((4, 4))
    for i in range(4):
        for j in range(i + 1, 4):
            irm[i, j] = 0.333333
    return irm


def test_gap_model_known_answers_of_the_reference(oracle):
    """GapSubstModelTest.java: testConditionalProbabilitiesJCGap (eq. 59 of Rivas & Eddy 2008, Sup. A1; 1e-5 as
    the reference asserts), ...JCNonGap, ...JCNoIndels, ...JTTNoIndels, testStationaryProbabilities*, testGapProbabilities"""
    base = oracle.Model(custom=(4, [0.25] * 4, _jc_irm(), 1, 1))
    mu, lam, t, alpha = 0.01, 0.02, 0.05, 0.3333333333
    P = oracle.GapModel(base, mu, lam).probs(t)
    for x in range(4):
        for y in range(4):
            want = 0.25 * (lam / (lam + mu)) + ((x == y) - 0.25) * np.exp(-(4 * alpha + mu) * t) + 0.25 * (mu / (lam + mu)) * np.exp(-(lam + mu) * t)
            assert abs(P[y, x] - want) < 1e-5
            exact = 0.25 * (lam / (lam + mu)) + ((x == y) - 0.25) * np.exp(-(4 / 3 + mu) * t) + 0.25 * (mu / (lam + mu)) * np.exp(-(lam + mu) * t)
            assert abs(P[y, x] - exact) < 1e-14          # the same formula at alpha = 1/3 exactly (the constructor normalises R)
    P0 = oracle.GapModel(base, 0, 0).probs(t)
    for x in range(4):
        for y in range(4):
            assert abs(P0[y, x] - (0.25 + ((x == y) - 0.25) * np.exp(-4 * alpha * t))) < 1e-5
    for name in ("JC", "JTT"):
        m = oracle.Model(name)
        g0 = oracle.GapModel(m, 0, 0)
        K = m.K
        assert np.abs(g0.probs(0.5)[:K, :K] - m.probs(0.5)).max() < 1e-5
        assert np.abs(g0.parts()["F"][:K] - m.parts()["F"]).max() < 1e-5
        assert g0.parts()["F"][K] == 0.0
    g = oracle.GapModel(oracle.Model("JTT"), 0.05, 0.05)
    F = g.parts()["F"]
    assert abs(F[20] - 0.5) < 1e-15 and abs(F[:20].sum() - 0.5) < 1e-6   # gap frequency mu / (mu + lambda)
    assert abs(g.ksi(0.05) - 0.5 * (1 - np.exp(-0.1 * 0.05))) < 1e-15 and g.ksi(0.05) == g.gamma(0.05)


@pytest.mark.parametrize("name", ["LG", "JTT", "WAG", "Dayhoff", "JC", "Yang"])
def test_gap_model_against_an_independent_matrix_exponential(oracle, name):
    """P_eps(t) = exp(R_eps t) by scipy.linalg.expm; rows sum to 1; gap column = (1 - e^{-(mu + lambda) t}) mu / (mu + lambda)"""
    from scipy.linalg import expm
    m = oracle.Model(name)
    g = oracle.GapModel(m, 0.07, 0.03)
    R = g.parts()["R"]
    K = m.K
    # residue rows sum to 0; the gap row sums to lambda * (sum(F) - 1): the published frequencies are rounded
    assert np.abs(R[:K].sum(1)).max() < 1e-12 and abs(R[K].sum()) < 1e-5 * 0.03
    assert np.allclose(R[:K, K], 0.07) and np.allclose(R[K, :K], 0.03 * m.parts()["F"])
    for t in (1e-5, 0.01, 0.3, 2.0, 20.0):
        P = g.probs(t)
        assert np.abs(P - expm(R * t)).max() < 1e-12
        assert np.abs(P.sum(1) - 1).max() < 1e-6
        assert np.abs(P[:K, K] - g.gamma(t)).max() < 1e-6


def _felsenstein_gap_linear(gm, parent, dist, obs_col, rate, geom):
    """the recursion of IndelPeeler (src/asr/IndelPeeler.java:229-356) in linear space with Python floats /
    fractions of numpy -- an independent second statement for small trees"""
    n = len(parent)
    K = gm.K1 - 1
    F = gm.parts()["F"]
    kids = [[] for _ in range(n)]
    for v, p in enumerate(parent):
        if p >= 0:
            kids[p].append(v)
    t = [0.0 if parent[v] < 0 else dist[v] * rate for v in range(n)]
    R = np.zeros((n, K))
    G = np.zeros(n)
    cg = np.zeros(n, bool)
    for v in range(n - 1, -1, -1):
        if not kids[v]:
            cg[v] = obs_col[v] == NONE
            if not cg[v]:
                R[v, obs_col[v]] = 1.0
            continue
        cg[v] = all(cg[c] for c in kids[v])
        ksi_v = gm.ksi(t[v])
        G[v] = sum((R[c] * (ksi_v * F[:K])).sum() + G[c] for c in kids[v] if cg[c])
        R[v] = 1.0
        for c in kids[v]:
            noins = 1 - gm.ksi(t[c])
            A = gm.probs(t[c])[:K, :K] * noins
            R[v] *= A @ R[c] + (noins * gm.gamma(t[c]) if cg[c] else 0.0)
    return np.log(G[0] + geom * (F[:K] * R[0]).sum())


def _random_alignment(rng, parent, n_cols, K, gap):
    leaf = leaves_of(parent)
    obs = np.full((len(parent), n_cols), NONE, np.uint8)
    vals = rng.integers(0, K, (len(parent), n_cols)).astype(np.uint8)
    obs[leaf] = vals[leaf]
    obs[(rng.random(obs.shape) < gap) & leaf[:, None]] = NONE
    return obs


def _clade_gaps(rng, parent, obs, cols, per_col=1):
    """whole clades lose a column (nodes are indexed depth-first: a subtree is a contiguous index range)"""
    n = len(parent)
    size = np.ones(n, np.int64)
    for v in range(n - 1, 0, -1):
        size[parent[v]] += size[v]
    for c in cols:
        for _ in range(per_col):
            v = int(rng.integers(1, n))
            obs[v:v + size[v], c] = NONE


def test_oracle_peeling_against_an_independent_linear_space_statement(oracle):
    rng = np.random.default_rng(5)
    checked = 0
    for trial in range(12):
        name = ["JTT", "LG", "JC", "Yang"][trial % 4]
        m = oracle.Model(name)
        parent, dist = random_tree(rng, int(rng.integers(2, 12)), max_children=int(rng.integers(2, 5)))
        gm = oracle.GapModel(m, *[(0.05, 0.03), (0.2, 0.2), (0.0, 0.0), (0.3, 0.001)][trial % 4])
        obs = _random_alignment(rng, parent, 14, m.K, float(rng.choice([0.0, 0.3, 0.7, 0.95])))
        obs[:, 0] = NONE                                       # an all-gap column
        rate, geom = float(rng.choice([0.5, 1.0, 2.0])), 0.02
        got = oracle.indel_column_logl(gm, parent, dist, obs, rate, geom)
        for c in range(obs.shape[1]):
            with np.errstate(divide="ignore"):
                want = _felsenstein_gap_linear(gm, parent, dist, obs[:, c], rate, geom)
            if np.isinf(want):
                assert got[c] == want
            else:
                assert abs(got[c] - want) <= 1e-10 * abs(want), (trial, c, got[c], want)
            checked += 1
    assert checked > 150


def test_no_indels_is_plain_felsenstein(oracle):
    """mu = lambda = 0 on gap-free columns: log P(column) = log(geom) + the log-likelihood of VarElim's
    sum-product (oracle.fast_marginal -- separate code, separate model object)"""
    rng = np.random.default_rng(8)
    m = oracle.Model("WAG")
    parent, dist = random_tree(rng, 40)
    obs = _random_alignment(rng, parent, 30, 20, 0.0)
    got = oracle.indel_column_logl(oracle.GapModel(m, 0, 0), parent, dist, obs, 1.7, 0.05, nthreads=2)
    _, want = oracle.fast_marginal(m, parent, dist, obs, None, np.full(30, 1.7), want_post=False)
    assert np.allclose(got, want + np.log(0.05), rtol=1e-11)


def test_oracle_alignment_likelihood_and_brent(oracle):
    """calcProbAlnGivenTree pieces (IndelPeeler.java:125-193) and optimiseMuLambda's search: the optimum is interior
    for an alignment simulated with indels, and no sampled neighbour does better"""
    rng = np.random.default_rng(3)
    m = oracle.Model("JTT")
    parent, dist = random_tree(rng, 24)
    obs = _random_alignment(rng, parent, 60, 20, 0.0)
    n = len(parent)
    _clade_gaps(rng, parent, obs, range(60), per_col=2)
    geom = 1 / 40.0
    gm = oracle.GapModel(m, 0.1, 0.1)
    cols = oracle.indel_column_logl(gm, parent, dist, obs, 1.0, geom)
    gapcol = oracle.indel_column_logl(gm, parent, dist, np.full((n, 1), NONE, np.uint8), 1.0, geom)[0]
    extra = oracle.indel_prob_extra_col(gm, parent, dist, geom)
    total = oracle.indel_aln_logl(gm, parent, dist, obs, geom)
    assert abs(total - ((extra - oracle.logm1exp(gapcol)) + cols.sum())) < 1e-9 * abs(total)
    kids_logs = sum(np.log(1 - gm.ksi(dist[v])) for v in range(1, n))
    assert abs(extra - (np.log(1 - geom) + kids_logs)) < 1e-10
    assert abs(oracle.logm1exp(-0.3) - np.log(1 - np.exp(-0.3))) < 1e-14 and abs(oracle.logm1exp(-5.0) - np.log1p(-np.exp(-5.0))) < 1e-15
    assert abs(oracle.logsumexp([-1000.0, -1001.0, -np.inf]) - (-1000 + np.log(1 + np.exp(-1)))) < 1e-12
    x, n_evals = oracle.indel_optimise_mulambda(m, parent, dist, obs, geom, 1e-4, 2.0, nthreads=4)
    assert 1e-3 < x < 1.9 and 8 <= n_evals <= 120
    f = lambda q: -oracle.indel_aln_logl(oracle.GapModel(m, q, q), parent, dist, obs, geom)   # noqa: E731
    assert all(f(x) <= f(x * s) + 1e-7 for s in (0.5, 0.8, 0.95, 1.05, 1.25, 2.0))


# ---------------------------------------------------------------------------------------------------------------
# device
# ---------------------------------------------------------------------------------------------------------------
def _close(got, want, tol=TOL):
    got, want = np.asarray(got), np.asarray(want)
    inf = np.isinf(want)
    assert np.array_equal(np.isinf(got), inf) and np.array_equal(got[inf], want[inf])
    err = np.abs(got[~inf] - want[~inf]) / np.abs(want[~inf])
    assert err.size == 0 or err.max() <= tol, err.max()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["JTT", "LG", "JC", "Yang"])
def test_gpu_gap_model_tables(bnk, oracle, name):
    """bnk_indel_probs: P_eps(t), ksi, gamma, stationary frequencies against the oracle (same arithmetic, fdlibm exp,
    no FMA contraction: bit for bit)"""
    m, om = bnk.Model(name), oracle.Model(name)
    h = bnk.Indel(m, bnk.Tree([-1, 0, 0], [0.0, 0.1, 0.2]), 4)
    for mu, lam in ((0.05, 0.03), (0.0, 0.0), (0.4, 0.4)):
        h.set_rates(mu, lam)
        g = oracle.GapModel(om, mu, lam)
        for t in (1e-5, 0.02, 0.7, 11.0):
            P, ksi, gamma, F = h.probs(t)
            assert np.array_equal(P, g.probs(t))
            assert abs(ksi - g.ksi(t)) <= 1e-15 and abs(gamma - g.gamma(t)) <= 1e-15
            assert np.array_equal(F, g.parts()["F"])
    h.close()


@pytest.mark.gpu
def test_gpu_column_likelihoods_random_trees(bnk, oracle):
    """IndelPeeler.decorate per column: random trees incl. multifurcations, 0-95 % gaps, all-gap columns, 20- and
    4-state models, mu != lambda, no indels, several indel rates"""
    rng = np.random.default_rng(17)
    for trial in range(24):
        name = ["JTT", "LG", "WAG", "Dayhoff", "JC", "Yang"][trial % 6]
        m, om = bnk.Model(name), oracle.Model(name)
        parent, dist = random_tree(rng, int(rng.integers(2, 90)), max_children=int(rng.integers(2, 6)), zero_branch_prob=0.1)
        n_cols = int(rng.integers(1, 300))
        obs = _random_alignment(rng, parent, n_cols, m.K, float(rng.choice([0.0, 0.3, 0.6, 0.95])))
        obs[:, rng.integers(0, n_cols)] = NONE
        mu, lam = [(0.05, 0.03), (0.2, 0.2), (0.0, 0.0), (0.3, 0.001), (0.001, 0.4)][trial % 5]
        rate, geom = float(rng.choice([0.25, 1.0, 3.0])), float(rng.choice([0.5, 0.02, 1e-3]))
        h = bnk.Indel(m, bnk.Tree(parent, dist), n_cols)
        h.set_alignment(obs)
        h.set_rates(mu, lam)
        got, s = h.column_logl(rate, geom)
        want = oracle.indel_column_logl(oracle.GapModel(om, mu, lam), parent, dist, obs, rate, geom, nthreads=4)
        _close(got, want)
        if np.isfinite(want).all():
            assert abs(s - want.sum()) <= TOL * abs(want.sum())
        h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["caterpillar", "balanced", "star"])
def test_gpu_column_likelihoods_deep_and_wide_trees(bnk, oracle, shape):
    """a 700-level caterpillar (everything in the walk kernel; exponents reach -3000), a 2,048-leaf balanced tree
    (level launches + walk), a 300-way star (one node, many children)"""
    from bnkit_b200 import synth
    rng = np.random.default_rng(23)
    if shape == "caterpillar":
        parent, dist = synth.caterpillar_tree(700, seed=2)
    elif shape == "balanced":
        parent, dist = synth.balanced_tree(2048, seed=2)
    else:
        parent = np.array([-1] + [0] * 300, np.int32)
        dist = np.concatenate([[0.0], rng.gamma(1.0, 0.2, 300) + 1e-3])
    n_cols = 150
    m, om = bnk.Model("LG"), oracle.Model("LG")
    obs = _random_alignment(rng, parent, n_cols, 20, 0.3)
    _clade_gaps(rng, parent, obs, range(0, n_cols, 3))   # all-gap subtrees of every size
    h = bnk.Indel(m, bnk.Tree(parent, dist), n_cols)
    h.set_alignment(obs)
    h.set_rates(0.08, 0.05)
    got, _ = h.column_logl(1.0, 0.01)
    want = oracle.indel_column_logl(oracle.GapModel(om, 0.08, 0.05), parent, dist, obs, 1.0, 0.01, nthreads=8)
    _close(got, want)
    h.close()


@pytest.mark.gpu
def test_gpu_alignment_likelihood_priors_and_brent_on_reference_data(bnk, oracle):
    """data/3_2_1_1_filt.* (BASELINE configs[1]; 39 extants x 1,307 columns, 61 % gaps) under JTT:
    calcProbAlnGivenTree, computeColumnPriors over five rate categories, optimiseMuLambda -- the same optimum after
    the same number of likelihood evaluations as the oracle's restatement of Minimise.brent"""
    fx = gold("c2_3_2_1_1_filt.json")
    parent, rows, n_pos, obs = fixture_rows(fx)
    parent = np.asarray(parent, np.int32)
    dist = np.asarray(fx["tree"]["Distances"])
    m, om = bnk.Model("JTT"), oracle.Model("JTT")
    lengths = [(obs[v] != NONE).sum() for v in range(len(parent)) if leaves_of(parent)[v]]
    geom = 1.0 / np.mean(lengths)                    # Mip.java:301: 1 / aln.getAvgSeqLength()
    h = bnk.Indel(m, bnk.Tree(parent, dist), n_pos)
    h.set_alignment(obs)
    h.set_rates(0.05, 0.05)
    gm = oracle.GapModel(om, 0.05, 0.05)
    want = oracle.indel_aln_logl(gm, parent, dist, obs, geom, nthreads=8)
    assert abs(h.aln_logl(geom) - want) <= TOL * abs(want)
    rates = np.array([0.1, 0.5, 1.0, 2.0, 5.0])
    pri = h.column_priors(rates, geom)
    for r, rate in enumerate(rates):
        _close(pri[:, r], oracle.indel_column_logl(gm, parent, dist, obs, rate, geom, nthreads=8))
    x, n_evals = h.optimise_mulambda(1e-4, 1.0, geom)
    wx, wn = oracle.indel_optimise_mulambda(om, parent, dist, obs, geom, 1e-4, 1.0, nthreads=8)
    # Brent stops when its bracket is narrower than EPS = 1e-5 (Minimise.java:9): that is the agreement to ask for
    assert abs(x - wx) <= 1.5e-5 and abs(n_evals - wn) <= 3
    h.close()


@pytest.mark.gpu
def test_gpu_indel_argument_errors(bnk):
    m = bnk.Model("LG")
    tree = bnk.Tree([-1, 0, 0], [0.0, 0.1, 0.2])
    h = bnk.Indel(m, tree, 8)
    with pytest.raises(bnk.BnkError):
        h.column_logl()                                   # nothing uploaded
    obs = np.full((3, 4), NONE, np.uint8)
    obs[1] = 3
    h.set_alignment(obs)
    with pytest.raises(bnk.BnkError):
        h.column_logl()                                   # no rates
    with pytest.raises(bnk.BnkError):
        h.set_rates(-0.5, 0.1)
    bad = obs.copy()
    bad[2, 1] = 77
    with pytest.raises(bnk.BnkError):
        h.set_alignment(bad)
    with pytest.raises(bnk.BnkError):
        h.set_alignment(np.full((3, 9), NONE, np.uint8))  # wider than max_cols
    with pytest.raises(bnk.BnkError):
        bnk.Indel(m, bnk.Tree([-1, -1, 0], [0.0, 0.0, 0.2]), 8)   # a forest
    h.close()


@pytest.mark.gpu
def test_gpu_indel_java_api_mirror_and_base_models(bnk, oracle):
    """phylo.IndelPeeler.optimiseMuLambda / calcProbAlnGivenTree / computeColumnPriors (the reference's static entry
    points) on data/66.*; the base model optimiseMuLambda builds for "JC" and "Yang" is symmetric + normalised
    (IndelPeeler.java:474-503), unlike SubstModel.createModel's; an unknown name raises as the reference does"""
    from bnkit_b200.phylo import GapSubstModel, IdxTree, IndelPeeler, SubstModel
    fx = gold("c1_66.json")
    parent, rows, n_pos, obs = fixture_rows(fx)
    parent = np.asarray(parent, np.int32)
    dist = np.asarray(fx["tree"]["Distances"])
    tree = IdxTree.fromJSON(fx["tree"])
    states = [[None] * n_pos for _ in range(tree.getSize())]
    for nm, seq in zip(fx["names"], fx["seqs"]):
        states[tree.getIndex(nm)] = [None if ch == "-" else ch for ch in seq]
    geom = 0.4
    om = oracle.Model("JTT")
    x = IndelPeeler.optimiseMuLambda(1e-4, 2.0, "JTT", tree, geom, states)
    wx, _ = oracle.indel_optimise_mulambda(om, parent, dist, obs, geom, 1e-4, 2.0, nthreads=4)
    assert abs(x - wx) <= 1.5e-5          # Brent's own stopping width (EPS = 1e-5, Minimise.java:9)
    gm = GapSubstModel(SubstModel.createModel("JTT"), x, x)
    assert gm.getDomain()[-1] == "-" and len(gm.getDomain()) == 21
    ogm = oracle.GapModel(om, x, x)
    want = oracle.indel_aln_logl(ogm, parent, dist, obs, geom)
    assert abs(IndelPeeler.calcProbAlnGivenTree(tree, states, gm, geom) - want) <= TOL * abs(want)
    pri = IndelPeeler.computeColumnPriors(tree, states, gm, geom, [0.5, 1.0, 2.0])
    assert pri.shape == (n_pos, 3)
    _close(pri[:, 2], oracle.indel_column_logl(ogm, parent, dist, obs, 2.0, geom))
    with pytest.raises(ValueError):
        IndelPeeler.optimiseMuLambda(1e-4, 2.0, "Blosum", tree, geom, states)
    # JC / Yang: F and Q of the named matrix pushed through the symmetric, normalising constructor
    yang_F = [0.308, 0.185, 0.199, 0.308]
    yang_Q = [[0.000, 0.139, 0.566, 0.115], [0.232, 0.000, 0.223, 0.882], [0.877, 0.207, 0.000, 0.215], [0.115, 0.530, 0.139, 0.000]]
    for name, custom in (("JC", (4, [0.25] * 4, np.full((4, 4), 0.25) - 0.25 * np.eye(4), 1, 1)), ("Yang", (4, yang_F, yang_Q, 1, 1))):
        got = bnk.Model(name, indel_base=True).parts()
        ref = oracle.Model(custom=custom).parts()
        assert np.array_equal(got["R"], ref["R"]) and np.array_equal(got["F"], ref["F"])
        assert not np.array_equal(got["R"], bnk.Model(name).parts()["R"])


@pytest.mark.gpu
def test_gpu_indel_several_devices_behind_one_handle(bnk, oracle):
    """bnk_indel_create_multi: column blocks on several devices (here: three parts, on as many devices as the box
    has -- a device may serve several parts), results identical to the single-device handle column by column, sums in
    column order, the same Brent trajectory; more parts than columns is legal"""
    rng = np.random.default_rng(29)
    m = bnk.Model("JTT")
    parent, dist = random_tree(rng, 60, max_children=3)
    n_cols = 101
    obs = _random_alignment(rng, parent, n_cols, 20, 0.4)
    _clade_gaps(rng, parent, obs, range(0, n_cols, 2))
    tree = bnk.Tree(parent, dist)
    nd = max(bnk.device_count(), 1)
    one = bnk.Indel(m, tree, n_cols)
    multi = bnk.Indel(m, tree, n_cols, devices=[0 % nd, 1 % nd, 2 % nd])
    for h in (one, multi):
        h.set_alignment(obs)
        h.set_rates(0.06, 0.04)
    a, sa = one.column_logl(1.3, 0.02)
    b, sb = multi.column_logl(1.3, 0.02)
    assert np.array_equal(a, b) and sa == sb
    assert one.aln_logl(0.02) == multi.aln_logl(0.02)
    assert np.array_equal(one.column_priors([0.5, 2.0], 0.02), multi.column_priors([0.5, 2.0], 0.02))
    assert one.optimise_mulambda(1e-4, 1.0, 0.02) == multi.optimise_mulambda(1e-4, 1.0, 0.02)
    tiny = bnk.Indel(m, tree, 2, devices=[0] * 5)          # five parts, two columns
    tiny.set_alignment(obs[:, :2])
    tiny.set_rates(0.06, 0.04)
    one.set_alignment(obs[:, :2])
    one.set_rates(0.06, 0.04)                              # (the Brent search above left its last evaluation's rates)
    assert np.array_equal(tiny.column_logl(1.0, 0.02)[0], one.column_logl(1.0, 0.02)[0])
    for h in (one, multi, tiny):
        h.close()
