"""The CPU oracle against everything the reference itself pins for this path:
tutorial known answers, the brute-force method of MaxLhoodJointTest / MaxLhoodMarginalTest,
the IdxTreeTest prune fixture, and the documented Recon output on data/66.*."""
import json
import os

import numpy as np
import pytest

from helpers import NONE, leaves_of, random_obs, random_present, random_tree

GOLD = os.path.join(os.path.dirname(__file__), "golden")
AA = "ARNDCQEGHILKMFPSTWYV"


def gold(name):
    with open(os.path.join(GOLD, name)) as fh:
        return json.load(fh)


def test_elementary_functions_within_one_ulp(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-745, 709, 20000), rng.uniform(-1, 1, 20000), [0.0, -0.0, 1e-300, -1e-300]])
    e = np.array([L.om_exp(float(x)) for x in xs])
    assert (np.abs(e - np.exp(xs)) <= np.spacing(np.exp(xs))).all()
    ys = np.concatenate([rng.uniform(1e-12, 3, 20000), np.exp(rng.uniform(-700, 700, 20000)), [5e-324, 1.0]])
    l = np.array([L.om_log(float(y)) for y in ys])
    assert (np.abs(l - np.log(ys)) <= np.spacing(np.abs(np.log(ys)))).all()
    assert L.om_log(0.0) == -np.inf and np.isnan(L.om_log(-1.0))
    assert L.om_exp(-np.inf) == 0.0 and L.om_exp(np.inf) == np.inf
    # Factorize.logSumOfLogs (Factorize.java:1199-1210)
    assert L.om_logsum(-np.inf, -3.0) == -3.0 and L.om_logsum(-2.0, -np.inf) == -2.0
    assert L.om_logsum(-np.inf, -np.inf) == -np.inf
    assert np.isclose(L.om_logsum(np.log(0.25), np.log(0.5)), np.log(0.75), rtol=1e-15)


def test_tutorial_rate_matrix_and_transition_probabilities(oracle):
    fx = gold("tutorial_lg.json")
    m = oracle.Model("LG")
    p = m.parts()
    # the tutorial prints S as 0.062 where LG.java has 0.061197: allow one unit in the last printed place
    assert np.abs(p["F"] - np.array(fx["F_3dp"])).max() < 1.0e-3
    assert np.abs(np.round(p["R"], 2) - np.array(fx["R_2dp"])).max() < 1e-12
    for t, P in fx["P_3dp"].items():
        assert np.abs(m.probs(float(t)) - np.array(P)).max() <= 0.0005 + 1e-12
    assert np.abs(p["evec"] @ p["ievec"] - np.eye(20)).max() < 1e-13
    assert np.abs(p["R"].sum(1)).max() < 1e-14                       # makeValid
    assert np.isclose(-(np.diag(p["R"]) * p["F"]).sum(), 1.0)        # normalize: one substitution per unit time
    assert np.isclose(p["prior"].sum(), 1.0) and np.allclose(p["prior"] * p["F"].sum(), p["F"])


def test_tutorial_joint_example(oracle):
    """tutorial_ML.html:205-238 == SubstNode.main (SubstNode.java:891-957): N0 = E, N1 = E."""
    fx = gold("tutorial_lg.json")["example"]
    m = oracle.Model("LG")
    par, dist = fx["tree"]["parent"], fx["tree"]["dist"]
    obs = np.full(4, NONE, np.uint8)
    obs[1], obs[3] = AA.index("E"), AA.index("D")
    b = oracle.brute(m, par, dist, obs)
    st = oracle.fast_joint(m, par, dist, obs[:, None])[:, 0]
    assert [AA[s] for s in st] == ["E", "E", "E", "D"] and list(b["state"]) == list(st)
    prior = m.parts()["prior"]
    P007, P001, P01 = m.probs(0.07), m.probs(0.01), m.probs(0.1)
    def prod(a, c, Pmid):
        a, c = AA.index(a), AA.index(c)
        return prior[a] * P007[a, AA.index("E")] * Pmid[a, c] * P007[c, AA.index("D")]
    for k, v in fx["joint_products_8dp"].items():
        assert abs(prod(k[0], k[1], P001) - v) < 5.1e-9
    tot = sum(prod(a, c, P001) for a in AA for c in AA)
    for k, v in fx["joint_normalised_3dp"].items():
        assert abs(prod(k[0], k[1], P001) / tot - v) < 5.1e-4
    tot = sum(prod(a, c, P01) for a in AA for c in AA)
    for k, v in fx["joint_normalised_t01_3dp"].items():
        assert abs(prod(k[0], k[1], P01) / tot - v) < 5.1e-4
    assert np.isclose(b["total_p"], sum(prod(a, c, P001) for a in AA for c in AA), rtol=1e-12)


def test_joint_equals_bruteforce_jc(oracle):
    """MaxLhoodJointTest.joint1 (test/asr/MaxLhoodJointTest.java:281-365): JC, 3..7 leaves, random leaf states."""
    m = oracle.Model("JC")
    for n_leaves in range(3, 8):
        for seed in range(40):
            rng = np.random.default_rng(1000 * n_leaves + seed)
            parent, dist = random_tree(rng, n_leaves)
            obs = random_obs(rng, parent, 1, 4)
            b = oracle.brute(m, parent, dist, obs[:, 0], want_post=False)
            st = oracle.fast_joint(m, parent, dist, obs)[:, 0]
            # the brute-force maximum is unique up to ties; compare likelihood of the two assignments
            if not np.array_equal(st, b["state"]):
                bb = oracle.brute(m, parent, dist, st, want_post=False)  # everything observed: P(assignment)
                assert np.isclose(bb["total_p"], b["best_p"], rtol=1e-12)


@pytest.mark.parametrize("name", ["JTT", "LG", "WAG", "Dayhoff", "JC", "Yang"])
def test_conditional_tables_behave_as_substnodetest_demands(oracle, name):
    """What SubstNodeTest.getBarebone asserts of every SubstNode along a 200-node chain under JTT
    (test/bn/ctmc/SubstNodeTest.java:47-125): P(sym | sym) falls as the branch grows, every entry is a probability,
    every row sums to 1 within 1 % -- here for all six models (no row renormalisation: SubstModel.java:303-327)."""
    m = oracle.Model(name)
    rng = np.random.default_rng(200)
    ts = np.sort(np.concatenate([rng.uniform(1e-5, 0.05, 40), rng.uniform(0.05, 5.0, 40)]))
    prev = np.ones(m.K)
    for t in ts:
        P = m.probs(float(t))
        assert (P >= 0.0).all() and (P <= 1.0).all()
        assert (np.abs(P.sum(axis=1) - 1.0) < 0.01).all()
        d = np.diag(P)
        assert (d <= prev).all()
        prev = d
    prior = m.parts()["prior"]
    assert np.isclose(prior.sum(), 1.0, atol=1e-12) and (prior > 0).all()      # the root's table: F / sum F


def test_reference_known_answers_infer1_infer1b(oracle):
    """The states the reference's own tests assert, MaxLhoodJointTest.infer1 / infer1b
    (test/asr/MaxLhoodJointTest.java:42-46, 148-191).  Nodes in IdxTree (depth-first) order; A = 0, C = 1."""
    A, Cc = 0, 1
    # infer1: ((Leaf1:0.02,Leaf2:0.12)Anc_left:0.19,(Leaf3:0.18,Leaf4:0.15)Anc_right:0.191)Root under JC, Leaf1 = A, Leaf3 = C
    parent = np.array([-1, 0, 1, 1, 0, 4, 4], dtype=np.int32)
    dist = np.array([0.0, 0.19, 0.02, 0.12, 0.191, 0.18, 0.15])
    obs = np.array([NONE, NONE, A, NONE, NONE, Cc, NONE], dtype=np.uint8)
    st = oracle.fast_joint(oracle.Model("JC"), parent, dist, obs[:, None])[:, 0]
    assert (st[2], st[1], st[4], st[0]) == (A, A, Cc, A)   # Leaf1, ancestor 1 (Anc_left), ancestor 2 (Anc_right), root
    # infer1b: JC over the two-letter alphabet {A, C} (JC.java:44-60: q = gamma / n off the diagonal, uniform F);
    # mini1 = ((Leaf1:0.09,Leaf2:0.11)N1:0.07,Leaf3:0.12)N0, mini2 = ((Leaf1:0.03,Leaf2:0.05)N1:0.13,Leaf3:0.12)N0,
    # Leaf1 = A, Leaf3 = C: the root follows Leaf1 in mini1 and Leaf3 in mini2, N1 is A in both
    Q = np.array([[0.0, 0.5], [0.5, 0.0]])
    jc2 = oracle.Model(custom=(2, [0.5, 0.5], Q, False, False))
    par = np.array([-1, 0, 1, 1, 0], dtype=np.int32)
    ob = np.array([NONE, NONE, A, NONE, Cc], dtype=np.uint8)
    s1 = oracle.fast_joint(jc2, par, np.array([0.0, 0.07, 0.09, 0.11, 0.12]), ob[:, None])[:, 0]
    s2 = oracle.fast_joint(jc2, par, np.array([0.0, 0.13, 0.03, 0.05, 0.12]), ob[:, None])[:, 0]
    assert (s1[0], s1[1]) == (A, A) and (s2[0], s2[1]) == (Cc, A)
    # the brute-force enumerator agrees on all of it
    for d_ in ([0.0, 0.07, 0.09, 0.11, 0.12], [0.0, 0.13, 0.03, 0.05, 0.12]):
        b = oracle.brute(jc2, par, np.array(d_), ob, want_post=False)
        assert np.array_equal(b["state"], oracle.fast_joint(jc2, par, np.array(d_), ob[:, None])[:, 0])


def test_joint_with_internal_evidence_wag(oracle):
    """MaxLhoodJointTest.infer0 (:55-147): 7-node tree, 3 randomly instantiated nodes incl. internal ones."""
    m = oracle.Model("WAG")
    parent = np.array([-1, 0, 1, 1, 0, 4, 4], dtype=np.int32)
    for seed in range(60):
        rng = np.random.default_rng(seed)
        dist = np.maximum(rng.gamma(1.1, 0.2, 7), 1e-7)
        obs = np.full(7, NONE, np.uint8)
        for v in rng.choice(7, 3, replace=False):
            obs[v] = rng.integers(0, 20)
        b = oracle.brute(m, parent, dist, obs, want_post=False)
        st = oracle.fast_joint(m, parent, dist, obs[:, None])[:, 0]
        if not np.array_equal(st, b["state"]):
            bb = oracle.brute(m, parent, dist, st, want_post=False)
            assert np.isclose(bb["total_p"], b["best_p"], rtol=1e-12)


def test_marginal_equals_bruteforce(oracle):
    """MaxLhoodMarginalTest.marginal0 / marginal1 (:205-310, :330-399) with 1e-9 instead of 0.01."""
    for name, K, sizes in (("LG", 20, (3, 4)), ("JC", 4, (3, 5, 8))):
        m = oracle.Model(name)
        for n_leaves in sizes:
            for seed in range(25):
                rng = np.random.default_rng(7000 + 100 * n_leaves + seed)
                parent, dist = random_tree(rng, n_leaves, max_children=3 if K == 4 else 2)
                obs = random_obs(rng, parent, 1, K, internal_obs_prob=0.2 if K == 4 else 0.0)
                b = oracle.brute(m, parent, dist, obs[:, 0])
                post, logl = oracle.fast_marginal(m, parent, dist, obs)
                free = obs[:, 0] == NONE
                assert np.allclose(post[free, 0, :], b["post"][free], rtol=1e-9, atol=1e-15)
                assert np.isclose(logl[0], np.log(b["total_p"]), rtol=1e-10)


def test_masks_and_forests_against_bruteforce(oracle):
    m = oracle.Model("JC")
    for seed in range(60):
        rng = np.random.default_rng(9000 + seed)
        parent, dist = random_tree(rng, int(rng.integers(3, 7)), max_children=3)
        obs = random_obs(rng, parent, 1, 4, gap_prob=0.3)
        present = random_present(rng, parent, obs, absent_prob=0.3)
        b = oracle.brute(m, parent, dist, obs[:, 0], present[:, 0])
        st = oracle.fast_joint(m, parent, dist, obs, present)[:, 0]
        post, logl = oracle.fast_marginal(m, parent, dist, obs, present)
        ok = ~np.isnan(post[:, 0, 0])
        assert np.allclose(post[ok, 0, :], b["post"][ok], rtol=1e-9, atol=1e-15)
        assert np.isclose(logl[0], np.log(b["total_p"]), rtol=1e-10)
        if not np.array_equal(st, b["state"]):
            o2 = st.copy()
            bb = oracle.brute(m, parent, dist, o2, present[:, 0], want_post=False)
            assert np.isclose(bb["total_p"], b["best_p"], rtol=1e-12)
        assert (st[present[:, 0] == 0] == NONE).all()


def test_prune_fixture(oracle):
    """IdxTreeTest.createPrunedTree (test/dat/phylo/IdxTreeTest.java:15-56)."""
    fx = gold("idxtree_prune.json")
    parent = np.array(fx["parents"], dtype=np.int32)
    n = len(parent)
    for case in fx["cases"]:
        present = np.ones(n, np.uint8)
        present[case["prune"]] = 0
        out, cnt, roots = oracle.prune(parent, present, remove_orphans=False)
        assert cnt == n - case["removed"] and roots == case["roots"]
        if case["same_size_with_orphan_removal"]:
            out2, cnt2, _ = oracle.prune(parent, present, remove_orphans=True)
            assert cnt2 == cnt


def test_orphan_removal_semantics(oracle):
    # chain 0 - 1 - 2(leaf), 1 - 3(leaf): pruning both leaves orphans the ancestors
    parent = np.array([-1, 0, 1, 1], dtype=np.int32)
    out, cnt, roots = oracle.prune(parent, np.array([1, 1, 0, 0], np.uint8), True)
    assert cnt == 0
    out, cnt, roots = oracle.prune(parent, np.array([1, 0, 1, 1], np.uint8), True)
    assert list(out) == [0, 0, 1, 1] and roots == 2   # root has no extant descendant through present nodes


def test_documented_recon_output_c1(oracle):
    """docs/json-api.md:97-122: data/66.*, JTT, joint => N0, N1, N2 present at columns {2, 4} with states A, H."""
    from bnkit_b200 import synth
    c1 = gold("c1_66.json")
    parent = np.array(c1["tree"]["Parents"], dtype=np.int32)
    assert c1["tree"]["Parents"] == c1["doc_parents"]          # our Newick indexing == IdxTree's
    dist = np.array(c1["tree"]["Distances"])
    labels = c1["tree"]["Labels"]
    n, C = len(parent), len(c1["seqs"][0])
    obs = np.full((n, C), NONE, np.uint8)
    for name, seq in zip(c1["names"], c1["seqs"]):
        v = labels.index(name)
        for c, ch in enumerate(seq):
            if ch != "-":
                obs[v, c] = AA.index(ch)
    leaf = leaves_of(parent)
    present = synth.path_presence_mask(parent, (obs != NONE) & leaf[:, None])
    m = oracle.Model("JTT")
    st = oracle.fast_joint(m, parent, dist, obs, present)
    doc = c1["doc_result"]
    for anc in (0, 1, 2):
        cols = [c for c in range(C) if present[anc, c]]
        assert cols == doc["indices"]
        assert [AA[st[anc, c]] for c in cols] == doc["values"]


def test_generic_bucket_elimination_agrees_with_tree_sweeps(oracle):
    """bnk_ve.c (factor tables + buckets + product trees, no tree knowledge) vs bnk_sweeps.c: identical
    joint states -- including on multifurcations, where the association of the additions follows
    Factorize.getProductTree -- and posteriors / log-likelihoods equal to rounding."""
    cases = 0
    for name, K in (("JC", 4), ("LG", 20)):
        m = oracle.Model(name)
        for seed in range(30 if K == 4 else 10):
            rng = np.random.default_rng(4000 + seed)
            parent, dist = random_tree(rng, int(rng.integers(3, 14)), max_children=int(rng.integers(2, 6)))
            obs = random_obs(rng, parent, 3, K, gap_prob=0.2, internal_obs_prob=0.15 if seed % 2 else 0.0)
            present = random_present(rng, parent, obs, absent_prob=0.25) if seed % 3 == 0 else None
            rate = float(rng.choice([0.3, 1.0, 2.2]))
            st = oracle.fast_joint(m, parent, dist, obs, present, np.full(3, rate))
            post, logl = oracle.fast_marginal(m, parent, dist, obs, present, np.full(3, rate))
            for c in range(3):
                pc = None if present is None else np.ascontiguousarray(present[:, c])
                oc = np.ascontiguousarray(obs[:, c])
                vst, _ = oracle.ve_joint(m, parent, dist, oc, pc, rate)
                assert np.array_equal(vst, st[:, c]), (name, seed, c)
                # single tree: VarElim.logLikelihood; forests: all components (see bnk_ve.c)
                assert np.isclose(oracle.ve_loglik(m, parent, dist, oc, pc, rate), logl[c], rtol=1e-12, atol=1e-12)
                for v in range(len(parent)):
                    vp = oracle.ve_marginal(m, parent, dist, oc, v, pc, rate)
                    if vp is None:
                        assert np.isnan(post[v, c]).all()
                    else:
                        assert np.allclose(vp, post[v, c], rtol=1e-10, atol=1e-14), (name, seed, c, v)
                        cases += 1
    assert cases > 300


def test_association_order_is_pinned_bitwise(oracle):
    """On single trees with leaf evidence only, the maximal log-probability the tree sweep reports must
    equal the generic bucket elimination's BIT FOR BIT: that is only true if every log-space addition
    happens in the same association (Factorize.getProductTree order) -- checked on multifurcations."""
    m = oracle.Model("LG")
    checked = 0
    for seed in range(40):
        rng = np.random.default_rng(6000 + seed)
        parent, dist = random_tree(rng, int(rng.integers(4, 30)), max_children=int(rng.integers(3, 7)))
        obs = random_obs(rng, parent, 2, 20)
        st, lm = oracle.fast_joint(m, parent, dist, obs, None, None, want_logmax=True)
        for c in range(2):
            vst, vlm = oracle.ve_joint(m, parent, dist, np.ascontiguousarray(obs[:, c]))
            assert np.array_equal(vst, st[:, c])
            assert vlm == lm[c], (seed, c, vlm, lm[c])
            checked += 1
    assert checked == 80


@pytest.mark.parametrize("name", ["LG", "JTT", "WAG", "Dayhoff", "JC", "Yang"])
def test_transition_probabilities_against_an_independent_matrix_exponential(oracle, bnk, name):
    """The eigensolver of the oracle (oracle/bnk_oracle.c) and of the product's host set-up
    (bnkit_b200/csrc/subst_eigen.inc) are the same EISPACK transliteration of Matrix.java, so comparing them with
    each other pins nothing.  Independent check: P(t) = expm(R t) with scipy's Pade
    scaling-and-squaring, for the rate matrix R both report, to 1e-12 absolute over four decades of t; and the
    rate matrix itself against its definition (S.pi off the diagonal, zero row sums, one expected substitution
    per unit time -- SubstModel.java:80-110, 192-219)."""
    from scipy.linalg import expm
    om = oracle.Model(name)
    hm = bnk.Model(name)                     # host-side model set-up of the library (no GPU involved)
    for parts in (om.parts(), hm.parts()):
        R, F = parts["R"], parts["F"]
        K = len(F)
        assert np.allclose(R.sum(1), 0.0, atol=1e-13)
        if name != "JC":   # JC.java builds its matrix un-normalised (rate 0.25 per pair: 0.75 substitutions per unit time)
            assert np.isclose(-(np.diag(R) * F).sum(), 1.0, rtol=1e-12)
        # detailed balance of a reversible model: pi_i R_ij = pi_j R_ji
        if name in ("LG", "JTT", "WAG", "Dayhoff"):   # built from a symmetric S (Yang is given as a full Q)
            piR = (F / F.sum())[:, None] * R
            assert np.allclose(piR, piR.T, atol=1e-12)
        ev, iev, lam = parts["evec"], parts["ievec"], parts["eval"]
        assert np.allclose(ev @ iev, np.eye(K), atol=1e-10)
        for t in (1e-5, 0.003, 0.01, 0.07, 0.22, 1.0, 3.7, 25.0):
            want = expm(R * t)
            got = np.abs((ev * np.exp(lam * t)[None, :]) @ iev)
            assert np.abs(got - want).max() < 1e-12, (name, t)
            assert np.abs(om.probs(t) - want).max() < 1e-12, (name, t)
