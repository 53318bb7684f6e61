"""GPU parity: libbnkgpu (through its C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): joint states bit-exact, ties in the reference's order;
posteriors and log-likelihoods within 1e-9 relative.
"""
import numpy as np
import pytest

from helpers import NONE, leaves_of, random_obs, random_present, random_tree

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _post_close(got, want, tol=TOL):
    """posteriors: same NaN pattern and |got - want| <= tol * |want| -- RELATIVE, as BASELINE.json's north_star
    states it, for every entry down to 1e-300 (below that both sides must be below 1e-300 * (1 + tol))"""
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    g, w = got[ok], want[ok]
    big = np.abs(w) >= 1e-300
    if big.any():
        rel = np.abs(g[big] - w[big]) / np.abs(w[big])
        assert rel.max() <= tol, "max relative error %g (posterior %g)" % (rel.max(), w[big][rel.argmax()])
    assert (np.abs(g[~big]) <= 1e-300 * (1 + tol)).all()


@pytest.mark.parametrize("name", ["LG", "JTT", "WAG", "Dayhoff", "JC", "Yang"])
def test_tables_bit_exact(bnk, oracle, name):
    """Device P(t) and log P(t) equal the oracle's bit for bit (same eigensystem, same fdlibm exp/log)."""
    m = bnk.Model(name)
    om = oracle.Model(name)
    tree = bnk.Tree([-1, 0, 0], [0.0, 0.1, 0.2])
    ws = bnk.Workspace(m, tree, 4)
    rng = np.random.default_rng(7)
    times = np.concatenate([[0.0, 1e-7, 1e-5, 0.01, 0.07, 0.1, 1.0, 5.0, 50.0], rng.gamma(1.1, 0.2, 40)])
    P, L = ws.probs(times)
    for i, t in enumerate(times):
        Po = om.probs(t)
        assert np.array_equal(P[i], Po), (name, t)
        Lo = np.array([[oracle.lib().om_log(x) for x in row] for row in Po])
        assert np.array_equal(L[i], Lo), (name, t)


def _run_case(bnk, oracle, name, rng, n_leaves, n_cols, max_children=2, gap=0.0, internal=0.0, mask=False,
              rates=None, few_states=None, tile_cols=0, cpt=0):
    m = bnk.Model(name)
    om = oracle.Model(name)
    parent, dist = random_tree(rng, n_leaves, max_children=max_children, zero_branch_prob=0.05)
    obs = random_obs(rng, parent, n_cols, m.K, gap_prob=gap, internal_obs_prob=internal, few_states=few_states)
    present = random_present(rng, parent, obs) if mask else None
    tree = bnk.Tree(parent, dist)
    ws = bnk.Workspace(m, tree, n_cols)
    if tile_cols or cpt:
        ws.configure(tile_cols, cpt)
    ws.set_rates(n_cols, rates)
    got = ws.joint(obs, present)
    want = oracle.fast_joint(om, parent, dist, obs, present, rates, nthreads=4)
    assert np.array_equal(got, want), "joint states differ at %d of %d" % ((got != want).sum(), got.size)
    post, logl = ws.marginal(obs, present)
    wpost, wlogl = oracle.fast_marginal(om, parent, dist, obs, present, rates, nthreads=4)
    internal_nodes = tree.internal_nodes()
    _post_close(post, wpost[internal_nodes])
    assert np.allclose(logl, wlogl, rtol=TOL, atol=1e-12)
    logl2, total = ws.loglik(obs, present)
    assert np.allclose(logl2, wlogl, rtol=TOL, atol=1e-12)
    assert np.isclose(total, wlogl.sum(), rtol=TOL)
    ws.close()


@pytest.mark.parametrize("name", ["LG", "JTT", "WAG", "Dayhoff"])
def test_fast_path_random_trees(bnk, oracle, name):
    rng = np.random.default_rng(11)
    for n_leaves in (2, 3, 7, 40, 200):
        _run_case(bnk, oracle, name, rng, n_leaves, 67)


def test_four_state_models(bnk, oracle):
    rng = np.random.default_rng(12)
    for name in ("JC", "Yang"):
        for n_leaves in (3, 8, 50):
            _run_case(bnk, oracle, name, rng, n_leaves, 33)


def test_multifurcations_and_rates(bnk, oracle):
    rng = np.random.default_rng(13)
    rates = rng.choice([0.2, 0.7, 1.0, 3.1], size=301)
    _run_case(bnk, oracle, "LG", rng, 60, 301, max_children=5, rates=rates)
    rates = rng.gamma(0.5, 2.0, size=45) + 1e-3  # every column its own rate
    _run_case(bnk, oracle, "WAG", rng, 25, 45, max_children=3, rates=rates)


def test_general_path_masks_and_internal_evidence(bnk, oracle):
    rng = np.random.default_rng(14)
    _run_case(bnk, oracle, "JTT", rng, 39, 130, max_children=3, gap=0.6, mask=True)
    _run_case(bnk, oracle, "LG", rng, 30, 90, gap=0.3, internal=0.2, mask=True, rates=rng.choice([0.5, 2.0], size=90))
    _run_case(bnk, oracle, "LG", rng, 12, 64, internal=0.3)             # evidence on internal nodes, no mask
    _run_case(bnk, oracle, "LG", rng, 20, 64, gap=0.4)                  # uninstantiated leaves, everything present


def test_ties_go_to_lowest_state(bnk, oracle):
    """JC with equal branch lengths and few distinct leaf states produces exact ties."""
    rng = np.random.default_rng(15)
    m, om = bnk.Model("JC"), oracle.Model("JC")
    for n_leaves in (4, 9, 33):
        parent, dist = random_tree(rng, n_leaves)
        dist[:] = 0.25
        obs = random_obs(rng, parent, 128, 4)
        ws = bnk.Workspace(m, bnk.Tree(parent, dist), 128)
        got = ws.joint(obs)
        want = oracle.fast_joint(om, parent, dist, obs)
        assert np.array_equal(got, want)
        ws.close()


@pytest.mark.parametrize("tile_cols,cpt", [(4, 1), (7, 1), (16, 2), (38, 1), (30, 4), (64, 2), (17, -1), (20, -2), (-12, 2), (-30, 1)])
def test_tile_geometries(bnk, oracle, tile_cols, cpt, monkeypatch):
    """cpt < 0 forces the general kernels on inputs the lean kernels would otherwise take;
    tile_cols < 0 keeps the whole tree in the column-owned walker (no wide level launches).
    The geometry applies to the CTA-tiled walkers, so the warp-tiled one is switched off here."""
    monkeypatch.setenv("BNK_WALKER", "lean")
    rng = np.random.default_rng(16)
    rates = rng.choice([0.3, 1.0, 2.5], size=257)
    _run_case(bnk, oracle, "LG", rng, 50, 257, rates=rates, tile_cols=tile_cols, cpt=cpt)
    _run_case(bnk, oracle, "LG", rng, 300, 150, rates=rates[:150], tile_cols=tile_cols, cpt=cpt)   # has wide levels
    _run_case(bnk, oracle, "LG", rng, 21, 100, gap=0.5, mask=True, tile_cols=tile_cols, cpt=cpt)


@pytest.mark.parametrize("env", [{}, {"BNK_WARP_W": "2", "BNK_WARP_NS": "2", "BNK_WARP_C": "1"},
                                 {"BNK_WARP_W": "3", "BNK_WARP_NS": "5", "BNK_WARP_C": "3"},
                                 {"BNK_WARP_W": "4", "BNK_WARP_NS": "8", "BNK_WARP_C": "2"},
                                 {"BNK_WARP_W": "5", "BNK_WARP_NS": "3", "BNK_WARP_C": "1"},
                                 {"BNK_WARP_W": "1", "BNK_WARP_NS": "4", "BNK_WARP_C": "2"},
                                 {"BNK_WARP_W": "7", "BNK_WARP_NS": "6", "BNK_WARP_C": "1"}, {"BNK_WALKER": "lean"},
                                 {"BNK_WALKER": "scalar"}, {"BNK_MMA_W": "1", "BNK_WARP_NS": "2"}, {"BNK_MMA_W": "3"},
                                 {"BNK_MMA_W": "7", "BNK_WARP_NS": "5"}, {"BNK_TRACEBACK": "warp"}])
def test_warp_tiled_walker(bnk, oracle, env, monkeypatch):
    """sweeps_warp.cu (warp-owned columns, TMA-fed stage ring; sum-product sweeps on DMMA by default, scalar with
    BNK_WALKER=scalar) against the oracle for every CTA shape and ring depth, and against the lean CTA walker: deep chains (> 4 program chunks), category tiles narrower than a
    warp, uninstantiated childless nodes without a mask, hybrid trees (children handed over by the level
    kernels), whole trees in the walker, forests."""
    from bnkit_b200 import synth
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(31)
    m, om = bnk.Model("LG"), oracle.Model("LG")
    rates5 = rng.choice([0.2, 0.7, 1.0, 1.9, 4.0], size=77, p=[0.05, 0.3, 0.3, 0.3, 0.05])
    cases = []
    parent, dist = synth.caterpillar_tree(333, seed=5)
    cases.append((parent, dist, random_obs(rng, parent, 77, 20, few_states=5), rates5, 0))
    cases.append((parent, dist, random_obs(rng, parent, 77, 20, gap_prob=0.2), None, 0))
    parent, dist = random_tree(rng, 300, zero_branch_prob=0.05)
    cases.append((parent, dist, random_obs(rng, parent, 77, 20, gap_prob=0.05), rates5, 0))    # hybrid
    cases.append((parent, dist, random_obs(rng, parent, 77, 20), rates5, -8))                   # walker only
    pa, da = random_tree(rng, 40)
    pb, db = synth.caterpillar_tree(25, seed=6)
    parent = np.concatenate([pa, np.where(pb < 0, -1, pb + len(pa))]).astype(np.int32)          # two roots
    dist = np.concatenate([da, db])
    cases.append((parent, dist, random_obs(rng, parent, 13, 20, gap_prob=0.1), None, 0))
    for parent, dist, obs, rates, tc in cases:
        n_cols = obs.shape[1]
        tree = bnk.Tree(parent, dist)
        ws = bnk.Workspace(m, tree, n_cols)
        if tc:
            ws.configure(tc, 0)
        ws.set_rates(n_cols, rates)
        got = ws.joint(obs)
        want = oracle.fast_joint(om, parent, dist, obs, None, rates, nthreads=4)
        assert np.array_equal(got, want), "joint states differ at %d of %d" % ((got != want).sum(), got.size)
        post, logl = ws.marginal(obs)
        wpost, wlogl = oracle.fast_marginal(om, parent, dist, obs, None, rates, nthreads=4)
        _post_close(post, wpost[tree.internal_nodes()])
        assert np.allclose(logl, wlogl, rtol=TOL, atol=1e-12)
        l2, tot = ws.loglik(obs)
        assert np.allclose(l2, wlogl, rtol=TOL, atol=1e-12) and np.isclose(tot, wlogl.sum(), rtol=TOL)
        q = tree.internal_nodes()[::7]
        p2, _ = ws.marginal(obs, query=q, want_logl=False)
        assert np.array_equal(np.isnan(p2), np.isnan(post[::7])) and np.allclose(p2, post[::7], rtol=1e-12, atol=0, equal_nan=True)
        ws.close()


def test_deep_caterpillar_and_spill(bnk, oracle):
    """A 600-level caterpillar (one launch, no per-level cost) and a wide tree."""
    from bnkit_b200 import synth
    rng = np.random.default_rng(17)
    m, om = bnk.Model("LG"), oracle.Model("LG")
    for parent, dist in (synth.caterpillar_tree(600, seed=3), synth.balanced_tree(512, seed=4)):
        obs = random_obs(rng, parent, 96, 20, few_states=6)
        ws = bnk.Workspace(m, bnk.Tree(parent, dist), 96)
        got = ws.joint(obs)
        assert np.array_equal(got, oracle.fast_joint(om, parent, dist, obs, nthreads=4))
        post, logl = ws.marginal(obs)
        wpost, wlogl = oracle.fast_marginal(om, parent, dist, obs, nthreads=4)
        _post_close(post, wpost[leaves_of(parent) == False])
        assert np.allclose(logl, wlogl, rtol=TOL)
        ws.close()


def test_large_irregular_tree_hybrid(bnk, oracle):
    """5,000-leaf random (agglomeration) tree: many wide height levels plus a long narrow tail for the
    walker; joint bit-exact, posteriors / logL 1e-9, with Gamma categories."""
    from bnkit_b200 import synth
    rng = np.random.default_rng(18)
    parent, dist = synth.random_tree(5000, seed=9)
    m, om = bnk.Model("WAG"), oracle.Model("WAG")
    n_cols = 96
    rates = rng.choice(synth.gamma_category_rates(0.5, 4), size=n_cols)
    obs = synth.simulate_alignment(parent, dist, m.parts(), n_cols, rates, seed=10)
    tree = bnk.Tree(parent, dist)
    ws = bnk.Workspace(m, tree, n_cols)
    ws.set_rates(n_cols, rates)
    assert np.array_equal(ws.joint(obs), oracle.fast_joint(om, parent, dist, obs, None, rates, nthreads=8))
    post, logl = ws.marginal(obs)
    wpost, wlogl = oracle.fast_marginal(om, parent, dist, obs, None, rates, nthreads=8)
    _post_close(post, wpost[tree.internal_nodes()])
    assert np.allclose(logl, wlogl, rtol=TOL)
    l2, tot = ws.loglik(obs)
    assert np.allclose(l2, wlogl, rtol=TOL) and np.isclose(tot, wlogl.sum(), rtol=TOL)
    q = tree.internal_nodes()[::311]
    p2, _ = ws.marginal(obs, query=q, want_logl=False)
    assert np.allclose(p2, post[::311], rtol=1e-12, atol=0)
    ws.close()


def test_large_tree_with_masks_and_evidence_hybrid(bnk, oracle):
    """Per-column forests on a 3,000-leaf tree (wide levels + walker, general kernels): gaps, presence
    masks from the stand-in indel rule, orphan pruning, a sprinkle of evidence on internal nodes."""
    from bnkit_b200 import synth
    rng = np.random.default_rng(23)
    m, om = bnk.Model("JTT"), oracle.Model("JTT")
    n_cols = 80
    makers = ((synth.balanced_tree, 11, {}), (synth.random_tree, 12, {}),
              (synth.random_tree, 13, {"max_children": 4}))   # polytomies: wide levels below them, walker above
    for maker, seed, kw in makers:
        parent, dist = maker(3000, seed=seed, **kw)
        rates = rng.choice([0.4, 1.0, 2.5], size=n_cols)
        obs = synth.simulate_alignment(parent, dist, m.parts(), n_cols, rates, seed=seed, gap_fraction=0.55)
        leaf = leaves_of(parent)
        present = synth.path_presence_mask(parent, (obs != NONE) & leaf[:, None])
        # knock out random ancestors too (forests with many components), then prune orphans like GRASP does
        present[(rng.random(present.shape) < 0.05) & ~leaf[:, None]] = 0
        tree = bnk.Tree(parent, dist)
        present = tree.prune_orphans(present)
        ev = (rng.random(obs.shape) < 0.01) & ~leaf[:, None] & (present == 1)
        obs[ev] = rng.integers(0, 20, int(ev.sum()))
        ws = bnk.Workspace(m, tree, n_cols)
        ws.set_rates(n_cols, rates)
        got = ws.joint(obs, present)
        want = oracle.fast_joint(om, parent, dist, obs, present, rates, nthreads=8)
        assert np.array_equal(got, want), (got != want).sum()
        post, logl = ws.marginal(obs, present)
        wpost, wlogl = oracle.fast_marginal(om, parent, dist, obs, present, rates, nthreads=8)
        _post_close(post, wpost[tree.internal_nodes()])
        assert np.allclose(logl, wlogl, rtol=TOL, atol=1e-9)
        l2, tot = ws.loglik(obs, present)
        assert np.allclose(l2, wlogl, rtol=TOL, atol=1e-9)
        ws.close()
        # the same through the walker only
        ws2 = bnk.Workspace(m, tree, n_cols)
        ws2.configure(-16, 1)
        ws2.set_rates(n_cols, rates)
        assert np.array_equal(ws2.joint(obs, present), want)
        p2, l3 = ws2.marginal(obs, present)
        _post_close(p2, wpost[tree.internal_nodes()])
        assert np.allclose(l3, wlogl, rtol=TOL, atol=1e-9)
        ws2.close()


def test_rate_category_chunks(bnk, oracle, monkeypatch):
    """Every column its own rate => as many categories as columns; a tiny table budget forces the
    categories through several chunks (tables rebuilt per chunk, per-chunk traceback column ranges)."""
    monkeypatch.setenv("BNK_TABLE_BUDGET_MB", "1")
    rng = np.random.default_rng(29)
    for name, leaves, mc, gap in (("LG", 40, 2, 0.0), ("WAG", 30, 3, 0.4), ("LG", 220, 2, 0.0), ("JTT", 220, 2, 0.3)):
        m, om = bnk.Model(name), oracle.Model(name)
        parent, dist = random_tree(rng, leaves, max_children=mc)
        n_cols = 57
        rates = rng.gamma(0.7, 1.5, size=n_cols) + 0.01
        obs = random_obs(rng, parent, n_cols, 20, gap_prob=gap)
        present = random_present(rng, parent, obs) if gap else None
        tree = bnk.Tree(parent, dist)
        ws = bnk.Workspace(m, tree, n_cols)
        ws.set_rates(n_cols, rates)
        assert np.array_equal(ws.joint(obs, present), oracle.fast_joint(om, parent, dist, obs, present, rates))
        post, logl = ws.marginal(obs, present)
        wpost, wlogl = oracle.fast_marginal(om, parent, dist, obs, present, rates)
        _post_close(post, wpost[tree.internal_nodes()])
        assert np.allclose(logl, wlogl, rtol=TOL, atol=1e-10)
        ws.close()


def test_cherry_pattern_compression(bnk, oracle, monkeypatch):
    """The height-1 level with site-pattern compression (cherry.cu): one evaluation per distinct pair of
    child states and category.  Against the oracle, and bit-identical to the plain level kernel
    (BNK_NO_CHERRY=1), incl. uninstantiated leaves, unary height-1 nodes, 3 categories with ragged sizes,
    a group boundary (> 2048 columns in one category) and a 4-state model."""
    from bnkit_b200 import synth
    rng = np.random.default_rng(31)
    for name, K, leaves, n_cols, gap in (("LG", 20, 700, 2600, 0.0), ("JTT", 20, 400, 900, 0.03), ("JC", 4, 300, 700, 0.02)):
        m, om = bnk.Model(name), oracle.Model(name)
        parent, dist = synth.random_tree(leaves, seed=leaves)
        rates = rng.choice([0.3, 1.0, 2.2], size=n_cols, p=[0.85, 0.1, 0.05]) if n_cols > 2048 else rng.choice([0.3, 1.0, 2.2], size=n_cols)
        obs = synth.simulate_alignment(parent, dist, m.parts(), n_cols, rates, seed=leaves + 1, gap_fraction=gap)
        tree = bnk.Tree(parent, dist)
        sub = slice(0, n_cols, 9)  # the oracle on a sample of the columns
        want = oracle.fast_joint(om, parent, dist, np.ascontiguousarray(obs[:, sub]), None, rates[sub], nthreads=8)
        wpost, wlogl = oracle.fast_marginal(om, parent, dist, np.ascontiguousarray(obs[:, sub]), None, rates[sub], nthreads=8)
        res = []
        # compression with the records read in place by the parent level ("cherry-direct", the default), compression
        # with every message row expanded, no compression
        for no_cherry, direct in (("0", "1"), ("0", "0"), ("1", "1")):
            monkeypatch.setenv("BNK_NO_CHERRY", no_cherry)
            monkeypatch.setenv("BNK_CHERRY_DIRECT", direct)
            ws = bnk.Workspace(m, tree, n_cols)
            ws.set_rates(n_cols, rates)
            st = ws.joint(obs, None)
            post, logl = ws.marginal(obs, None)
            ws.close()
            assert np.array_equal(st[:, sub], want)
            _post_close(post[:, sub], wpost[tree.internal_nodes()])
            assert np.allclose(logl[sub], wlogl, rtol=TOL, atol=1e-9)
            res.append((st, post, logl))
        for other in res[1:]:
            assert np.array_equal(res[0][0], other[0])
            assert np.array_equal(res[0][1], other[1], equal_nan=True)
            assert np.array_equal(res[0][2], other[2])


def test_fused_height2_down_sweep(bnk, oracle, monkeypatch):
    """The down-sweep of the height-2 level evaluates its height-1 children in the same thread (wide.cu:
    down_wide2_kernel; their from-above vectors never reach HBM).  Against the oracle, and within a few ulp of the
    level-by-level kernels (BNK_NO_FUSE2=1): irregular trees (height-1 nodes under taller parents keep the plain
    kernel), uninstantiated leaves (the staged sum row), unary nodes, ragged categories, a query subset that
    skips some of the fused children."""
    from bnkit_b200 import synth
    rng = np.random.default_rng(57)
    for name, leaves, n_cols, gap in (("LG", 900, 700, 0.0), ("WAG", 600, 500, 0.04)):
        m, om = bnk.Model(name), oracle.Model(name)
        parent, dist = synth.random_tree(leaves, seed=leaves + 3)
        rates = rng.choice([0.3, 1.0, 2.2], size=n_cols, p=[0.6, 0.3, 0.1])
        obs = synth.simulate_alignment(parent, dist, m.parts(), n_cols, rates, seed=leaves + 4, gap_fraction=gap)
        tree = bnk.Tree(parent, dist)
        internal = tree.internal_nodes()
        sub = slice(0, n_cols, 7)
        wpost, wlogl = oracle.fast_marginal(om, parent, dist, np.ascontiguousarray(obs[:, sub]), None, rates[sub], nthreads=8)
        q = internal[::3]
        res = []
        for no_fuse in ("0", "1"):
            monkeypatch.setenv("BNK_NO_FUSE2", no_fuse)
            ws = bnk.Workspace(m, tree, n_cols)
            ws.set_rates(n_cols, rates)
            post, logl = ws.marginal(obs, None)
            pq, _ = ws.marginal(obs, None, query=q, want_logl=False)
            ws.close()
            _post_close(post[:, sub], wpost[internal])
            assert np.allclose(logl[sub], wlogl, rtol=TOL, atol=1e-9)
            res.append((post, pq, logl))
        # same operations in the same order; the compiler contracts multiply-adds differently in the two kernels,
        # so the two paths agree to a few ulp rather than bit for bit
        for other in res[1:]:
            assert np.allclose(res[0][0], other[0], rtol=1e-13, atol=0, equal_nan=True)
            assert np.allclose(res[0][1], other[1], rtol=1e-13, atol=0, equal_nan=True)
            assert np.array_equal(res[0][2], other[2])
        assert np.array_equal(res[0][1], res[0][0][::3], equal_nan=True)


def test_uninstantiated_childless_nodes_get_a_posterior(bnk, oracle):
    """A present but uninstantiated childless node (orphan flooding keeps such ancestors-turned-leaves,
    IdxTree.java:429-432) has a BN variable like any other: queried explicitly it returns the distribution the
    oracle's elimination gives it, not NaN; listing a node twice returns the same row twice."""
    rng = np.random.default_rng(41)
    for name, mask in (("LG", False), ("JTT", True), ("JC", False)):
        m, om = bnk.Model(name), oracle.Model(name)
        parent, dist = random_tree(rng, 30, max_children=3)
        n_cols = 45
        obs = random_obs(rng, parent, n_cols, m.K, gap_prob=0.35, internal_obs_prob=0.05 if mask else 0.0)
        present = None
        if mask:  # gap leaves stay PRESENT in half of the columns, internal nodes are dropped at random
            present = random_present(rng, parent, obs, absent_prob=0.15)
            leaf = leaves_of(parent)
            keep = rng.random(obs.shape) < 0.5
            present[leaf[:, None] & (obs == NONE) & keep] = 1
        tree = bnk.Tree(parent, dist)
        ws = bnk.Workspace(m, tree, n_cols)
        leaves = np.nonzero(leaves_of(parent))[0]
        query = np.concatenate([leaves[:6], tree.internal_nodes()[:3], leaves[:2]]).astype(np.int32)
        post, logl = ws.marginal(obs, present, query=query)
        wpost, wlogl = oracle.fast_marginal(om, parent, dist, obs, present, nthreads=4)
        _post_close(post, wpost[query])
        assert np.isfinite(post[:6]).any(), "no childless query had a posterior: the case is not exercised"
        assert np.allclose(logl, wlogl, rtol=TOL, atol=1e-12)
        ws.close()


def test_big_multifurcations(bnk, oracle):
    """Factorize.getProduct multiplies any number of factors (Factorize.java:331-375): stars with 40 and 70
    children (beyond the 31 / 16 the first version accepted), joint bit-exact incl. the queue-pairing order."""
    rng = np.random.default_rng(43)
    for n_kids in (40, 70):
        # root -> hub with n_kids children, a third of them small subtrees
        parent = [-1, 0, 0]
        hub = 1
        for k in range(n_kids):
            parent.append(hub)
            if k % 3 == 0:
                v = len(parent) - 1
                parent += [v, v]
        parent = np.asarray(parent, dtype=np.int32)
        dist = np.maximum(rng.gamma(1.1, 0.2, size=len(parent)), 1e-7)
        dist[0] = 0.0
        m, om = bnk.Model("WAG"), oracle.Model("WAG")
        n_cols = 40
        obs = random_obs(rng, parent, n_cols, m.K, gap_prob=0.1)
        present = random_present(rng, parent, obs, absent_prob=0.1)
        tree = bnk.Tree(parent, dist)
        ws = bnk.Workspace(m, tree, n_cols)
        for pres in (None, present):
            got = ws.joint(obs, pres)
            assert np.array_equal(got, oracle.fast_joint(om, parent, dist, obs, pres, nthreads=4))
            post, logl = ws.marginal(obs, pres)
            wpost, wlogl = oracle.fast_marginal(om, parent, dist, obs, pres, nthreads=4)
            _post_close(post, wpost[tree.internal_nodes()])
            assert np.allclose(logl, wlogl, rtol=TOL, atol=1e-12)
        ws.close()


def test_bad_state_bytes_are_rejected(bnk):
    """a state byte that is neither 255 nor < K would index past the K x K tables: BNK_ERR_ARG"""
    m = bnk.Model("JC")
    parent = np.array([-1, 0, 0], dtype=np.int32)
    ws = bnk.Workspace(m, bnk.Tree(parent, np.array([0.0, 0.1, 0.2])), 8)
    obs = np.full((3, 8), NONE, np.uint8)
    obs[1:] = 2
    ws.joint(obs)
    obs[2, 5] = 7   # a 20-letter code handed to a 4-state workspace
    for call in (ws.joint, ws.marginal, ws.loglik):
        with pytest.raises(bnk.BnkError) as e:
            call(obs)
        assert e.value.code == -1
    with pytest.raises(ValueError):
        ws.joint(obs[:2])   # short array: caught before the library reads past its end
    ws.close()


def test_configure_keeps_rates_and_unchanged_rates_keep_tables(bnk, oracle):
    """bnk_ws_configure must not silently drop the caller's rates; bnk_set_rates with the same
    assignment must not force a table rebuild (a per-ancestor getMarginal loop calls it every time)."""
    rng = np.random.default_rng(47)
    m, om = bnk.Model("LG"), oracle.Model("LG")
    parent, dist = random_tree(rng, 60)
    n_cols = 80
    obs = random_obs(rng, parent, n_cols, 20)
    rates = rng.choice([0.3, 1.0, 2.5], n_cols)
    ws = bnk.Workspace(m, bnk.Tree(parent, dist), n_cols)
    ws.set_rates(n_cols, rates)
    ws.configure(16, 1)
    _, wl = oracle.fast_marginal(om, parent, dist, obs, None, rates, nthreads=4, want_post=False)
    logl, _ = ws.loglik(obs)
    assert np.allclose(logl, wl, rtol=TOL)
    n0 = ws.launch_count()
    ws.set_rates(n_cols, rates)
    ws.loglik(obs)
    n1 = ws.launch_count()
    ws.set_rates(n_cols, rates * 1.5)
    ws.loglik(obs)
    n2 = ws.launch_count()
    assert n2 - n1 == (n1 - n0) + 1, "a changed assignment rebuilds the tables (one more launch), an unchanged one does not"
    ws.close()


@pytest.mark.parametrize("name", ["LG", "JTT", "WAG", "Dayhoff", "JC", "Yang"])
def test_device_tables_against_an_independent_matrix_exponential(bnk, name):
    """bnk_probs (device P(t)) against scipy.linalg.expm(R t) to 1e-12 -- independent of the transliterated
    eigensolver that both the oracle and the host set-up share."""
    from scipy.linalg import expm
    m = bnk.Model(name)
    R = m.parts()["R"]
    ws = bnk.Workspace(m, bnk.Tree([-1, 0, 0], [0.0, 0.1, 0.2]), 4)
    times = np.array([1e-5, 0.003, 0.01, 0.07, 0.22, 1.0, 3.7, 25.0])
    P, L = ws.probs(times)
    for i, t in enumerate(times):
        want = expm(R * t)
        assert np.abs(P[i] - want).max() < 1e-12, (name, t)
        # log P(t) is the log of the DEVICE's P(t) (the reference takes Math.log of getProbs' entries,
        # SubstNode.java:498-518): entries of 1e-14 carry the eigen-route's absolute error of ~1e-17, so their logs are
        # compared with the device's own P, and with expm only where 1e-12 absolute means <= 1e-9 relative
        pos = P[i] > 0
        assert np.abs(L[i][pos] - np.log(P[i][pos])).max() < 1e-10, (name, t)
        big = want > 1e-3
        assert np.abs(L[i][big] - np.log(want[big])).max() < 1e-9, (name, t)
    ws.close()


@pytest.mark.parametrize("walker", ["", "general"])
def test_masked_walker_per_column_forests(bnk, oracle, walker, monkeypatch):
    """Per-column forests (presence masks, evidence on extants only) through the masked warp-tiled walker (default)
    and the general walker (BNK_WALKER=general): a 700-level caterpillar (walker only, > 20 program chunks), a balanced
    tree (general level kernels + walker for the top), irregular binary trees with per-column rates, uninstantiated
    present extants, fully absent columns and a column where only one extant is present -- against the oracle."""
    from bnkit_b200 import synth
    if walker:
        monkeypatch.setenv("BNK_WALKER", walker)
    rng = np.random.default_rng(71)
    m, om = bnk.Model("LG"), oracle.Model("LG")
    cases = [synth.caterpillar_tree(700, seed=5), synth.balanced_tree(700, seed=6), random_tree(rng, 300), random_tree(rng, 37)]
    for k, (parent, dist) in enumerate(cases):
        n_cols = [97, 130, 64, 33][k]
        obs = random_obs(rng, parent, n_cols, 20, gap_prob=0.35, few_states=7)
        present = random_present(rng, parent, obs, absent_prob=[0.25, 0.3, 0.5, 0.15][k])
        leaf = leaves_of(parent)
        # present extants without a state (maximised / marginalised), an all-absent column, a single present extant
        un = (rng.random(obs.shape) < 0.03) & leaf[:, None] & (present == 1)
        obs[un] = NONE
        present[:, 0] = 0
        present[:, 1] = 0
        present[np.nonzero(leaf)[0][0], 1] = 1
        rates = rng.choice([0.4, 1.0, 2.2], n_cols) if k >= 2 else None
        ws = bnk.Workspace(m, bnk.Tree(parent, dist), n_cols)
        ws.set_rates(n_cols, rates)
        got = ws.joint(obs, present)
        want = oracle.fast_joint(om, parent, dist, obs, present, rates, nthreads=8)
        assert np.array_equal(got, want), "case %d: joint states differ at %d entries" % (k, (got != want).sum())
        post, logl = ws.marginal(obs, present)
        wpost, wlogl = oracle.fast_marginal(om, parent, dist, obs, present, rates, nthreads=8)
        _post_close(post, wpost[~leaf])
        assert np.allclose(logl, wlogl, rtol=TOL, atol=1e-12)
        one, _ = ws.marginal(obs, present, query=[int(np.nonzero(~leaf)[0][-1])], want_logl=False)
        _post_close(one, wpost[[int(np.nonzero(~leaf)[0][-1])]])
        ws.close()
