"""Synthetic trees and alignments of the shapes named in BASELINE.json.

Host-side numpy only (workload generation is not on the hot path).  Trees come out in
dat.phylo.IdxTree order (depth-first pre-order, parent index < child index); branch lengths
are Gamma(shape 1.1, rate 5.0) with a 1e-7 floor like the reference's test generator
Tree.Random(n, seed, 1.1, 5.0, 2, 2) (src/dat/phylo/Tree.java:244-272); leaf states are
simulated down the tree from the model's root prior with each column's rate.
"""
import numpy as np

ALPHABET_AA = "ARNDCQEGHILKMFPSTWYV"  # bn/ctmc/matrix/LG.java:32
NONE = 255


def _branch_lengths(n, rng, shape=1.1, rate=5.0):
    d = rng.gamma(shape, 1.0 / rate, size=n)
    return np.maximum(d, 1e-7)


def balanced_tree(n_leaves, seed=1):
    """Complete binary tree, pre-order indices."""
    parent = []

    def rec(k, par):
        idx = len(parent)
        parent.append(par)
        if k > 1:
            left = (k + 1) // 2
            rec(left, idx)
            rec(k - left, idx)

    import sys
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
    rec(n_leaves, -1)
    parent = np.asarray(parent, dtype=np.int32)
    rng = np.random.default_rng(seed)
    dist = _branch_lengths(len(parent), rng)
    dist[0] = 0.0
    return parent, dist


def caterpillar_tree(n_leaves, seed=2):
    """Every internal node = (leaf, rest): depth n_leaves - 1."""
    assert n_leaves >= 2
    parent = [-1]
    cur = 0
    for i in range(n_leaves - 1):
        parent.append(cur)            # leaf child
        if i < n_leaves - 2:
            parent.append(cur)        # next internal node
            cur = len(parent) - 1
        else:
            parent.append(cur)        # last leaf
    parent = np.asarray(parent, dtype=np.int32)
    rng = np.random.default_rng(seed)
    dist = _branch_lengths(len(parent), rng)
    dist[0] = 0.0
    return parent, dist


def random_tree(n_leaves, seed=5, max_children=2):
    """Random agglomeration (join random clusters until one remains), then pre-order indexing."""
    rng = np.random.default_rng(seed)
    children = {}
    clusters = list(range(n_leaves))  # ids < n_leaves are leaves
    nxt = n_leaves
    while len(clusters) > 1:
        k = 2 if max_children <= 2 else int(rng.integers(2, max_children + 1))
        k = min(k, len(clusters))
        picks = sorted(rng.choice(len(clusters), size=k, replace=False).tolist(), reverse=True)
        kids = [clusters.pop(i) for i in picks]
        children[nxt] = kids[::-1]
        clusters.append(nxt)
        nxt += 1
    root = clusters[0]
    parent = []
    stack = [(root, -1)]
    while stack:
        node, par = stack.pop()
        idx = len(parent)
        parent.append(par)
        for c in reversed(children.get(node, [])):
            stack.append((c, idx))
    parent = np.asarray(parent, dtype=np.int32)
    dist = _branch_lengths(len(parent), rng)
    dist[0] = 0.0
    return parent, dist


def gamma_category_rates(alpha=0.5, ncat=4):
    """Mean rates of ncat equiprobable categories of Gamma(alpha, alpha) (Yang 1994)."""
    from scipy.special import gammainc
    from scipy.stats import gamma as gdist
    cuts = gdist.ppf(np.arange(1, ncat) / ncat, a=alpha, scale=1.0 / alpha)
    edges = np.concatenate([[0.0], cuts, [np.inf]])
    upper = gammainc(alpha + 1.0, edges * alpha)  # integral of x f(x) up to the edge (mean = 1)
    return (upper[1:] - upper[:-1]) * ncat


def transition_matrices(parts, times):
    """P(t) = |V exp(L t) V^-1| for an array of times (numpy; used only to *simulate* data)."""
    ev, iev, lam = parts["evec"], parts["ievec"], parts["eval"]
    e = np.exp(np.asarray(times)[:, None] * lam[None, :])
    return np.abs(np.einsum("ik,nk,kj->nij", ev, e, iev))


def simulate_alignment(parent, dist, parts, n_cols, rates=None, seed=3, gap_fraction=0.0):
    """Leaf states [n_nodes][n_cols] (internal rows = NONE).  rates: per-column or None."""
    rng = np.random.default_rng(seed)
    n = len(parent)
    K = len(parts["prior"])
    rates = np.ones(n_cols) if rates is None else np.asarray(rates, dtype=np.float64)
    cats, cat_idx = np.unique(rates, return_inverse=True)
    has_child = np.zeros(n, dtype=bool)
    has_child[parent[parent >= 0]] = True
    state = np.empty((n, n_cols), dtype=np.uint8)
    prior_cdf = np.cumsum(parts["prior"] / parts["prior"].sum())
    cols = np.arange(n_cols)
    for v in range(n):
        u = rng.random(n_cols)
        if parent[v] < 0:
            state[v] = np.minimum((u[:, None] > prior_cdf[None, :]).sum(1), K - 1)
            continue
        P = transition_matrices(parts, dist[v] * cats)  # [ncat][K][K]
        cdf = np.cumsum(P / P.sum(-1, keepdims=True), axis=-1)
        row = cdf[cat_idx, state[parent[v]]]
        state[v] = np.minimum((u[:, None] > row).sum(1), K - 1)
    del cols
    obs = np.full((n, n_cols), NONE, dtype=np.uint8)
    obs[~has_child] = state[~has_child]
    if gap_fraction > 0:
        gaps = rng.random((n, n_cols)) < gap_fraction
        obs[gaps & ~has_child[:, None]] = NONE
    return obs


def path_presence_mask(parent, leaf_present):
    """Stand-in indel rule (NOT the reference's BEP inference): an
    ancestor is present in a column iff it lies on the tree path between two present leaves
    (or is itself a present leaf).  leaf_present: [n_nodes][n_cols] bool, True only at leaves."""
    n, n_cols = leaf_present.shape
    cnt = leaf_present.astype(np.int32).copy()  # present leaves in the subtree
    for v in range(n - 1, 0, -1):
        cnt[parent[v]] += cnt[v]
    total = cnt[parent < 0].sum(0) if (parent < 0).sum() > 1 else cnt[0]
    # child subtrees holding >= 1 present leaf
    nz_children = np.zeros((n, n_cols), dtype=np.int32)
    for v in range(1, n):
        nz_children[parent[v]] += cnt[v] > 0
    inside = (cnt > 0) & (cnt < total[None, :])   # some present leaf below and some elsewhere
    branching = nz_children >= 2
    present = leaf_present | branching | (inside & (nz_children >= 1) & ~leaf_present)
    has_child = np.zeros(n, dtype=bool)
    has_child[parent[parent >= 0]] = True
    present[~has_child] = leaf_present[~has_child]
    return present.astype(np.uint8)
