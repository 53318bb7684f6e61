"""Shared generators for the parity tests (seeded, small)."""
import numpy as np

NONE = 255


def random_tree(rng, n_leaves, max_children=2, zero_branch_prob=0.0):
    """Random agglomeration tree in IdxTree order; returns (parent, dist)."""
    children = {}
    clusters = list(range(n_leaves))
    nxt = n_leaves
    while len(clusters) > 1:
        k = 2 if max_children <= 2 else int(rng.integers(2, max_children + 1))
        k = min(k, len(clusters))
        picks = sorted(rng.choice(len(clusters), size=k, replace=False).tolist(), reverse=True)
        kids = [clusters.pop(i) for i in picks]
        children[nxt] = kids[::-1]
        clusters.append(nxt)
        nxt += 1
    parent = []
    stack = [(clusters[0], -1)]
    while stack:
        node, par = stack.pop()
        idx = len(parent)
        parent.append(par)
        for c in reversed(children.get(node, [])):
            stack.append((c, idx))
    parent = np.asarray(parent, dtype=np.int32)
    dist = np.maximum(rng.gamma(1.1, 0.2, size=len(parent)), 1e-7)
    if zero_branch_prob > 0:
        dist[rng.random(len(parent)) < zero_branch_prob] = 1e-5  # Newick.java:54-56 maps 0.0 to 1e-5
    dist[0] = 0.0
    return parent, dist


def leaves_of(parent):
    has_child = np.zeros(len(parent), dtype=bool)
    has_child[parent[parent >= 0]] = True
    return ~has_child


def random_obs(rng, parent, n_cols, K, gap_prob=0.0, internal_obs_prob=0.0, few_states=None):
    """obs [n][n_cols]: leaves observed (or gap -> NONE), internal nodes optionally observed."""
    n = len(parent)
    leaf = leaves_of(parent)
    hi = K if few_states is None else few_states
    obs = np.full((n, n_cols), NONE, dtype=np.uint8)
    vals = rng.integers(0, hi, size=(n, n_cols)).astype(np.uint8)
    obs[leaf] = vals[leaf]
    if gap_prob > 0:
        g = rng.random((n, n_cols)) < gap_prob
        obs[g & leaf[:, None]] = NONE
    if internal_obs_prob > 0:
        g = rng.random((n, n_cols)) < internal_obs_prob
        sel = g & ~leaf[:, None]
        obs[sel] = vals[sel]
    return obs


def random_present(rng, parent, obs, absent_prob=0.3):
    """A per-column presence mask: gap leaves are absent, internal nodes randomly absent."""
    n, n_cols = obs.shape
    leaf = leaves_of(parent)
    present = np.ones((n, n_cols), dtype=np.uint8)
    present[leaf[:, None] & (obs == NONE)] = 0
    drop = rng.random((n, n_cols)) < absent_prob
    present[drop & ~leaf[:, None]] = 0
    return present


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.maximum(np.abs(b), 1e-300)
    return np.abs(a - b) / denom
