"""Bi-directional edge parsimony: oracle/bep.py against the reference's documented output, the
bit-set formulation the device kernels use against the oracle's explicit Sankoff tables (CPU), and
bnk_bep_presence (C ABI, device sweeps) against the oracle (GPU)."""
import json
import os

import numpy as np
import pytest

from helpers import NONE, random_tree

HERE = os.path.dirname(os.path.abspath(__file__))
AA = "ARNDCQEGHILKMFPSTWYV"


def gold(name):
    with open(os.path.join(HERE, "golden", name)) as fh:
        return json.load(fh)


def fixture_rows(fx):
    """(parent, rows, n_pos, obs): alignment rows placed at their tree nodes."""
    parent = fx["tree"]["Parents"]
    labels = fx["tree"]["Labels"]
    n, n_pos = len(parent), len(fx["seqs"][0])
    rows = [None] * n
    obs = np.full((n, n_pos), NONE, np.uint8)
    for name, seq in zip(fx["names"], fx["seqs"]):
        v = labels.index(name)
        rows[v] = [ch != "-" for ch in seq]
        for c, ch in enumerate(seq):
            if ch != "-":
                obs[v, c] = AA.index(ch)
    return parent, rows, n_pos, obs


def random_case(rng, n_leaves, n_pos, gap, max_children=2):
    parent, dist = random_tree(rng, n_leaves, max_children=max_children)
    has_child = np.zeros(len(parent), bool)
    has_child[parent[parent >= 0]] = True
    obs = np.full((len(parent), n_pos), NONE, np.uint8)
    rows = [None] * len(parent)
    for v in range(len(parent)):
        if not has_child[v]:
            occ = rng.random(n_pos) > gap
            rows[v] = occ.tolist()
            obs[v, occ] = rng.integers(0, 20, size=int(occ.sum()))
    return parent, dist, rows, obs


def test_oracle_reproduces_documented_recon_graphs():
    """docs/json-api.md:113-122: ancestors 0, 1, 2 of data/66.* get nodes {2, 4}, edges (-1,2), (2,4), (4,5)."""
    from oracle import bep
    parent, rows, n_pos, _ = fixture_rows(gold("c1_66.json"))
    mask, anc = bep.presence_mask(parent, rows, n_pos)
    doc = gold("c1_66.json")["doc_result"]
    for j in (0, 1, 2):
        assert [c for c in range(n_pos) if anc[j].node[c]] == doc["indices"]
        assert sorted(anc[j].edges()) == [(-1, 2), (2, 4), (4, 5)]
        assert sorted(anc[j].start) == [2] and sorted(anc[j].end) == [4]


def _bitset_parsimony(children, inst, recode_null):
    """bep.cu's forward / backward on Python ints (A = arg-min set, B = states one above it)."""
    n = len(children)
    vals = []
    for v in inst:
        if v is not None and v not in vals:
            vals.append(v)
    ns = len(vals)
    ntot = ns + (1 if recode_null else 0)
    valid = (1 << ntot) - 1
    A, B, opt, parent = [0] * n, [0] * n, [0] * n, [-1] * n
    for v in range(n - 1, -1, -1):
        for u in children[v]:
            parent[u] = v
        if not children[v] or ntot == 0:
            continue
        cnt = [0] * ntot
        for u in children[v]:
            Au = A[u] if children[u] else ((1 << ns) if recode_null else valid) if inst[u] is None else 1 << vals.index(inst[u])
            for i in range(ntot):
                cnt[i] += not (Au >> i) & 1
        A[v] = sum(1 << i for i in range(ntot) if cnt[i] == min(cnt))
        B[v] = sum(1 << i for i in range(ntot) if cnt[i] == min(cnt) + 1)
    out = [None] * n
    for v in range(n):
        if children[v]:
            ob = opt[parent[v]] if parent[v] >= 0 else 0
            opt[v] = A[v] if parent[v] < 0 else (ob & (A[v] | B[v])) | (A[v] if ob & ~A[v] else 0)
            out[v] = sorted(vals[i] for i in range(ns) if (opt[v] >> i) & 1)
    return out


def test_parsimony_known_answers_of_the_reference_tests():
    """asr.Parsimony's all-optimal sets on the hand-checked cases of ParsimonyTests.testParsimony2 / testParsimony3
    (test/dat/phylo/ParsimonyTests.java:29-30, 93-186; `infer(ti, false)`: no NULL recoding) -- through the restatement
    and through the bit-set form the device sweeps use."""
    from bnkit_b200 import phylo
    from oracle import bep
    t = phylo.IdxTree.fromNewick("((((x01,x02)X01_02,x03)X01_03,(x04,((x05,x06)X05_06,(x07,x08)X07_08)X05_08)X04_08)X01_08,"
                                 "(x09,(x10,x11)X10_11)X09_11)X01_11;")
    n = t.getSize()
    children = [list(t.getChildren(i)) for i in range(n)]
    is_leaf = [not c for c in children]
    names = ["x%02d" % i for i in range(1, 12)]
    a, b, c, d, e, f, g = "ABCDEFG"

    def optimal(values, leaves=names):
        inst = [None] * n
        for name, v in zip(leaves, values):
            inst[t.labels.index(name)] = v
        ref = bep.parsimony(children, inst, is_leaf, False)
        bits = _bitset_parsimony(children, inst, False)
        for v in range(n):
            if children[v]:
                assert sorted(ref[v] or []) == bits[v]
        return {lab: set(ref[t.labels.index(lab)]) for lab in ("X01_03", "X04_08")}
    o = optimal([a, a, c, d, b, b, b, b, e, f, g])                 # init1: every symbol optimal at both nodes
    assert o["X01_03"] == set("ABCDEFG") and o["X04_08"] == set("ABCDEFG")
    o = optimal([a, a, c, d, b, b, b, b, e, b, a])                 # init2
    assert {a, b, c} <= o["X01_03"] and d not in o["X01_03"]
    assert {a, b, d} <= o["X04_08"] and c not in o["X04_08"]
    o = optimal([a, a, c, d, b, b, b, b, e, b, f])                 # init3
    assert {a, b, c} <= o["X01_03"]
    assert b in o["X04_08"] and not ({a, c, d} & o["X04_08"])
    o = optimal([a, b, c], names[:3])                              # testParsimony3: the other eight leaves uninstantiated
    assert {a, b, c} <= o["X01_03"]


def test_bitset_formulation_equals_sankoff_tables():
    """Unit costs collapse asr.Parsimony's score tables and traceback lists to two bit sets per node."""
    from oracle import bep
    rng = np.random.default_rng(1)
    checked = 0
    for trial in range(60):
        parent, _, rows, _ = random_case(rng, int(rng.integers(3, 22)), int(rng.integers(2, 10)),
                                         float(rng.choice([0.2, 0.5, 0.7])), max_children=int(rng.integers(2, 5)))
        children = bep.children_of(parent.tolist())
        is_leaf = [len(c) == 0 for c in children]
        n_pos = len(rows[[i for i, r in enumerate(rows) if r is not None][0]])
        for recode in (True, False):
            for i in range(-1, n_pos + 1):
                for fwd in (True, False):
                    inst = bep.edge_instance(rows, i, fwd, n_pos)
                    if all(x is None for x in inst):
                        continue
                    ref = bep.parsimony(children, inst, is_leaf, recode)
                    got = _bitset_parsimony(children, inst, recode)
                    for v in range(len(parent)):
                        if children[v]:
                            assert sorted(ref[v] or []) == got[v]
                            checked += 1
    assert checked > 5000


@pytest.mark.gpu
def test_gpu_bep_c1_c2(bnk):
    from oracle import bep
    for name in ("c1_66.json", "c2_3_2_1_1_filt.json"):
        parent, rows, n_pos, obs = fixture_rows(gold(name))
        want, _ = bep.presence_mask(parent, rows, n_pos)
        tree = bnk.Tree(np.asarray(parent, np.int32), np.asarray(gold(name)["tree"]["Distances"]))
        got, info = tree.bep_presence(obs, want_info=True)
        assert np.array_equal(got, np.asarray(want, np.uint8)), "%s: %d entries differ" % (name, (got != np.asarray(want)).sum())
        assert info[3] >= 4


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", ["levels", "instance"])
def test_gpu_bep_random_alignments(bnk, schedule, monkeypatch):
    """Random trees (binary and multifurcating) and gappy alignments, including discontinuous ancestors that
    go through the patch rounds; one case with more than 31 distinct edge targets (two-word bit sets).  Both
    schedules of the device sweeps: one launch per height level, and one thread per instance walking the tree."""
    from oracle import bep
    monkeypatch.setenv("BNK_BEP_SCHEDULE", schedule)
    rng = np.random.default_rng(7)
    patched = 0
    for trial in range(40):
        parent, dist, rows, obs = random_case(rng, int(rng.integers(3, 40)), int(rng.integers(1, 30)),
                                              float(rng.choice([0.05, 0.3, 0.6, 0.85])), max_children=int(rng.integers(2, 6)))
        n_pos = obs.shape[1]
        want, _ = bep.presence_mask(parent.tolist(), rows, n_pos)
        got, info = bnk.Tree(parent, dist).bep_presence(obs, want_info=True)
        assert np.array_equal(got, np.asarray(want, np.uint8)), "trial %d: %d entries differ" % (trial, (got != np.asarray(want)).sum())
        patched += int(info[0] > 0)
    assert patched > 0
    # 45 sequences that all occupy column 0 and then one column each: 45 distinct targets from column 0
    parent, dist = random_tree(rng, 45)
    has_child = np.zeros(len(parent), bool)
    has_child[parent[parent >= 0]] = True
    n_pos = 50
    obs = np.full((len(parent), n_pos), NONE, np.uint8)
    rows = [None] * len(parent)
    for k, v in enumerate(np.nonzero(~has_child)[0]):
        obs[v, 0] = 0
        obs[v, 1 + k] = 1
        rows[v] = (obs[v] != NONE).tolist()
    want, _ = bep.presence_mask(parent.tolist(), rows, n_pos)
    got, info = bnk.Tree(parent, dist).bep_presence(obs, want_info=True)
    assert info[2] >= 45
    assert np.array_equal(got, np.asarray(want, np.uint8))


@pytest.mark.gpu
def test_gpu_bep_edge_cases(bnk):
    """Degenerate inputs: a single column, all-gap columns, an empty sequence, every sequence empty, two leaves,
    a star tree, a gap-free alignment (every ancestor present everywhere)."""
    from oracle import bep
    rng = np.random.default_rng(3)

    def check(parent, obs):
        parent = np.asarray(parent, np.int32)
        has_child = np.zeros(len(parent), bool)
        has_child[parent[parent >= 0]] = True
        rows = [None if has_child[v] else (obs[v] != NONE).tolist() for v in range(len(parent))]
        want, _ = bep.presence_mask(parent.tolist(), rows, obs.shape[1])
        got = bnk.Tree(parent, np.full(len(parent), 0.1)).bep_presence(obs)
        assert np.array_equal(got, np.asarray(want, np.uint8))
        return got

    two = [-1, 0, 0]
    for n_pos in (1, 2, 7):
        for trial in range(6):
            obs = np.full((3, n_pos), NONE, np.uint8)
            obs[1:] = np.where(rng.random((2, n_pos)) < 0.5, 0, NONE)
            check(two, obs)
    parent, _ = random_tree(rng, 12)
    leaves = np.nonzero(np.bincount(parent[parent >= 0], minlength=len(parent)) == 0)[0]
    obs = np.full((len(parent), 9), NONE, np.uint8)
    check(parent, obs)                                   # every sequence empty
    obs[leaves] = 3
    got = check(parent, obs)                             # no gaps at all
    assert got.all()
    obs[leaves[0]] = NONE                                # one empty sequence
    obs[:, 4] = NONE                                     # an all-gap column
    got = check(parent, obs)
    assert not got[:, 4].any()
    star = [-1] + [0] * 9                                # one ancestor, nine children
    obs = np.full((10, 6), NONE, np.uint8)
    obs[1:] = np.where(rng.random((9, 6)) < 0.6, 1, NONE)
    check(star, obs)


@pytest.mark.gpu
def test_gpu_bep_feeds_the_sweeps(bnk, oracle):
    """End to end on data/66.*: BEP mask (device) -> orphan pruning -> joint reconstruction (device) equals the
    oracle chain, and reproduces the documented states A, H of ancestors 0-2 under JTT."""
    from oracle import bep
    fx = gold("c1_66.json")
    parent, rows, n_pos, obs = fixture_rows(fx)
    parent = np.asarray(parent, np.int32)
    dist = np.asarray(fx["tree"]["Distances"])
    tree = bnk.Tree(parent, dist)
    present = tree.prune_orphans(tree.bep_presence(obs))
    m = bnk.Model("JTT")
    ws = bnk.Workspace(m, tree, n_pos)
    st = ws.joint(obs, present)
    want_mask, _ = bep.presence_mask(parent.tolist(), rows, n_pos)
    want_present = np.stack([oracle.prune(parent, np.ascontiguousarray(np.asarray(want_mask, np.uint8)[:, c]), True)[0]
                             for c in range(n_pos)], axis=1)
    assert np.array_equal(present, want_present)
    assert np.array_equal(st, oracle.fast_joint(oracle.Model("JTT"), parent, dist, obs, present))
    for anc in (0, 1, 2):
        cols = [c for c in range(n_pos) if present[anc, c]]
        assert cols == fx["doc_result"]["indices"]
        assert [AA[st[anc, c]] for c in cols] == fx["doc_result"]["values"]
    ws.close()


@pytest.mark.gpu
def test_gpu_bep_device_assembly_equals_host_assembly(bnk, monkeypatch):
    """Graph assembly on the device (bep_reach.cu: alive1 / alive2 passes, only ancestors without a path go back to
    the host) against the host assembly of the same optimal sets (BNK_BEP_ASSEMBLY=host, the r1 path) and against
    the oracle: larger alignments than the random cases above -- columns beyond one 32-column tile, several
    ancestors per CTA, clade-wise indels that leave many ancestors without a path"""
    from oracle import bep
    rng = np.random.default_rng(31)
    crippled = 0
    for trial in range(6):
        parent, dist, rows, obs = random_case(rng, int(rng.integers(40, 120)), int(rng.integers(33, 150)),
                                              float(rng.choice([0.3, 0.6])), max_children=int(rng.integers(2, 4)))
        n = len(parent)
        size = np.ones(n, np.int64)
        for v in range(n - 1, 0, -1):
            size[parent[v]] += size[v]
        for c in range(0, obs.shape[1], 2):                  # clade-wise deletions
            v = int(rng.integers(1, n))
            obs[v:v + size[v], c] = NONE
        has_child = np.zeros(n, bool)
        has_child[parent[parent >= 0]] = True
        rows = [None if has_child[v] else (obs[v] != NONE).tolist() for v in range(n)]
        tree = bnk.Tree(parent, dist)
        monkeypatch.delenv("BNK_BEP_ASSEMBLY", raising=False)
        got, info = tree.bep_presence(obs, want_info=True)
        monkeypatch.setenv("BNK_BEP_ASSEMBLY", "host")
        host, hinfo = tree.bep_presence(obs, want_info=True)
        assert np.array_equal(got, host) and info[0] == hinfo[0] and info[1] == hinfo[1]
        crippled += int(info[0])
        if trial < 3:
            want, _ = bep.presence_mask(parent.tolist(), rows, obs.shape[1])
            assert np.array_equal(got, np.asarray(want, np.uint8))
    assert crippled > 0


@pytest.mark.gpu
def test_gpu_bep_wide_multifurcations(bnk):
    """more than 63 children under one ancestor (r1 refused them: 6-bit child counters): a 100-way star and a tree
    with a 70-way node below the root, both assembly paths"""
    from oracle import bep
    rng = np.random.default_rng(12)
    star = np.array([-1] + [0] * 100, np.int32)
    mixed = np.array([-1, 0, 0] + [2] * 70 + [1, 1], np.int32)
    for parent in (star, mixed):
        n = len(parent)
        has_child = np.zeros(n, bool)
        has_child[parent[parent >= 0]] = True
        obs = np.full((n, 12), NONE, np.uint8)
        obs[~has_child] = np.where(rng.random((int((~has_child).sum()), 12)) < 0.55, 2, NONE)
        rows = [None if has_child[v] else (obs[v] != NONE).tolist() for v in range(n)]
        want, _ = bep.presence_mask(parent.tolist(), rows, 12)
        tree = bnk.Tree(parent, np.full(n, 0.1))
        assert np.array_equal(tree.bep_presence(obs), np.asarray(want, np.uint8))
        mask, graphs = tree.bep_infer(obs)
        assert np.array_equal(mask, np.asarray(want, np.uint8))
