"""Host-side logic that needs no GPU: the Java-API mirror's parsers and tree bookkeeping, the
synthetic workload generators, and the column sharding used by the multi-GPU path."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_newick_matches_documented_idxtree(bnk):
    from bnkit_b200.phylo import IdxTree
    c1 = json.load(open(os.path.join(ROOT, "tests", "golden", "c1_66.json")))
    t = IdxTree.fromJSON(c1["tree"])
    assert t.getSize() == 131 and len(t.getLeaves()) == 66 and t.getNRoots() == 1
    assert [int(p) for p in t.parent] == c1["doc_parents"]         # docs/json-api.md:103-107
    # ancestor labels N<k> in depth-first order <-> the doc's "0", "1", ...
    anc = [l for l in c1["doc_labels"] if l.isdigit()]
    mine = [t.getLabel(i) for i in t.getAncestors()]
    assert mine[: len(anc)] == ["N" + a for a in anc]
    assert IdxTree.fromJSON(t.toJSON()).toJSON() == t.toJSON()


def test_newick_zero_distance_and_labels(bnk):
    from bnkit_b200.phylo import IdxTree
    t = IdxTree.fromNewick("((a:0.0,b:0.5)x:0.25,(c:1e-3,d:2):0.0,e:3);")
    assert t.getSize() == 8 and [int(p) for p in t.parent] == [-1, 0, 1, 1, 0, 4, 4, 0]
    assert t.getLabel(1) == "x" and t.getLabel(0) == "N0" and t.getLabel(4) == "N2"
    assert t.getDistance(2) == 0.00001 and t.getDistance(4) == 0.00001   # Newick.java:54-56, :103-105
    assert t.getChildren(0) == [1, 4, 7] and t.isLeaf(7) and t.isConnected(0)
    with pytest.raises(RuntimeError):
        IdxTree.fromNewick("(a:xyz,b:1);")


def test_pruned_tree_fixture_through_the_mirror(bnk):
    from bnkit_b200.phylo import IdxTree
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "idxtree_prune.json")))
    t = IdxTree(fx["parents"])
    for case in fx["cases"]:
        idx = t.getPrunedIndex(set(case["prune"]), removeOrphans=False)
        pruned = IdxTree.createPrunedTree(t, idx)
        assert pruned.getSize() == t.getSize() - case["removed"] and pruned.getNRoots() == case["roots"]
        if case["same_size_with_orphan_removal"]:
            idx2 = t.getPrunedIndex(set(case["prune"]), removeOrphans=True)
            assert IdxTree.createPrunedTree(t, idx2).getSize() == pruned.getSize()


def test_substmodel_registry(bnk):
    from bnkit_b200.phylo import SubstModel
    assert SubstModel.createModel("nope") is None
    m = SubstModel.createModel("JTT")
    assert "".join(m.getDomain()) == "ARNDCQEGHILKMFPSTWYV" and m.getName() == "JTT"
    assert "".join(SubstModel.createModel("JC").getDomain()) == "ACGT"


def test_synthetic_generators(bnk):
    from bnkit_b200 import lib, synth
    for maker, n_leaves in ((synth.balanced_tree, 37), (synth.caterpillar_tree, 37), (synth.random_tree, 37)):
        parent, dist = maker(n_leaves)
        assert len(parent) == 2 * n_leaves - 1 and (parent[1:] < np.arange(1, len(parent))).all() and parent[0] == -1
        assert (dist[1:] >= 1e-7).all()
    parent, _ = synth.balanced_tree(1024)
    depth = np.zeros(len(parent), int)
    for v in range(1, len(parent)):
        depth[v] = depth[parent[v]] + 1
    assert depth.max() == 10
    parent, _ = synth.caterpillar_tree(100)
    depth = np.zeros(len(parent), int)
    for v in range(1, len(parent)):
        depth[v] = depth[parent[v]] + 1
    assert depth.max() == 99
    r = synth.gamma_category_rates(0.5, 4)
    assert np.isclose(r.mean(), 1.0) and (np.diff(r) > 0).all()
    parts = lib.Model("LG").parts()
    parent, dist = synth.balanced_tree(16)
    obs = synth.simulate_alignment(parent, dist, parts, 50, r[np.arange(50) % 4], seed=1)
    leaf = np.ones(len(parent), bool); leaf[parent[parent >= 0]] = False
    assert (obs[leaf] < 20).all() and (obs[~leaf] == 255).all()
    mask = synth.path_presence_mask(parent, (obs != 255) & leaf[:, None])
    assert mask.all()     # no gaps: every ancestor lies on a path between present leaves


def test_shard_columns(bnk):
    from bnkit_b200.dist import shard_columns
    for n_cols in (1, 7, 5000, 2001):
        for world in (1, 2, 3, 8):
            blocks = [shard_columns(n_cols, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n_cols
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_rates_file_as_grasp_reads_it(tmp_path):
    """GRASP -rf: "Rate" column (else the first), optional 1-based "Site" column (src/asr/GRASP.java:433-456)."""
    from bnkit_b200.phylo import ASRRuntimeException, load_rates
    f = tmp_path / "r.tsv"
    f.write_text("# IQ-TREE style\nSite\tRate\tCat\n2\t0.5\t1\n1\t2.25\t3\n3\t1e-3\t0\n")
    assert list(load_rates(str(f))) == [2.25, 0.5, 1e-3]
    f.write_text("Rate\n1.0\n0.25\n")
    assert list(load_rates(str(f))) == [1.0, 0.25]
    f.write_text("whatever\tx\n0.7\ta\n1.3\tb\n")
    assert list(load_rates(str(f))) == [0.7, 1.3]
    f.write_text("Site\tRate\n1\tabc\n")
    with pytest.raises(ASRRuntimeException):
        load_rates(str(f))


def test_ancestor_numbers_follow_the_reference_rules():
    """Newick.parse (src/dat/file/Newick.java:217-254): a label N<number> fixes the ancestor's number, all other
    ancestors (unlabelled or e.g. bootstrap values) take the lowest free numbers depth-first; IdxTree.fromJSON
    (src/dat/phylo/IdxTree.java:312-321) numbers the ancestors 0, 1, ... in index order whatever their labels;
    IdxTree.getIndex (:997-1022) finds "N<k>" and k."""
    from bnkit_b200.phylo import IdxTree
    t = IdxTree.fromNewick("((a:1,b:1)0.9:1,((c:1,d:1)N1:1,e:1)0.9:1)N3;")
    anc = t.getAncestors()
    assert [t.getID(i) for i in anc] == [3, 0, 2, 1]
    assert t.getIndex("N2") == anc[2] and t.getIndex(1) == anc[3] and t.getIndex("N3") == 0 and t.getIndex("c") == t.labels.index("c")
    assert t.getIndex("N9") == -1 and t.getIndex("zz") == -1
    j = IdxTree.fromJSON(t.toJSON())
    assert [j.getID(i) for i in j.getAncestors()] == [0, 1, 2, 3]
    import pytest
    with pytest.raises(RuntimeError):
        IdxTree.fromNewick("((a:1,b:1)N1:1,(c:1,d:1)N1:1);")
    with pytest.raises(RuntimeError):
        IdxTree.fromNewick("((a:1,b:1)Nx:1,c:1);")


@pytest.mark.parametrize("fixture, model_name", [("c1_66.json", "LG"), ("c2_3_2_1_1_filt.json", "JTT")])
def test_node_instances_follow_the_alignment_occupancy(bnk, fixture, model_name):
    """POGTreeTest.getNodeInstance (test/dat/pog/POGTreeTest.java:44-53): per column, the nodes that carry an instance
    are exactly the extants with a residue there -- what the `obs` matrix of the C ABI holds (row a4 of SURVEY 8), and
    the residues are the alignment's own in the model's alphabet order (SubstNodeTest's P(sym | sym) indexing)."""
    from bnkit_b200.phylo import IdxTree, Prediction, SubstModel
    with open(os.path.join(os.path.dirname(__file__), "golden", fixture)) as fh:
        d = json.load(fh)
    tree = IdxTree.fromJSON(d["tree"])
    n_cols = len(d["seqs"][0])
    states = [[None] * n_cols for _ in range(tree.getSize())]
    for name, seq in zip(d["names"], d["seqs"]):
        states[tree.getIndex(name)] = [None if ch == "-" else ch for ch in seq]
    model = SubstModel.createModel(model_name)
    obs = Prediction(tree, states)._obs(model)
    occupancy = np.array([sum(seq[c] != "-" for seq in d["seqs"]) for c in range(n_cols)])
    assert np.array_equal((obs != 255).sum(axis=0), occupancy)
    assert not (obs[[v for v in range(tree.getSize()) if not tree.isLeaf(v)]] != 255).any()   # GRASP instantiates extants only
    dom = list(model.domain)
    for name, seq in zip(d["names"], d["seqs"]):
        row = obs[tree.getIndex(name)]
        assert all((row[c] == 255) if ch == "-" else (dom[row[c]] == ch) for c, ch in enumerate(seq))


def test_tree_instance_encode_decode_as_the_reference_test():
    """TreeInstanceTest.encode1 / encode2 / decode1 / decode2 (test/dat/phylo/TreeInstanceTest.java:40-150): seven
    leaves carry 11, 11, 22, 11, 33, 33, 33; encoding renumbers them, gives null leaves the extra symbol only when asked
    to, never touches ancestors; decoding with the key restores every assigned value."""
    from bnkit_b200.phylo import IdxTree, TreeInstance
    with open(os.path.join(os.path.dirname(__file__), "golden", "c1_66.json")) as fh:
        tree = IdxTree.fromJSON(json.load(fh)["tree"])
    leaves = [i for i in range(tree.getSize()) if tree.isLeaf(i)]
    assign = dict(zip(leaves[:7], [11, 11, 22, 11, 33, 33, 33]))
    unique = len(set(assign.values()))
    for recode in (True, False):
        ti = TreeInstance(tree, dict(assign))
        assert all(ti.getInstance(i) == v for i, v in assign.items())
        key = ti.encode(recode)
        assert len(key) == unique + (1 if recode else 0) and (key[-1] is None) == recode
        for i in range(tree.getSize()):
            if i in assign:
                assert ti.getInstance(i) != assign[i] and key[ti.getInstance(i)] == assign[i]
            elif ti.getInstance(i) is not None:           # newly assigned: a leaf, the symbol after the last value
                assert recode and tree.isLeaf(i) and ti.getInstance(i) == unique
        n_null_leaves = sum(tree.isLeaf(i) and ti.getInstance(i) is None for i in range(tree.getSize()))
        assert n_null_leaves == (0 if recode else len(leaves) - 7)
        ti.decode(key)
        assert all(ti.getInstance(i) == v for i, v in assign.items())
        assert sum(tree.isLeaf(i) and ti.getInstance(i) is None for i in range(tree.getSize())) == len(leaves) - 7


def test_enum_distrib_json_roundtrip():
    """EnumDistribTest.toJSON (test/bn/prob/EnumDistribTest.java:10-15): toJSON -> fromJSON -> toJSON is the identity;
    predefined domains travel by name (Enumerable.java:248-283)."""
    from bnkit_b200.phylo import ASRRuntimeException, EnumDistrib
    rng = np.random.default_rng(1)
    for dom in ("ACGT", "ACDEFGHIKLMNPQRSTVWY", "ARNDCQEGHILKMFPSTWYV", "AC"):
        p = rng.random(len(dom))
        d = EnumDistrib(list(dom), p / p.sum())
        j = d.toJSON()
        assert ("Predef" in j["Domain"]) == (dom in ("ACGT", "ACDEFGHIKLMNPQRSTVWY"))
        r = EnumDistrib.fromJSON(json.loads(json.dumps(j)))
        assert json.dumps(r.toJSON()) == json.dumps(j) and r.domain == list(dom) and np.array_equal(r.distrib, d.distrib)
    with pytest.raises(ASRRuntimeException):
        EnumDistrib.fromJSON({"Domain": {"Predef": "Klingon"}, "Pr": [1.0]})
