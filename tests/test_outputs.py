"""Downstream assembly and outputs: ancestors' partial-order graphs, consensus paths,
sequences, the ancestors FASTA, the DISTRIB table and ASR.json -- `bnkit_b200/pog.py` and the output half of
`bnkit_b200/phylo.Prediction`, against the documented Recon result (docs/json-api.md:113-122, fixture
tests/golden/c1_66.json "doc_ancestors"), the oracle's graphs (oracle/bep.py) and brute-force path enumeration.
CPU tests drive the host logic with the oracle's inference results; GPU tests run the device chain."""
import itertools
import json
import os

import numpy as np
import pytest

from helpers import NONE
from test_bep import fixture_rows, gold, random_case

AA = "ARNDCQEGHILKMFPSTWYV"


# ---------------------------------------------------------------------------------------------------------------
# writers
# ---------------------------------------------------------------------------------------------------------------
def test_java_double_to_string():
    """java.lang.Double.toString: decimal for 1e-3 <= |x| < 1e7, otherwise computerised scientific notation"""
    from bnkit_b200.pog import java_double
    for x, want in [(0.0, "0.0"), (1.0, "1.0"), (0.5, "0.5"), (0.001, "0.001"), (0.00099, "9.9E-4"), (1e-4, "1.0E-4"),
                    (1.25e-5, "1.25E-5"), (9999999.0, "9999999.0"), (1e7, "1.0E7"), (1.5e10, "1.5E10"), (123.456, "123.456"),
                    (0.1, "0.1"), (100.0, "100.0"), (4.74e-4, "4.74E-4"), (-2.5, "-2.5"), (3.0e-300, "3.0E-300"),
                    (0.9999998, "0.9999998"), (float("nan"), "NaN"), (float("inf"), "Infinity")]:
        assert java_double(x) == want, (x, java_double(x), want)
    rng = np.random.default_rng(0)
    for x in np.concatenate([rng.random(200), rng.random(200) * 1e-6, np.exp(rng.normal(0, 30, 200))]):
        s = java_double(x)
        assert float(s.replace("E", "e")) == x          # round-trips (shortest digits)
        assert ("E" in s) == (not (1e-3 <= x < 1e7))


def test_fasta_and_tsv_layout():
    """FastaWriter.save(String[], Object[][]) (src/dat/file/FastaWriter.java:117-142) breaks BEFORE every 60th symbol;
    the DISTRIB table (src/asr/GRASP.java:599-630): Index + alphabet header, 1-based rows, empty cells where absent"""
    from bnkit_b200.phylo import EnumDistrib
    from bnkit_b200.pog import distrib_tsv_text, fasta_text
    seq = list("ACDEFGHIKL" * 13)
    seq[3] = None
    txt = fasta_text(["N0", "N1"], [seq, ["A", None]])
    lines = txt.split("\n")
    assert lines[0] == ">N0" and [len(x) for x in lines[1:4]] == [60, 60, 10] and lines[1][3] == "-"
    assert lines[4] == ">N1" and lines[5] == "A-" and txt.endswith("\n")
    d = [EnumDistrib("ACGT", np.array([0.25, 0.5, 0.125, 0.125])), None, EnumDistrib("ACGT", np.array([1e-5, 0.99999, 0, 0]))]
    tsv = distrib_tsv_text(d, "ACGT").split("\n")
    assert tsv[0] == "Index\tA\tC\tG\tT"
    assert tsv[1] == "1\t0.25\t0.5\t0.125\t0.125"
    assert tsv[2] == "2\t\t\t\t"
    assert tsv[3] == "3\t1.0E-5\t0.99999\t0.0\t0.0"


# ---------------------------------------------------------------------------------------------------------------
# POGraph: JSON of the documented ancestors, consensus search
# ---------------------------------------------------------------------------------------------------------------
def _canon(j):
    """a POGraph JSON with its edge list as a dict (the reference writes edges in HashMap order)"""
    edges = {tuple(k): (e["Forward"], e["Backward"], e["Recip"], float(e["Weight"])) for k, e in zip(j.get("Edgeindices", []), j.get("Edges", []))}
    return {"Size": j["Size"], "Indices": j["Indices"], "Adjacent": j["Adjacent"], "Starts": j["Starts"], "Ends": j["Ends"],
            "Nodes": j.get("Nodes"), "Nodetype": j.get("Nodetype"), "Name": j.get("Name"), "Directed": j["Directed"],
            "Terminated": j["Terminated"], "Datatype": j["Datatype"], "Edgetype": j.get("Edgetype"), "edges": edges}


def test_pograph_json_roundtrip_of_the_documented_ancestors():
    from bnkit_b200.pog import POGraph
    for doc in gold("c1_66.json")["doc_ancestors"]:
        g = POGraph.fromJSON(doc)
        assert g.getIndices() == [2, 4] and g.getStarts() == [2] and g.getEnds() == [4] and g.getForward(2) == [4]
        assert g.getEdge(2, 4).getReciprocated() and g.getEdge(-1, 2).forward and g.getEdge(4, 5).backward
        assert _canon(g.toJSON()) == _canon(doc)
        assert _canon(POGraph.fromJSON(json.dumps(g.toJSON())).toJSON()) == _canon(doc)
        assert g.getMostSupported() == [2, 4]


def _all_paths(g):
    """every start-to-end path of a POGraph with its DijkstraSearch cost (unreciprocated edges x 1000)"""
    out = []

    def cost(a, b):
        e = g.getEdge(a, b)
        return None if e is None else (e.weight if e.getReciprocated() else 1000 * e.weight)

    def walk(v, path, c):
        if g.isEndNode(v):
            ce = cost(v, g.maxsize())
            if ce is not None:
                out.append((c + ce, list(path)))
        for w in g.getForward(v):
            cw = cost(v, w)
            if cw is not None:
                path.append(w)
                walk(w, path, c + cw)
                path.pop()
    for s in g.getStarts():
        cs = cost(-1, s)
        if cs is not None:
            walk(s, [s], cs)
    return out


def _pograph_test_graph():
    """POGraphTest.setupPOG (test/dat/pog/POGraphTest.java:26-79): 30 nodes, chains of step 2, 7 and 9 from the start."""
    from bnkit_b200 import pog
    N = 30
    g = pog.POGraph(N)
    for i in range(N):
        g.addNode(i)
    touched = set()
    for step in (2, 7, 9):
        last, frm = -1, -1
        while frm < N - 10:
            to = min(frm + step, N)
            g.addEdge(frm, to)
            touched.add(to)
            last = to
            frm += step
        if last < N:
            g.addEdge(last, N)
    return g, touched, N


def test_pograph_primitives_against_the_reference_test_values():
    """The numbers POGraphTest asserts (:165-243, :288-297): three hops on the shortest way in either direction (an edge
    to the end terminal marks an end node, it is not a forward edge), every chained node visited, isPath."""
    from oracle import bep
    g, touched, N = _pograph_test_graph()

    def hops(idx, nrounds, nxt):
        n = nxt(idx)
        return nrounds if not n else min(hops(v, nrounds + 1, nxt) for v in n)

    def visit(idx, nxt, acc):
        for v in nxt(idx):
            if v not in acc:
                acc.add(v)
                visit(v, nxt, acc)
        return acc
    assert hops(-1, 0, g.getForward) == 3 and hops(N, 0, g.getBackward) == 3
    assert visit(-1, g.getForward, set()) >= touched - {N} and visit(N, g.getBackward, set()) >= touched - {N}
    assert g.isPath(6, 19) and g.isPath(7, 26) and g.isPath(7, 7) and g.isPath(-1, N)
    assert not g.isPath(8, 20) and not g.isPath(19, 26) and not g.isPath(26, 7) and not g.isPath(0, N)
    assert g.isContiguous()
    # the BEP restatement's graph answers the same questions the same way
    o = bep.POG(N)
    for i in range(N):
        o.add_node(i)
    for a in [-1] + list(range(N)):
        for b in g.getForward(a):
            o.add_edge(a, b)
    for a in g.getEnds():
        o.add_edge(a, N)
    assert o.is_contiguous()
    o.remove_node(20); o.remove_node(21); o.remove_node(26)   # all three ends gone: no way through
    assert not o.is_contiguous()


def test_most_supported_path_is_a_cheapest_path():
    """POGraph.getMostSupported = DijkstraSearch(-1 -> N) (src/dat/pog/GraphSearch.java:105-206): on random terminated
    DAGs the returned path is a start-to-end path whose cost equals the minimum over all paths (enumerated)."""
    from bnkit_b200.pog import BidirEdge, POGraph
    rng = np.random.default_rng(7)
    found = 0
    for trial in range(200):
        n = int(rng.integers(2, 9))
        g = POGraph(n)
        nodes = [i for i in range(n) if rng.random() < 0.8]
        for i in nodes:
            g.addNode(i)
        for a, b in itertools.combinations(nodes, 2):
            if rng.random() < 0.45:
                g.addEdge(a, b, BidirEdge(rng.random() < 0.7, rng.random() < 0.7, float(rng.integers(0, 6)) / 2))
        for i in nodes:
            if rng.random() < 0.35:
                g.addEdge(-1, i, BidirEdge(True, True, 0.0))
            if rng.random() < 0.35:
                g.addEdge(i, n, BidirEdge(rng.random() < 0.7, True, float(rng.integers(0, 4))))
        paths = _all_paths(g)
        got = g.getMostSupported()
        if not paths:
            assert got is None
            continue
        found += 1
        best = min(c for c, _ in paths)
        assert got in [p for c, p in paths if c == best], (trial, got, sorted(paths)[:3])
    assert found > 60


# ---------------------------------------------------------------------------------------------------------------
# the chain on the reference's data with the oracle's inference results (CPU)
# ---------------------------------------------------------------------------------------------------------------
def _oracle_prediction(name, oracle, model_name):
    """a Prediction over the oracle's BEP graphs, with the oracle's joint states and posteriors filled in"""
    from oracle import bep
    from bnkit_b200.phylo import EnumDistrib, IdxTree, POGraph, Prediction
    fx = gold(name)
    parent, rows, n_pos, obs = fixture_rows(fx)
    tree = IdxTree.fromJSON(fx["tree"])
    mask, anc = bep.presence_mask(parent, rows, n_pos)
    mask = np.asarray(mask, np.uint8)
    present = np.stack([oracle.prune(np.asarray(parent, np.int32), np.ascontiguousarray(mask[:, c]), True)[0] for c in range(n_pos)], axis=1)
    graphs = {}
    for v, pg in anc.items():
        fl = pg.edge_flags()
        es = sorted(fl)
        graphs[v] = POGraph.fromEdges(n_pos, [a for a, _ in es], [b for _, b in es], [fl[e] for e in es], name=Prediction._ancestor_id(tree, v))
    states = [[None] * n_pos for _ in range(tree.getSize())]
    for nm, seq in zip(fx["names"], fx["seqs"]):
        states[tree.getIndex(nm)] = [None if ch == "-" else ch for ch in seq]
    pred = Prediction(tree, states, present, ancestors=graphs)
    om = oracle.Model(model_name)
    dist = np.asarray(fx["tree"]["Distances"])
    st = oracle.fast_joint(om, np.asarray(parent, np.int32), dist, obs, present)
    pred.joint_states = [[None if s == NONE else AA[s] for s in row] for row in st]
    post, _ = oracle.fast_marginal(om, np.asarray(parent, np.int32), dist, obs, present)
    for v in tree.getAncestors():
        pred.distribs[v] = [None if np.isnan(post[v, c]).any() else EnumDistrib(AA, post[v, c]) for c in range(n_pos)]
    return fx, tree, pred


def test_documented_recon_ancestors_json_from_the_oracle_chain(oracle):
    """docs/json-api.md:113-122: the Recon job on data/66.* (JTT, joint, BEP) returns ancestors "0", "1", "2" as
    POGs with nodes {2, 4} = A, H, edges (-1,2), (2,4), (4,5) all inferred in both directions."""
    from bnkit_b200.phylo import Prediction
    fx, tree, pred = _oracle_prediction("c1_66.json", oracle, "JTT")
    pogs = pred.getAncestors(Prediction.JOINT)
    assert len(pogs) == len(tree.getAncestors())
    for doc in fx["doc_ancestors"]:
        assert _canon(pogs[doc["Name"]].toJSON()) == _canon(doc)
    out = pred.toJSON()
    assert out["Datatype"] == "Prediction" and len(out["Ancestors"]) == len(tree.getAncestors())
    for doc, got in zip(fx["doc_ancestors"], out["Ancestors"][:3]):     # "Ancestors" is in ancestor order
        assert _canon(got) == _canon(doc)
    assert pred.getSequence("0", Prediction.JOINT, True) == [None, None, "A", None, "H"]
    assert pred.getSequence("N1", Prediction.JOINT, False) == ["A", "H"]
    assert pred.getSequence(2, Prediction.MARGINAL, True) == [None, None, "A", None, "H"]
    fa = pred.ancestorsFasta(Prediction.JOINT).split("\n")
    assert fa[0] == ">N0" and fa[1] == "--A-H" and fa[2] == ">N1"
    assert len([x for x in fa if x.startswith(">")]) == len(tree.getAncestors())


@pytest.mark.parametrize("name", ["c1_66.json", "c2_3_2_1_1_filt.json"])
def test_asr_json_roundtrip_and_consensus_properties(oracle, tmp_path, name):
    """Prediction.save / load (ASR.json, src/asr/Prediction.java:133-268): tree, extants, ancestors' graphs with
    their values and edge flags survive; the mask rebuilt from the graphs equals the one the run used; every
    consensus path is a start-to-end path of the ancestor's graph through present columns."""
    from bnkit_b200.phylo import Prediction
    fx, tree, pred = _oracle_prediction(name, oracle, "JTT")
    pred.getAncestors(Prediction.JOINT)            # decorate
    f = tmp_path / "ASR.json"
    pred.save(str(f))
    back = Prediction.load(str(f))
    assert back.tree.toJSON() == tree.toJSON()
    assert np.array_equal(back.present, pred.present)
    assert back.states == pred.states
    for v in tree.getAncestors():
        assert _canon(back.ancarr[v].toJSON()) == _canon(pred.ancarr[v].toJSON())
        for c in back.ancarr[v].getIndices():
            assert back.ancarr[v].nodes[c] == pred.joint_states[v][c]
    assert json.loads(json.dumps(back.toJSON())) == json.loads(f.read_text())
    for v in tree.getAncestors():
        g = pred.ancarr[v]
        path = pred.getConsensus(v)
        assert path is not None and g.isStartNode(path[0]) and g.isEndNode(path[-1])
        assert all(g.isEdge(a, b) for a, b in zip(path, path[1:]))
        assert all(pred.present[v, c] for c in path)
        seq = pred.getSequence("N%d" % tree.getID(v), Prediction.JOINT, True)
        assert [c for c, x in enumerate(seq) if x is not None] == path
        # the consensus is a cheapest path under the weights getConsensus installed
        best = min(c for c, _ in _all_paths(g)) if g.size() <= 12 else None
        if best is not None:
            assert path in [p for c, p in _all_paths(g) if c == best]
    tsv = pred.distribTsv(tree.getID(tree.getAncestors()[0]), type("M", (), {"getDomain": lambda self: AA})()).split("\n")
    assert tsv[0].split("\t") == ["Index"] + list(AA) and len(tsv) == pred.n_cols + 2
    for c, line in enumerate(tsv[1:-1]):
        cells = line.split("\t")
        assert cells[0] == str(c + 1)
        if pred.present[tree.getAncestors()[0], c]:
            assert abs(sum(float(x) for x in cells[1:]) - 1) < 1e-12
        else:
            assert cells[1:] == [""] * 20


def test_from_json_rejects_malformed_input():
    from bnkit_b200.phylo import ASRRuntimeException, Prediction
    with pytest.raises(ASRRuntimeException):
        Prediction.fromJSON({"Datatype": "Dataset"})
    with pytest.raises(ASRRuntimeException):
        Prediction.fromJSON({"Datatype": "Prediction"})
    with pytest.raises(ASRRuntimeException):
        Prediction.fromJSON({"Datatype": "Prediction", "Input": {"Tree": {"Parents": [-1, 0, 0]}}})


# ---------------------------------------------------------------------------------------------------------------
# device chain
# ---------------------------------------------------------------------------------------------------------------
def _graphs_equal(got, want_pog, n_pos):
    fr, to, fl = got
    have = {(int(a), int(b)): int(f) for a, b, f in zip(fr, to, fl)}
    return have == want_pog.edge_flags()


@pytest.mark.gpu
def test_gpu_bep_infer_graphs_and_flags_equal_the_oracle(bnk):
    """bnk_bep_infer: every ancestor's edges and their direction flags (POGraph.BidirEdge) on C1, C2 and random
    alignments incl. multifurcations -- the patch rounds overwrite flags as IdxEdgeGraph.addEdge does"""
    from oracle import bep
    cases = []
    for name in ("c1_66.json", "c2_3_2_1_1_filt.json"):
        parent, rows, n_pos, obs = fixture_rows(gold(name))
        cases.append((np.asarray(parent, np.int32), rows, n_pos, obs))
    rng = np.random.default_rng(21)
    for trial in range(40):
        parent, _, rows, obs = random_case(rng, int(rng.integers(3, 40)), int(rng.integers(2, 30)),
                                           float(rng.choice([0.2, 0.5, 0.7])), max_children=int(rng.integers(2, 5)))
        cases.append((parent, rows, obs.shape[1], obs))
    patched = 0
    for parent, rows, n_pos, obs in cases:
        want_mask, anc = bep.presence_mask(list(map(int, parent)), rows, n_pos)
        tree = bnk.Tree(parent, np.full(len(parent), 0.1))
        mask, graphs = tree.bep_infer(obs)
        assert np.array_equal(mask, np.asarray(want_mask, np.uint8))
        assert set(graphs) == set(anc)
        for v in anc:
            assert _graphs_equal(graphs[v], anc[v], n_pos), v
            patched += any(f in (1, 2) for f in graphs[v][2])
    assert patched > 0


@pytest.mark.gpu
def test_gpu_documented_recon_json_and_outputs(bnk, tmp_path):
    """the documented Recon job end to end on the device: BEP -> orphan pruning -> joint under JTT -> ancestors'
    POG JSON equal to docs/json-api.md:117-119; ASR.json written, re-loaded, and re-inferred to the same states"""
    from bnkit_b200.phylo import IdxTree, Prediction, SubstModel
    fx = gold("c1_66.json")
    tree = IdxTree.fromJSON(fx["tree"])
    n_pos = len(fx["seqs"][0])
    states = [[None] * n_pos for _ in range(tree.getSize())]
    for nm, seq in zip(fx["names"], fx["seqs"]):
        states[tree.getIndex(nm)] = [None if ch == "-" else ch for ch in seq]
    model = SubstModel.createModel("JTT")
    pred = Prediction.PredictByBidirEdgeParsimony(tree, states)
    joint = pred.getJoint(model)
    pogs = pred.getAncestors(Prediction.JOINT)
    for doc in fx["doc_ancestors"]:
        assert _canon(pogs[doc["Name"]].toJSON()) == _canon(doc)
    assert pred.ancestorsFasta(Prediction.JOINT).split("\n")[:2] == [">N0", "--A-H"]
    pred.getMarginal(0, model)
    tsv = pred.distribTsv("N0", model).split("\n")
    assert tsv[1] == "1" + "\t" * 20 and abs(sum(float(x) for x in tsv[3].split("\t")[1:]) - 1) < 1e-12
    f = tmp_path / "ASR.json"
    pred.save(str(f))
    back = Prediction.load(str(f))
    assert np.array_equal(back.present, pred.present)
    assert back.getJoint(model) == joint
    pred.close()
    back.close()


@pytest.mark.gpu
def test_gpu_c2_outputs_equal_the_oracle_chain(bnk, oracle):
    """BASELINE configs[1] through the device chain: consensus sequences of all 37 ancestors (joint and marginal
    modes) equal those assembled from the oracle's graphs, states and posteriors"""
    from bnkit_b200.phylo import IdxTree, Prediction, SubstModel
    fx, tree, want = _oracle_prediction("c2_3_2_1_1_filt.json", oracle, "JTT")
    model = SubstModel.createModel("JTT")
    pred = Prediction.PredictByBidirEdgeParsimony(tree, want.states)
    assert np.array_equal(pred.present, want.present)
    pred.getJoint(model)
    for v in tree.getAncestors():
        pred.getMarginal(v, model)
    assert pred.ancestorsFasta(Prediction.JOINT) == want.ancestorsFasta(Prediction.JOINT)
    assert pred.ancestorsFasta(Prediction.MARGINAL, gappy=False) == want.ancestorsFasta(Prediction.MARGINAL, gappy=False)
    for v in tree.getAncestors():
        assert pred.getConsensus(v) == want.getConsensus(v)
    pred.close()


def test_clustal_and_newick_outputs_of_a_prediction(oracle):
    """asr.GRASP's other output formats over the same consensus sequences: CLUSTAL (AlnWriter) and the tree with
    ancestors labelled N<number> (Newick.save MODE_ANCESTOR)"""
    from bnkit_b200.files import loadAlignment
    from bnkit_b200.phylo import IdxTree, Prediction
    fx, tree, pred = _oracle_prediction("c1_66.json", oracle, "JTT")
    aln = pred.ancestorsClustal(Prediction.JOINT)
    names, rows = loadAlignment(aln, AA, is_text=True)
    assert names[:3] == ["N0", "N1", "N2"] and rows[0] == [None, None, "A", None, "H"]
    fa_names, fa_rows = loadAlignment(pred.ancestorsFasta(Prediction.JOINT), AA, is_text=True)
    assert (fa_names, fa_rows) == (names, rows)
    nwk = pred.treeNewick()
    back = IdxTree.fromNewick(nwk)
    assert np.array_equal(back.parent, tree.parent) and [back.getID(i) for i in back.getAncestors()] == [tree.getID(i) for i in tree.getAncestors()]
    assert np.allclose(back.distance[1:], tree.distance[1:], rtol=0, atol=0)
