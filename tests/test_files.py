"""Input parsers and writers as a library (SURVEY.md section 8 f-4 + the writers of f-2): bnkit_b200/files.py and
IdxTree.fromNewick against the reference's own test resources (tests/golden/reference_text_files.json, copied
verbatim by tools/make_golden_fixtures.py): test/dat/phylo/TreeTest.java:99-121 writes saveme.nwk / .anwk / .snwk
and default.anwk with Newick.save -- those files are the golden vectors of the Newick writer."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
AA = "ARNDCQEGHILKMFPSTWYV"


@pytest.fixture(scope="module")
def ref_files():
    with open(os.path.join(HERE, "golden", "reference_text_files.json")) as fh:
        return json.load(fh)


def test_format_sniffing(ref_files):
    from bnkit_b200.files import sniff_format
    assert sniff_format(ref_files["data/66.aln"]) == "FASTA"
    assert sniff_format(ref_files["test/resources/default.aln"]) == "CLUSTAL"
    assert sniff_format(ref_files["data/66.nwk"]) == "NEWICK"
    assert sniff_format("#NEXUS\nbegin trees;") == "NEXUS" and sniff_format("\n\nhello") is None and sniff_format("") is None


def test_fasta_reader_rules(ref_files):
    """FastaReader.loadGappy (src/dat/file/FastaReader.java:255-306, :362-405): name = first token of the defline,
    symbols upper-cased, '-' -> null, anything outside the alphabet is an error (X is not an ambiguity code)"""
    from bnkit_b200.files import ASRException, loadAlignment, parse_fasta_gappy
    names, rows = loadAlignment(ref_files["data/66.aln"], AA, is_text=True)
    fx = json.load(open(os.path.join(HERE, "golden", "c1_66.json")))
    assert names == fx["names"] and ["".join("-" if x is None else x for x in r) for r in rows] == fx["seqs"]
    got = parse_fasta_gappy(">a desc, more;x\nac-D\n\n  ef  \n>b\tz\n----\n", "ACDEF")
    assert got == [("a", ["A", "C", None, "D", "E", "F"]), ("b", [None] * 4)]
    lower = loadAlignment(ref_files["test/resources/basic_5.aln"], AA, is_text=True)    # holds lower-case residues
    assert len(lower[0]) == 5 and all(x is None or x in AA for r in lower[1] for x in r)
    with pytest.raises(ASRException) as e:
        loadAlignment(">s1\nACDX\n", AA, is_text=True)
    assert "invalid symbol" in str(e.value) and "Unrecognised symbol: X at index 3" in str(e.value)
    with pytest.raises(ASRException) as e:
        loadAlignment(ref_files["test/resources/input_test_aln_diff_length_4.aln"], AA, is_text=True)
    assert "different lengths" in str(e.value)
    with pytest.raises(ASRException):
        loadAlignment("hello world\n", AA, is_text=True)
    with pytest.raises(ASRException):
        loadAlignment(ref_files["data/66.nwk"], AA, is_text=True)          # a tree is not an alignment


def test_clustal_reader_rules(ref_files):
    """AlnReader.load (src/dat/file/AlnReader.java:60-196): blocks, identifiers with blanks, /start-end suffixes,
    conservation lines skipped, names matched by row when repeated; equals the FASTA reading of the same data"""
    from bnkit_b200.files import ASRException, clustal_text, loadAlignment, parse_clustal
    names, rows = loadAlignment(ref_files["test/resources/default.aln"], AA, is_text=True)
    assert names[:3] == ["s1", "s2", "s3"] and len(set(len(r) for r in rows)) == 1 and len(rows[0]) == 22
    assert "".join("-" if x is None else x for x in rows[0]) == "---R--SFLTKSVKQIKEGRL-"
    txt = ("CLUSTAL W (1.83) multiple sequence alignment\n\n\n"
           "P25101/1-20     ---METLCLR\nhomo sapiens 2  MQPPPSLCGR\nP30556          ----------\n                   :  .   \n\n"
           "P25101/1-20     ASFWLALV\nhomo sapiens 2  ALVALVLA\nP30556          --MILNSS\n")
    # (the sequence is the LONGEST token after the first, so names with blanks need chunks longer than their words)
    assert parse_clustal(txt, AA) == [("P25101", [None] * 3 + list("METLCLRASFWLALV")), ("homo sapiens 2", list("MQPPPSLCGRALVALVLA")),
                                      ("P30556", [None] * 12 + list("MILNSS"))]
    with pytest.raises(RuntimeError):
        parse_clustal("MUSCLE\n\na  ACD\n", AA)
    with pytest.raises(ASRException):
        loadAlignment("CLUSTAL\n\na  ACB\n", AA, is_text=True)
    # AlnWriter.save -> AlnReader.load round trip, 60 columns per block
    wide = [list(AA * 4), [None] * 80]
    out = clustal_text(["one", "two"], wide)
    assert out.startswith("CLUSTAL\n\none\t" + AA * 3 + "\ntwo\t" + "-" * 60 + "\n\n\none\t" + AA + "\n")
    assert loadAlignment(out, AA, is_text=True) == (["one", "two"], wide)


def test_check_data(ref_files):
    """Utils.checkData (src/dat/file/Utils.java:140-189)"""
    from bnkit_b200.files import ASRException, checkData, loadAlignment
    from bnkit_b200.phylo import IdxTree
    tree = IdxTree.fromNewick(ref_files["data/66.nwk"])
    names, _ = loadAlignment(ref_files["data/66.aln"], AA, is_text=True)
    checkData(names, tree)
    with pytest.raises(ASRException) as e:
        checkData(names[:-1] + [names[0], "stranger"], tree)
    msg = str(e.value)
    assert "Duplicate identifiers in alignment" in msg and names[0] in msg
    assert "missing in tree stranger" in msg and "missing in alignment " + names[-1] in msg
    with pytest.raises(ASRException) as e:
        checkData(["a", "b"], IdxTree.fromNewick("((a:1,a:1):1,b:1);"))
    assert "Duplicate identifiers in tree" in str(e.value)


def test_newick_reader_quotes_blocks_and_writer_golden(ref_files):
    """Newick.parse keeps quoted labels verbatim and takes label="..." out of [..] blocks (:170-203); Newick.save in
    its three modes against the files the reference's TreeTest wrote (saveme.nwk / .anwk / .snwk, default.anwk)"""
    from bnkit_b200.files import NEWICK_ANCESTOR, NEWICK_DEFAULT, NEWICK_STRIPPED, newick_text
    from bnkit_b200.phylo import IdxTree
    # saveme.nwk holds the ancestor label "N_42": Newick.parse reads N<...> as an ancestor number and fails on it
    # (Integer.parseInt, src/dat/file/Newick.java:224-235) -- so does the mirror
    with pytest.raises(RuntimeError) as e:
        IdxTree.fromNewick(ref_files["test/resources/saveme.nwk"])
    assert "Invalid ancestor name: N_42" in str(e.value)
    # the same tree from its ancestor-mode file (labels N0..N3), relabelled through the arrays as TreeTest's ss3 has it
    t = IdxTree.fromNewick(ref_files["test/resources/saveme.anwk"])
    assert "D@ --" in t.labels and t.getIndex("N2") > 0 and [t.getID(i) for i in t.getAncestors()] == [0, 1, 2, 3]
    assert newick_text(t, NEWICK_DEFAULT) == newick_text(t, NEWICK_ANCESTOR) == ref_files["test/resources/saveme.anwk"].strip() + "\n"
    t2 = IdxTree(t.parent, t.distance, [{"N1": "N_42", "N3": "_mysubtree is cool"}.get(lab, lab) for lab in t.labels])
    t2.has_distance = list(t.has_distance)
    assert newick_text(t2, NEWICK_DEFAULT) == ref_files["test/resources/saveme.nwk"].strip() + "\n"
    assert newick_text(t2, NEWICK_ANCESTOR) == ref_files["test/resources/saveme.anwk"].strip() + "\n"
    assert newick_text(t2, NEWICK_STRIPPED) == ref_files["test/resources/saveme.snwk"].strip() + "\n"
    blk = IdxTree.fromNewick("((E:0.1,'D@ --':0.5)[&label=\"N7\"]:0.2,(C:0.4,(B:0.4,A:0.3)'_my sub':0.1):0.2)")
    assert blk.labels[1] == "N7" and blk.getID(1) == 7 and blk.getIndex("_my sub") > 0
    with pytest.raises(RuntimeError):
        IdxTree.fromNewick("((E:0.1,'D:0.5):0.2,C:1);")
    with pytest.raises(RuntimeError):
        newick_text(IdxTree([-1, 0, 0], [0, 1, 1], ["r", "a(b", "c"]), NEWICK_STRIPPED)
    # default.nwk -> default.anwk: bootstrap values as ancestor labels become N<k>, distances print as Double.toString
    d = IdxTree.fromNewick(ref_files["test/resources/default.nwk"])
    assert newick_text(d, NEWICK_ANCESTOR) == ref_files["test/resources/default.anwk"].strip() + "\n"
    back = IdxTree.fromNewick(newick_text(d, NEWICK_ANCESTOR))
    assert np.array_equal(back.parent, d.parent) and np.array_equal(back.distance, d.distance)
    assert [back.getID(i) for i in back.getAncestors()] == [d.getID(i) for i in d.getAncestors()]
    # a 3,000-level caterpillar does not hit the recursion limit in the writer
    from bnkit_b200 import synth
    parent, dist = synth.caterpillar_tree(3000, seed=1)
    deep = IdxTree(parent, dist, ["n%d" % i for i in range(len(parent))])
    assert IdxTree.fromNewick(newick_text(deep)).getSize() == len(parent)


def test_empirical_frequencies():
    """TSVFile.loadEmpiricalFreqFile (src/dat/file/TSVFile.java:920-981) + GRASP's renormalisation (GRASP.java:385-401)"""
    from bnkit_b200.files import load_empirical_freqs
    rng = np.random.default_rng(2)
    f = rng.random(20)
    order = list(AA)
    rng.shuffle(order)
    txt = "Character\tCount\tProportion\n" + "".join("%s\t7\t%r\n" % (ch, float(f[AA.index(ch)])) for ch in order)
    got = load_empirical_freqs(txt, "LG", is_text=True)
    total = 0.0
    for x in f:
        total += x
    assert got == [x / total for x in f]                  # renormalised: the sum is not 1
    g = f / f.sum()
    txt = "Proportion\tCharacter\n" + "".join("%r\t%s\n" % (float(g[i]), AA[i]) for i in range(20))
    assert load_empirical_freqs(txt, "JTT", is_text=True) == [float(x) for x in g]     # within 1e-5 of 1: untouched
    with pytest.raises(RuntimeError):
        load_empirical_freqs("Char\tProportion\nA\t1\n", "LG", is_text=True)
    with pytest.raises(RuntimeError):
        load_empirical_freqs("Character\tProportion\nA\t1\n", "LG", is_text=True)          # 1 entry, 20 states
    with pytest.raises(TypeError):
        load_empirical_freqs("Character\tProportion\n" + "".join("%s\t0.05\n" % ch for ch in AA[:-1] + "X"), "LG", is_text=True)
    with pytest.raises(ValueError):
        load_empirical_freqs("Character\tProportion\n" + "".join("%s\tabc\n" % ch for ch in "ACGT"), "Yang", is_text=True)
    with pytest.raises(ValueError):
        load_empirical_freqs("Character\tProportion\nA\t1\n", "JC", is_text=True)


@pytest.mark.gpu
def test_gpu_empirical_frequencies_reach_the_model(bnk, oracle):
    """asr.GRASP -ef: SubstModel.createModel(name, EMPIRICAL_FREQS) (GRASP.java:413-415) -- the frequencies change R,
    the eigensystem and the root prior; joint states and log-likelihoods equal the oracle's under the same model"""
    from bnkit_b200.files import load_empirical_freqs
    from helpers import random_obs, random_tree
    rng = np.random.default_rng(6)
    f = rng.dirichlet(np.full(20, 3.0)) * 1.3             # does not sum to 1: renormalised by the loader
    txt = "Character\tProportion\n" + "".join("%s\t%r\n" % (AA[i], float(f[i])) for i in range(20))
    F = load_empirical_freqs(txt, "WAG", is_text=True)
    m, om = bnk.Model("WAG", F=F), oracle.Model("WAG", F=F)
    assert np.array_equal(m.parts()["R"], om.parts()["R"]) and not np.array_equal(m.parts()["R"], bnk.Model("WAG").parts()["R"])
    parent, dist = random_tree(rng, 30)
    obs = random_obs(rng, parent, 40, 20)
    ws = bnk.Workspace(m, bnk.Tree(parent, dist), 40)
    assert np.array_equal(ws.joint(obs), oracle.fast_joint(om, parent, dist, obs))
    _, wl = oracle.fast_marginal(om, parent, dist, obs, want_post=False)
    assert np.allclose(ws.loglik(obs)[0], wl, rtol=1e-9)
    ws.close()
