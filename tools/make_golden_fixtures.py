#!/usr/bin/env python3
"""Turn the reference's data files and pinned test expectations for the hot path into small
JSON fixtures under tests/golden/.  Run in the build container only (needs /root/reference);
the outputs are committed, nothing reads /root/reference at test time.

Sources:
  data/66.aln, data/66.nwk                       BASELINE.json configs[0] (C1)
  data/3_2_1_1_filt.aln, data/3_2_1_1_filt.nwk   BASELINE.json configs[1] (C2)
  docs/json-api.md:97-122                        the same tree as IdxTree JSON ("Parents") + the documented
                                                 Recon result (JTT, joint, BEP): N0..N2 present at columns
                                                 {2,4} with states {A,H}
  test/resources/default.nwk + test/dat/phylo/IdxTreeTest.java:15-56   pinned prune sizes / root counts
"""
import json, pathlib, re, sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bnkit_b200.phylo import IdxTree, load_alignment  # parser only; no GPU needed

REF = pathlib.Path("/root/reference")
OUT = ROOT / "tests" / "golden"


def dataset(stem):
    tree = IdxTree.fromNewick((REF / "data" / (stem + ".nwk")).read_text())
    names, rows = load_alignment(str(REF / "data" / (stem + ".aln")))
    return {"tree": tree.toJSON(), "names": names, "seqs": ["".join("-" if c is None else c for c in r) for r in rows]}


c1 = dataset("66")
doc = (REF / "docs" / "json-api.md").read_text()
m = re.search(r'"Parents":\[([-\d,]+)\]', doc)
c1["doc_parents"] = [int(x) for x in m.group(1).split(",")]
m = re.search(r'"Labels":\[([^\]]+)\]', doc)
c1["doc_labels"] = json.loads("[" + m.group(1) + "]")
c1["doc_result"] = {"model": "JTT", "mode": "joint", "indel": "BEP", "ancestors": ["0", "1", "2"],
                    "indices": [2, 4], "values": ["A", "H"], "source": "docs/json-api.md:113-122"}
# the three ancestor POGs the documentation prints in full (docs/json-api.md:117-119), verbatim
c1["doc_ancestors"] = [json.loads(x) for x in re.findall(r'^\s+(\{ "Adjacent".*"Edgetype":"class dat.pog.POGraph\$BidirEdge" \}),\s*$', doc, re.M)]
assert len(c1["doc_ancestors"]) == 3
(OUT / "c1_66.json").write_text(json.dumps(c1))
(OUT / "c2_3_2_1_1_filt.json").write_text(json.dumps(dataset("3_2_1_1_filt")))

t = IdxTree.fromNewick((REF / "test" / "resources" / "default.nwk").read_text())
cases = [  # (pruned indices, size delta, roots without orphan removal, size equal with orphan removal?)
    ([1, 4], 2, 5, None), ([0], 1, 2, None), ([3], 1, 1, None), ([7, 8, 9], 3, 1, None), ([8, 9], 2, 1, None),
    ([7], 1, 3, True), ([2, 5, 7, 8, 13], 5, 6, True), ([0, 4, 7, 9, 10, 15], 6, 7, True)]
fix = {"source": "test/dat/phylo/IdxTreeTest.java:15-56 on test/resources/default.nwk",
       "parents": t.toJSON()["Parents"], "cases": [{"prune": c[0], "removed": c[1], "roots": c[2], "same_size_with_orphan_removal": c[3]} for c in cases]}
(OUT / "idxtree_prune.json").write_text(json.dumps(fix))
# small text files of the reference's own test resources, verbatim: golden vectors for the readers / writers
# (test/dat/phylo/TreeTest.java:99-121 writes saveme.* and default.anwk with Newick.save)
files = {}
for rel in ("test/resources/saveme.nwk", "test/resources/saveme.anwk", "test/resources/saveme.snwk",
            "test/resources/default.nwk", "test/resources/default.anwk", "test/resources/default.aln",
            "test/resources/input_test_aln_diff_length_4.aln", "test/resources/input_test_aln_diff_length_4.nwk",
            "test/resources/basic_5.aln"):
    files[rel] = (REF / rel).read_text()
files["data/66.aln"] = (REF / "data" / "66.aln").read_text()
files["data/66.nwk"] = (REF / "data" / "66.nwk").read_text()
(OUT / "reference_text_files.json").write_text(json.dumps(files))
print("wrote", [p.name for p in OUT.iterdir()])
