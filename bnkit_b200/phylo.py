"""Host-side mirror of the reference's interface for the hot path, over libbnkgpu.

The reference is Java; this image has no JDK, so (as the task allows) the host side above
the C ABI is written in Python with the reference's names, argument meaning and error
behaviour for exactly this path:

    bn.ctmc.SubstModel.createModel / getDomain / getProbs   src/bn/ctmc/SubstModel.java:303-385
    dat.phylo.IdxTree (parents / distances / labels, JSON)   src/dat/phylo/IdxTree.java:27-32, :277-351, :411-503
    dat.file.Newick.parse (0.0 -> 1e-5)                      src/dat/file/Newick.java:47-131, :170-257
    dat.phylo.TreeInstance                                   src/dat/phylo/TreeInstance.java:21-57
    asr.MaxLhoodJoint / asr.MaxLhoodMarginal (TreeDecor<E>)  src/asr/MaxLhoodJoint.java:39-170, src/asr/MaxLhoodMarginal.java:39-108
    asr.Prediction.getJoint / getMarginal                    src/asr/Prediction.java:555-649

The per-column decorators exist for API parity (and are what the parity tests drive, like the
reference's own MaxLhoodJointTest / MaxLhoodMarginalTest); the batch entry points of
Prediction are the drop-in that matters: one call, every column, on the GPU.
Everything numeric happens in libbnkgpu.so; nothing here falls back to the CPU.
"""
import json

import numpy as np

from . import lib
from . import pog as _pog
from .pog import ASRRuntimeException, POGraph, POGTree  # noqa: F401  (re-exported: the mirror's public names)

ALPHABETS = {
    "LG": "ARNDCQEGHILKMFPSTWYV", "JTT": "ARNDCQEGHILKMFPSTWYV", "WAG": "ARNDCQEGHILKMFPSTWYV",
    "Dayhoff": "ARNDCQEGHILKMFPSTWYV", "JC": "ACGT", "Yang": "ACGT",
}
NONE = lib.NONE


class EnumDistrib:
    """bn.prob.EnumDistrib: a normalised distribution over the model's domain (index order = model alphabet)."""

    def __init__(self, domain, probs):
        self.domain = domain
        self.distrib = np.asarray(probs, dtype=np.float64)

    def get(self, key):
        return float(self.distrib[key if isinstance(key, (int, np.integer)) else self.domain.index(key)])

    def getMaxIndex(self):  # EnumDistrib.java:274-281: first strict maximum, ascending
        mi = 0
        for i in range(1, len(self.distrib)):
            if self.distrib[mi] < self.distrib[i]:
                mi = i
        return mi

    def getMax(self):
        return self.domain[self.getMaxIndex()]

    _PREDEF = {"DNA": "ACGT", "RNA": "ACGU", "DNA with N": "ACGTN", "RNA with N": "ACGUN", "Protein": "ACDEFGHIKLMNPQRSTVWY",
               "Protein with X": "ACDEFGHIKLMNPQRSTVWYX", "Protein with gap": "ACDEFGHIKLMNPQRSTVWY-"}

    def toJSON(self):
        """EnumDistrib.toJSON (src/bn/prob/EnumDistrib.java:165-170) with Enumerable.toJSON (src/dat/Enumerable.java:271-283):
        a predefined domain goes by name, any other one by its values"""
        dom = "".join(self.domain) if all(isinstance(c, str) and len(c) == 1 for c in self.domain) else None
        name = next((k for k, v in self._PREDEF.items() if v == dom), None)
        domain = {"Predef": name} if name else {"Size": len(self.domain), "Values": list(self.domain),
                                                "Datatype": "Character" if dom is not None else "Object"}
        return {"Domain": domain, "Pr": [float(p) for p in self.distrib]}

    @staticmethod
    def fromJSON(obj, domain=None):
        """EnumDistrib.fromJSON (:180-193)"""
        if domain is None:
            d = obj["Domain"]
            if d.get("Predef") is not None:
                if d["Predef"] not in EnumDistrib._PREDEF:
                    raise ASRRuntimeException("Invalid predefined domain name: \"%s\"" % d["Predef"])
                domain = list(EnumDistrib._PREDEF[d["Predef"]])
            else:
                domain = list(d["Values"])
        pr = [float(x) for x in obj["Pr"]]
        if len(pr) != len(domain):
            raise ASRRuntimeException("EnumDistrib: %d probabilities for a domain of %d" % (len(pr), len(domain)))
        return EnumDistrib(domain, pr)

    def __repr__(self):
        return "<" + " ".join("%s:%.3f" % (s, p) for s, p in zip(self.domain, self.distrib)) + ">"


class SubstModel:
    """bn.ctmc.SubstModel facade; the 20x20 arithmetic lives in libbnkgpu (host eigensystem, device P(t))."""

    def __init__(self, handle, name, domain):
        self.handle, self.name, self.domain = handle, name, list(domain)

    @staticmethod
    def createModel(name, empiricalFreqs=None):
        if name not in ALPHABETS:
            return None  # SubstModel.createModel returns null for unknown names (SubstModel.java:334-338)
        return SubstModel(lib.Model(name, F=empiricalFreqs), name, ALPHABETS[name])

    @staticmethod
    def fromEigen(F, evec, ievec, evals, domain, name="custom"):
        """What the JNI stub does with the JVM's own SubstModel.getRexp()."""
        return SubstModel(lib.Model(eigen=(len(F), F, evec, ievec, evals)), name, domain)

    def getName(self):
        return self.name

    def getDomain(self):
        return self.domain

    def getProbs(self, t):
        """P[i][j] = P(child = j | parent = i, t), computed on the device (SubstModel.java:303-327)."""
        ws = lib.Workspace(self.handle, lib.Tree([-1], [0.0]), 1)
        try:
            return ws.probs([t], want_log=False)[0]
        finally:
            ws.close()


class IdxTree:
    """dat.phylo.IdxTree: nodes indexed depth-first, parent index < child index."""

    def __init__(self, parents, distances=None, labels=None, ancids=None):
        self.parent = np.asarray(parents, dtype=np.int32)
        n = len(self.parent)
        self.distance = np.zeros(n) if distances is None else np.asarray(distances, dtype=np.float64).copy()
        self.labels = list(labels) if labels is not None else [str(i) for i in range(n)]
        self.children = [[] for _ in range(n)]
        for i, p in enumerate(self.parent):
            if p >= 0:
                self.children[p].append(i)
        # BranchPoint.getID (src/dat/phylo/BranchPoint.java:88-90): a leaf's ID is its label, an ancestor's its ancestor
        # number.  Default = what IdxTree.fromJSON assigns (:312-321): 0, 1, ... over the ancestors in index order.
        if ancids is None:
            ancids = {v: k for k, v in enumerate(i for i in range(n) if self.children[i])}
        self.ancid = dict(ancids)
        # which branch points carry a distance (BranchPoint.getDistance throws for the others, so Newick.sprint
        # leaves ':d' out): all of them when built from arrays / JSON; from Newick text, those that had one
        self.has_distance = [True] * n
        self._handle = None

    # ---- construction ----
    @staticmethod
    def fromJSON(obj):
        """{"Parents": [...], "Labels": [...], "Distances": [...], "Branchpoints": n} (IdxTree.java:277-351)"""
        if isinstance(obj, str):
            obj = json.loads(obj)
        return IdxTree(obj["Parents"], obj.get("Distances"), obj.get("Labels"))

    def toJSON(self):
        return {"Parents": [int(p) for p in self.parent], "Labels": list(self.labels),
                "Distances": [float(d) for d in self.distance], "Branchpoints": self.getSize()}

    @staticmethod
    def fromNewick(text):
        """Newick -> IdxTree.  Ancestors are named N0, N1, ... in depth-first order unless labelled;
        a distance of exactly 0.0 becomes 1e-5 (Newick.java:54-56, :103-105)."""
        # white space is dropped outside quoted labels and [..] blocks (Newick.parse keeps those verbatim, :170-203)
        buf, q = [], None
        for ch in text:
            if q is None:
                if ch in "'[":
                    q = "'" if ch == "'" else "]"
                    buf.append(ch)
                elif not ch.isspace():
                    buf.append(ch)
            else:
                buf.append(ch)
                if ch == q:
                    q = None
        if q is not None:
            raise RuntimeError("Non-matched quotes, about here: " + "".join(buf)[-20:] if q == "'" else "Non-matched block, about here: " + "".join(buf)[-20:])
        s = "".join(buf)
        if s.endswith(";"):
            s = s[:-1]
        parents, dists, labels = [], [], []
        named, with_dist = {}, set()
        pos = 0
        anc = [0]

        def node(par):
            nonlocal pos
            idx = len(parents)
            parents.append(par)
            dists.append(0.0)
            labels.append(None)
            internal = False
            if pos < len(s) and s[pos] == "(":
                internal = True
                my_anc = anc[0]
                anc[0] += 1
                pos += 1
                while True:
                    node(idx)
                    if s[pos] == ",":
                        pos += 1
                        continue
                    if s[pos] == ")":
                        pos += 1
                        break
                    raise ValueError("Newick: expected ',' or ')' at %d" % pos)
            name = ""
            while pos < len(s) and s[pos] not in ",():":
                if s[pos] == "'":                      # a quoted label: taken verbatim without the quotes
                    end = s.index("'", pos + 1)
                    name += s[pos + 1:end]
                    pos = end + 1
                elif s[pos] == "[":                    # a block: label="..." inside it, else its content (:186-194)
                    end = s.index("]", pos + 1)
                    blk = s[pos:end + 1]
                    k = blk.find('label="')
                    if k >= 0:
                        name = blk[k + 7:k + 7 + blk[k + 7:].index('"')]
                    elif not name:
                        name = blk[1:-1]
                    pos = end + 1
                else:
                    name += s[pos]
                    pos += 1
            if pos < len(s) and s[pos] == ":":
                pos += 1
                start = pos
                while pos < len(s) and s[pos] not in ",()":
                    pos += 1
                try:
                    d = float(s[start:pos])
                except ValueError:
                    raise RuntimeError("A distance value couldn't be parsed as a number. The value is \"%s\"" % s[start:pos])
                dists[idx] = 0.00001 if d == 0.0 else d
                with_dist.add(idx)
            labels[idx] = name if name else ("N%d" % my_anc if internal else "")
            if internal:
                named[idx] = bool(name)
            return idx

        import sys
        old = sys.getrecursionlimit()
        sys.setrecursionlimit(max(old, 100000))
        try:
            node(-1)
        finally:
            sys.setrecursionlimit(old)
        # ancestor numbers as Newick.parse assigns them (src/dat/file/Newick.java:217-254): a label N<number> fixes
        # the number; every other ancestor takes the lowest numbers still free, in depth-first order
        ancids, taken, rest = {}, set(), []
        for idx in sorted(named):
            lab = labels[idx]
            if named[idx] and lab.startswith("N"):
                try:
                    k = int(lab[1:])
                except ValueError:
                    raise RuntimeError("Invalid ancestor name: " + lab + ". Needs to follow N<number> format.")
                if k in taken:
                    raise RuntimeError("Duplicate ancestor name/number: " + lab)
                taken.add(k)
                ancids[idx] = k
            else:
                rest.append(idx)
        count = 0
        for idx in rest:
            while count in taken:
                count += 1
            ancids[idx] = count
            taken.add(count)
            count += 1
        t = IdxTree(parents, dists, labels, ancids)
        t.has_distance = [i in with_dist for i in range(len(parents))]
        return t

    # ---- accessors (IdxTree.java) ----
    def getSize(self):
        return len(self.parent)

    def getParent(self, idx):
        return int(self.parent[idx])

    def getChildren(self, idx):
        return list(self.children[idx])

    def isLeaf(self, idx):
        return len(self.children[idx]) == 0

    def getDistance(self, idx):
        return float(self.distance[idx])

    def getLabel(self, idx):
        return self.labels[idx]

    def getID(self, idx):
        """BranchPoint.getID: the label of a leaf, the ancestor number (int) of an ancestor -- what the reference's
        IdxTree.getLabel(idx) returns (src/dat/phylo/IdxTree.java:759-763)"""
        return self.ancid[idx] if self.children[idx] else self.labels[idx]

    def getIndex(self, label):
        """IdxTree.getIndex (:997-1022): a string matches a branch point's label or the text of its ID, then
        "N<k>" the ancestor numbered k; an int matches an ancestor number; -1 if there is none"""
        if isinstance(label, str):
            for i, l in enumerate(self.labels):
                if l == label or str(self.getID(i)) == label:
                    return i
            if label.startswith("N"):
                try:
                    return self.getIndex(int(label[1:]))
                except ValueError:
                    pass
            return -1
        for i, k in self.ancid.items():
            if k == label:
                return i
        return -1

    def getLeaves(self):
        return [i for i in range(self.getSize()) if self.isLeaf(i)]

    def getAncestors(self):
        return [i for i in range(self.getSize()) if not self.isLeaf(i)]

    def isConnected(self, idx):  # IdxTree.java:746-751
        return len(self.children[idx]) > 0 or self.parent[idx] != -1

    def handle(self):
        if self._handle is None:
            self._handle = lib.Tree(self.parent, self.distance)
        return self._handle

    def getPrunedMask(self, pruneMe, removeOrphans=True):
        """getPrunedIndex + createPrunedTree (IdxTree.java:411-503) as a presence mask over the
        ORIGINAL indices: mask[i] = 1 iff node i survives (the GPU path keeps one shared topology)."""
        present = np.ones((self.getSize(), 1), dtype=np.uint8)
        for i in pruneMe:
            present[i, 0] = 0
        if removeOrphans:
            present = self.handle().prune_orphans(present)
        return present[:, 0]

    def getPrunedIndex(self, pruneMe, removeOrphans=True):
        mask = self.getPrunedMask(pruneMe, removeOrphans)
        match, k = [], 0
        for i in range(self.getSize()):
            if mask[i]:
                match.append(k)
                k += 1
            else:
                match.append(-1)
        return match

    @staticmethod
    def createPrunedTree(source, prunedIndex):
        keep = [i for i, m in enumerate(prunedIndex) if m >= 0]
        parents = []
        for i in keep:
            p = source.parent[i]
            parents.append(prunedIndex[p] if p >= 0 else -1)
        return IdxTree(parents, [source.distance[i] for i in keep], [source.labels[i] for i in keep])

    def getNRoots(self):
        return int((self.parent < 0).sum())


class TreeInstance:
    """dat.phylo.TreeInstance: one value (or None) per branch point of a tree."""

    def __init__(self, tree, assign):
        self.tree = tree
        if isinstance(assign, dict):
            arr = [None] * tree.getSize()
            for k, v in assign.items():
                arr[k] = v
            assign = arr
        self.instance = list(assign)

    def getInstance(self, idx):
        return self.instance[idx]

    def getSize(self):
        return len(self.instance)

    def getTree(self):
        return self.tree

    def _possible(self):
        out = []
        for v in self.instance:          # the reference collects them in a HashSet (TreeInstance.java:46-58): any order
            if v is not None and v not in out:
                out.append(v)
        return out

    def encode(self, recodeNullAtLeaves=True):
        """TreeInstance.encode (src/dat/phylo/TreeInstance.java:141-165): values become 0 .. k - 1, a null at a leaf
        becomes k when recodeNullAtLeaves; returns the key decode() takes.  In place, as in the reference."""
        possible = self._possible()
        key = list(possible) + ([None] if recodeNullAtLeaves else [])
        for i, v in enumerate(self.instance):
            if v is not None:
                self.instance[i] = possible.index(v)
            elif recodeNullAtLeaves and self.tree.isLeaf(i):
                self.instance[i] = len(possible)
        return key

    def decode(self, key):
        """TreeInstance.decode (:191-206): back to the symbols of the key (the null symbol back to None)"""
        if len(key) < 1:
            return
        for i, v in enumerate(self.instance):
            if v is not None:
                self.instance[i] = key[int(v)]


def _encode(values, domain):
    out = np.full(len(values), NONE, dtype=np.uint8)
    for i, v in enumerate(values):
        if v is not None:
            try:
                out[i] = domain.index(v)
            except ValueError:
                raise ASRRuntimeException("Value %r is not in the model's domain" % (v,))
    return out


class MaxLhoodJoint:
    """asr.MaxLhoodJoint: joint ML reconstruction of every uninstantiated, connected branch point."""

    def __init__(self, tree, model, rate=1.0):
        self.tree, self.model, self.rate = tree, model, float(rate)
        self.values = None

    def decorate(self, ti):
        obs = _encode(ti.instance, self.model.domain)[:, None]
        ws = lib.Workspace(self.model.handle, self.tree.handle(), 1)
        try:
            ws.set_rates(1, [self.rate])
            st = ws.joint(obs)[:, 0]
        finally:
            ws.close()
        self.values = [None if s == NONE else self.model.domain[s] for s in st]

    def getDecoration(self, idx):
        return self.values[idx]


class MaxLhoodMarginal:
    """asr.MaxLhoodMarginal: posterior distribution at one branch point."""

    def __init__(self, bpidx, tree, model, rate=1.0):
        self.bpidx, self.tree, self.model, self.rate = int(bpidx), tree, model, float(rate)
        self.value = None

    def decorate(self, ti):
        obs = _encode(ti.instance, self.model.domain)[:, None]
        ws = lib.Workspace(self.model.handle, self.tree.handle(), 1)
        try:
            ws.set_rates(1, [self.rate])
            post, _ = ws.marginal(obs, query=[self.bpidx], want_logl=False)
        finally:
            ws.close()
        p = post[0, 0]
        self.value = None if np.isnan(p).any() else EnumDistrib(self.model.domain, p)

    def getDecoration(self, idx):
        if idx == self.bpidx:
            return self.value
        raise ASRRuntimeException("Invalid ancestor index, not defined for marginal inference: %d" % idx)


class Prediction:
    """The character-inference half of asr.Prediction: getJoint / getMarginal over all columns.

    states:  [n_nodes][n_cols] characters (or None) at the extant leaves, as POGTree.getNodeInstance
             (src/dat/pog/POGTree.java:419-436) yields per column;
    present: [n_nodes][n_cols] 0/1, the per-column pruned trees of Prediction.getTree
             (src/asr/Prediction.java:329-402) as a mask, or None when every node is present.
    """

    JOINT, MARGINAL = "JOINT", "MARGINAL"   # GRASP.Inference

    def __init__(self, tree, states, present=None, device=-1, ancestors=None):
        self.tree = tree
        self.states = states
        self.present = None if present is None else np.ascontiguousarray(present, dtype=np.uint8)
        self.device = device
        self.n_cols = len(states[0]) if len(states) else 0
        self._ws = None
        self._ws_model = None
        # output side (section 8 f-2): the ancestors' partial-order graphs {branch point index: POGraph}, the extants'
        # graphs, and what the last getJoint / getMarginal calls inferred
        self.ancarr = ancestors
        self._pogTree = None              # the extants' graphs, built when the output side first needs them
        self.joint_states = None          # Object[][] states[bpidx][pos]
        self.distribs = {}                # bpidx -> EnumDistrib[] over positions

    @staticmethod
    def PredictByBidirEdgeParsimony(tree, states, device=-1):
        """asr.Prediction.PredictByBidirEdgeParsimony (src/asr/Prediction.java:1443-1524): indel inference by
        bi-directional edge parsimony; the returned Prediction carries the per-column trees Prediction.getTree
        (:329-402) would build -- ancestors present where their partial-order graph has the column, then
        (GRASP.REMOVE_INDEL_ORPHANS, src/asr/GRASP.java:42) subtrees without an extant removed."""
        n = tree.getSize()
        n_cols = len(states[0]) if len(states) else 0
        occ = np.full((n, n_cols), NONE, dtype=np.uint8)
        for v in range(n):
            if tree.isLeaf(v):
                for c in range(n_cols):
                    if states[v][c] is not None:
                        occ[v, c] = 0
        h = tree.handle()
        raw, graphs = h.bep_infer(occ, device=device)
        present = h.prune_orphans(raw)
        anc = {v: POGraph.fromEdges(n_cols, fr, to, fl, name=Prediction._ancestor_id(tree, v)) for v, (fr, to, fl) in graphs.items()}
        return Prediction(tree, states, present, device, ancestors=anc)

    @property
    def pogTree(self):
        if self._pogTree is None:
            self._pogTree = POGTree.fromAlignment(self.tree, self.states)
        return self._pogTree

    @pogTree.setter
    def pogTree(self, value):
        self._pogTree = value

    @staticmethod
    def _ancestor_id(tree, idx):
        """the ancestor's ID as text: BranchPoint IDs of ancestors are integers (POGraph names, FASTA names N<id>)"""
        return str(tree.getID(idx))

    def getBranchpointIndex(self, ancID):
        """index of the branch point with this ID in the tree, -1 if there is none (Prediction.java:453-456)"""
        if isinstance(ancID, (int, np.integer)):
            return self.tree.getIndex(int(ancID))
        i = self.tree.getIndex(str(ancID))
        if i == -1 and str(ancID).isdigit():
            i = self.tree.getIndex(int(ancID))
        return i

    def getAncestorIndices(self):
        return self.tree.getAncestors()

    def getPositions(self):
        return self.n_cols

    def _obs(self, model):
        n = self.tree.getSize()
        obs = np.full((n, self.n_cols), NONE, dtype=np.uint8)
        lut = {c: i for i, c in enumerate(model.domain)}
        for v in range(n):
            row = self.states[v]
            for c in range(self.n_cols):
                ch = row[c]
                if ch is not None:
                    if ch not in lut:
                        raise ASRRuntimeException("Invalid character %r for model %s" % (ch, model.getName()))
                    obs[v, c] = lut[ch]
        return obs

    def _workspace(self, model, rates):
        if self._ws is None or self._ws_model is not model:
            if self._ws is not None:
                self._ws.close()
            self._ws = lib.Workspace(model.handle, self.tree.handle(), max(self.n_cols, 1), device=self.device)
            self._ws_model = model
        self._ws.set_rates(self.n_cols, rates)  # None => 1.0 per column (Prediction.java:615-618)
        return self._ws

    def getJoint(self, model, rates=None):
        """Object[][] states[bpidx][pos]: Character from model.getDomain() or None (Prediction.java:614-649).  As in the
        reference only the ancestors' rows are filled (`for idx in getAncestorIndices()`, :636-641): extants' rows stay
        None -- the library's own output also carries the observed characters there (bnk_joint passes them through)."""
        ws = self._workspace(model, rates)
        st = ws.joint(self._obs(model), self.present)
        dom = model.domain
        leaf = [self.tree.isLeaf(v) for v in range(self.tree.getSize())]
        self.joint_states = [[None] * self.n_cols if leaf[v] else [None if s == NONE else dom[s] for s in row] for v, row in enumerate(st)]
        return self.joint_states

    def getMarginal(self, ancestor, model, rates=None):
        """EnumDistrib[] over positions for one ancestor, None where it is absent (Prediction.java:555-597); an
        unknown ancestor raises ASRRuntimeException (:561-562).  `ancestor`: an ID as text ("N3", "3", a label) as
        the reference takes it, or -- this mirror only -- an int = the branch point's index in the tree."""
        idx = ancestor if isinstance(ancestor, (int, np.integer)) else self.getBranchpointIndex(ancestor)
        if idx < 0 or idx >= self.tree.getSize():
            raise ASRRuntimeException("Invalid ancestor ID " + str(ancestor))
        ws = self._workspace(model, rates)
        try:
            post, _ = ws.marginal(self._obs(model), self.present, query=[idx], want_logl=False)
        except lib.BnkError as e:
            if e.code == -7:
                raise ASRRuntimeException("Invalid ancestor ID " + str(ancestor))
            raise
        d = [None if np.isnan(post[0, c]).any() else EnumDistrib(model.domain, post[0, c]) for c in range(self.n_cols)]
        self.distribs[idx] = d
        return d

    # ---- downstream assembly and outputs ----
    def getAncestors(self, mode):
        """{ancestor ID: POGraph} decorated with the inferred states (JOINT) or distributions (MARGINAL); ancestors
        that have not been inferred are not included (Prediction.java:483-503)"""
        if self.ancarr is None:
            raise ASRRuntimeException("No ancestor POGs: indel inference has not been performed")
        out = {}
        for idx in self.getAncestorIndices():
            g = self.ancarr.get(idx)
            if g is None:
                continue
            if mode == Prediction.JOINT and self.joint_states is not None:
                g.decorateNodes(self.joint_states[idx], "class dat.pog.SymNode")
                out[Prediction._ancestor_id(self.tree, idx)] = g
            elif mode == Prediction.MARGINAL and self.distribs.get(idx) is not None:
                g.decorateNodes(self.distribs[idx], "class dat.pog.EnumNode")
                out[Prediction._ancestor_id(self.tree, idx)] = g
        return out

    def getConsensus(self, anc):
        """positions of the ancestor's POG that make up its most supported start-to-end path
        (Prediction.getConsensus, src/asr/Prediction.java:723-780): every edge out of a visited node is weighted
        -log of the length-weighted share of the extants below the ancestor that make the same jump
        (POGTree.getEdgeRates), capped at 10000; then POGraph.getMostSupported (Dijkstra; unreciprocated edges
        cost 1000 x)."""
        bpidx = anc if isinstance(anc, (int, np.integer)) else self.getBranchpointIndex(anc)
        if bpidx is None or bpidx < 0 or self.ancarr is None or self.ancarr.get(bpidx) is None:
            raise ASRRuntimeException("Ancestor has not been inferred: index is " + str(bpidx))
        g = self.ancarr[bpidx]
        leaves = self._leaves_under(bpidx)
        import heapq
        queue = list(g.getForward())
        heapq.heapify(queue)
        visiting = set(queue)
        while queue:
            curr = heapq.heappop(queue)
            nexts = g.getForward(curr)
            if g.isEndNode(curr):
                nexts = nexts + [g.maxsize()]
            rates = self.pogTree.getEdgeRates(curr, nexts, leaves)
            for nx, r in zip(nexts, rates):
                edge = g.getEdge(curr, nx)
                w = -np.log(r)
                if edge is None:   # the target node may not have been inferred
                    edge = _pog.BidirEdge(False, False)
                    g.addEdge(curr, nx, edge)
                edge.weight = 10000.0 if w > 10000 else float(w)
            for nx in nexts:
                if nx != g.maxsize() and nx not in visiting:
                    heapq.heappush(queue, nx)
                    visiting.add(nx)
        return g.getMostSupported()

    def _leaves_under(self, bpidx):
        out, stack = [], [bpidx]
        while stack:
            v = stack.pop()
            ch = self.tree.getChildren(v)
            if not ch:
                out.append(v)
            stack.extend(reversed(ch))
        return out

    def getSequence(self, ancID, mode, gappy):
        """the inferred ancestor as a sequence along its consensus path (Prediction.getSequence :674-708): gappy =
        one entry per alignment column (None off the path), else only the path's entries"""
        bpidx = self.getBranchpointIndex(ancID)
        if bpidx == -1:
            raise ASRRuntimeException("Invalid ancestor ID: " + str(ancID))
        idxs = self.getConsensus(bpidx)
        if idxs is None:
            raise ASRRuntimeException("Failed to find optimal path for ancestor ID: " + str(ancID))
        elems = [None] * (self.n_cols if gappy else len(idxs))
        if mode == Prediction.JOINT:
            if self.joint_states is None:
                raise ASRRuntimeException("Joint inference has not been performed: " + str(ancID))
            for i, c in enumerate(idxs):
                elems[c if gappy else i] = self.joint_states[bpidx][c]
        else:
            d = self.distribs.get(bpidx)
            if d is None:
                raise ASRRuntimeException("Marginal inference has not been performed: " + str(ancID))
            for i, c in enumerate(idxs):
                elems[c if gappy else i] = d[c].getMax()
        return elems

    def ancestorsFasta(self, mode, gappy=True):
        """the text of <prefix>_ancestors.fa as asr.GRASP writes it (src/asr/GRASP.java:531-596): every inferred
        ancestor N<id> along its consensus path, in ancestor-ID order"""
        pogs = self.getAncestors(mode)
        ids = sorted(pogs, key=lambda k: (0, int(k)) if str(k).isdigit() else (1, str(k)))
        return _pog.fasta_text(["N" + str(k) for k in ids], [self.getSequence(k, mode, gappy) for k in ids])

    def ancestorsClustal(self, mode):
        """the same sequences in CLUSTAL format (asr.GRASP -of CLUSTAL, dat.file.AlnWriter.save :92-120)"""
        from . import files
        pogs = self.getAncestors(mode)
        ids = sorted(pogs, key=lambda k: (0, int(k)) if str(k).isdigit() else (1, str(k)))
        return files.clustal_text(["N" + str(k) for k in ids], [self.getSequence(k, mode, True) for k in ids])

    def treeNewick(self, ancestor_labels=True):
        """the tree as asr.GRASP writes <prefix>_ancestors.nwk: Newick.save(tree, file, MODE_ANCESTOR) -- ancestors
        labelled N<number> so that the FASTA / TSV outputs can be matched to branch points"""
        from . import files
        return files.newick_text(self.tree, files.NEWICK_ANCESTOR if ancestor_labels else files.NEWICK_DEFAULT)

    def distribTsv(self, ancID, model):
        """the text of <prefix>_N<id>.tsv (GRASP's DISTRIB format, src/asr/GRASP.java:599-630)"""
        bpidx = self.getBranchpointIndex(ancID)
        d = self.distribs.get(bpidx)
        if d is None:
            raise ASRRuntimeException("Marginal inference has not been performed: " + str(ancID))
        return _pog.distrib_tsv_text(d, model.getDomain())

    def toJSON(self):
        """ASR.json (Prediction.toJSON, src/asr/Prediction.java:213-226): version, datatype, the input (tree + extant
        POGs) and the ancestors' POGs in ancestor order, each carrying the columns it is present at ("Indices" =
        one row of the presence mask) and, once decorated, its node values"""
        if self.ancarr is None:
            raise ASRRuntimeException("No ancestor POGs: indel inference has not been performed")
        return {"GRASP_version": _pog.GRASP_VERSION, "Datatype": "Prediction", "Input": self.pogTree.toJSON(),
                "Ancestors": [self.ancarr[i].toJSON() for i in self.tree.getAncestors()]}

    @staticmethod
    def fromJSON(obj, device=-1):
        """Prediction.fromJSON (:133-155): the tree, the extants and the ancestors' POGs of a saved run -- e.g. an
        ASR.json a JVM GRASP wrote: its indel inference then feeds the GPU sweeps directly (GRASP -i <folder>,
        src/asr/GRASP.java:468-474)"""
        if isinstance(obj, str):
            obj = json.loads(obj)
        dt = obj.get("Datatype")
        if dt is not None and dt != "Prediction":
            raise ASRRuntimeException("Invalid input file: Wrong datatype " + str(dt) + " should be Prediction")
        if obj.get("Input") is None:
            raise ASRRuntimeException("Missing \"Input\" field in JSON")
        pt = POGTree.fromJSON(obj["Input"], IdxTree)
        jancs = obj.get("Ancestors")
        if jancs is None:
            raise ASRRuntimeException("Missing \"Ancestors\" field in JSON")
        tree = pt.tree
        if len(jancs) != len(tree.getAncestors()):
            raise ASRRuntimeException("Number of ancestors %d does not match tree %d" % (len(jancs), len(tree.getAncestors())))
        by_name = {}
        for a in jancs:
            g = POGraph.fromJSON(a)
            by_name[str(g.getName())] = g
        n_cols = jancs[0]["Size"] if jancs else 0
        states = [[None] * n_cols for _ in range(tree.getSize())]
        present = np.zeros((tree.getSize(), n_cols), dtype=np.uint8)
        for v, g in pt.extants.items():
            for c in g.getIndices():
                states[v][c] = g.nodes[c]
                present[v, c] = 1
        anc = {}
        for i in tree.getAncestors():
            g = by_name.get(Prediction._ancestor_id(tree, i))
            if g is None:
                raise ASRRuntimeException("Ancestor %s is missing from the JSON" % tree.getLabel(i))
            anc[i] = g
            for c in g.getIndices():
                present[i, c] = 1
        present = tree.handle().prune_orphans(present)   # Prediction.getTree removes orphans (GRASP.REMOVE_INDEL_ORPHANS)
        pred = Prediction(tree, states, present, device, ancestors=anc)
        pred.pogTree = pt
        return pred

    def save(self, filename):
        with open(filename, "w") as fh:
            fh.write(json.dumps(self.toJSON()))
            fh.write("\n")

    @staticmethod
    def load(filename, device=-1):
        with open(filename) as fh:
            return Prediction.fromJSON(json.loads(fh.read()), device)

    def getMarginalAll(self, model, rates=None):
        """All ancestors at once (what `-m` looped over every N<k> costs the reference 37 full passes)."""
        ws = self._workspace(model, rates)
        return ws.marginal(self._obs(model), self.present)

    def close(self):
        if self._ws is not None:
            self._ws.close()
            self._ws = None


class GapSubstModel:
    """bn.ctmc.GapSubstModel (src/bn/ctmc/GapSubstModel.java): a substitution model whose alphabet has the gap
    state appended, with deletion rate mu and insertion rate lambda (Rivas & Eddy 2008)"""

    def __init__(self, base, mu, lam):
        self.base = base                      # a SubstModel (its handle carries R and F)
        self.mu, self.lam = float(mu), float(lam)
        if self.mu + self.lam < 0:
            raise ValueError("mu + lambda must be >= 0")

    def getMu(self):
        return self.mu

    def getLambda(self):
        return self.lam

    def getName(self):
        return "GapSubstModel"

    def getDomain(self):
        return list(self.base.getDomain()) + ["-"]


class IndelPeeler:
    """asr.IndelPeeler's static entry points (src/asr/IndelPeeler.java) over bnk_indel_*: every column's
    gap-augmented likelihood in one device sweep instead of one thread-pool job per column.
    `states` is the alignment as Prediction takes it: states[node][col], None = gap, extants' rows only."""

    def __init__(self, tree, states, base, device=-1):
        self.tree, self.base = tree, base
        n_cols = len(states[0]) if len(states) else 0
        dom = {ch: i for i, ch in enumerate(base.getDomain())}
        obs = np.full((tree.getSize(), n_cols), NONE, dtype=np.uint8)
        for v in tree.getLeaves():
            row = states[v]
            if row is None:
                continue
            for c, ch in enumerate(row):
                if ch is not None:
                    if ch not in dom:
                        raise ASRRuntimeException("Invalid character %r for model %s" % (ch, base.getName()))
                    obs[v, c] = dom[ch]
        self.obs = obs
        self.h = lib.Indel(base.handle, tree.handle(), max(n_cols, 1), device)
        self.h.set_alignment(obs)

    def close(self):
        self.h.close()

    @staticmethod
    def computeColumnPriors(tree, states, model, geometricSeqLenParam, rates, device=-1):
        """double[numCols][numRates]: log P(column | rate) (IndelPeeler.java:93-123)"""
        p = IndelPeeler(tree, states, model.base, device)
        try:
            p.h.set_rates(model.mu, model.lam)
            return p.h.column_priors(rates, geometricSeqLenParam)
        finally:
            p.close()

    @staticmethod
    def calcProbAlnGivenTree(tree, states, model, geometricSeqLenParam, device=-1):
        """log P(alignment | tree, model) incl. the normalisation for unobserved all-gap columns (:125-161)"""
        p = IndelPeeler(tree, states, model.base, device)
        try:
            p.h.set_rates(model.mu, model.lam)
            return p.h.aln_logl(geometricSeqLenParam)
        finally:
            p.close()

    @staticmethod
    def optimiseMuLambda(min_val, max_val, substModelName, tree, geometricSeqLenParam, states, device=-1):
        """the mu = lambda that maximises calcProbAlnGivenTree, by Brent's method (:469-510); an unknown model
        name raises ValueError as the reference raises IllegalArgumentException"""
        try:
            handle = lib.Model(substModelName, indel_base=True)
        except lib.BnkError as e:
            raise ValueError(str(e))
        base = SubstModel(handle, substModelName, ALPHABETS[substModelName])
        p = IndelPeeler(tree, states, base, device)
        try:
            x, _ = p.h.optimise_mulambda(min_val, max_val, geometricSeqLenParam)
            return x
        finally:
            p.close()


def load_alignment(path, alphabet=None):
    """FASTA or CLUSTAL -> (names, rows of characters with '-' -> None)   (dat/file/Utils.java:103-138); with an
    alphabet (a model name or a string of symbols) unknown residues are an error as in the reference's readers"""
    from . import files
    if alphabet is None:
        alphabet = "ABCDEFGHIJKLMNOPQRSTUVWXYZ*."      # harness use: no validation beyond the format
    elif alphabet in ALPHABETS:
        alphabet = ALPHABETS[alphabet]
    try:
        return files.loadAlignment(path, alphabet)
    except files.ASRException as e:
        raise ASRRuntimeException(str(e))


def load_rates(path):
    """Position-specific evolutionary rates as asr.GRASP reads them (src/asr/GRASP.java:433-456): a tab-separated
    file with a header; the column named "Rate" holds the rates (first column if there is no such name) and the
    optional column "Site" the 1-based alignment column each rate belongs to (file order otherwise) -- e.g. the
    output of IQ-TREE's --rate.  Lines starting with '#' are comments.  Returns a float64 array."""
    with open(path) as fh:
        lines = [ln.rstrip("\n") for ln in fh if ln.strip() and not ln.startswith("#")]
    if not lines:
        raise ASRRuntimeException("Rates file is empty: %s" % path)
    header = lines[0].split("\t")
    rate_col = header.index("Rate") if "Rate" in header else 0
    site_col = header.index("Site") if "Site" in header else -1
    rows = [ln.split("\t") for ln in lines[1:]]
    rates = np.zeros(len(rows), dtype=np.float64)
    for i, r in enumerate(rows):
        try:
            idx = i if site_col < 0 else int(r[site_col]) - 1
            rates[idx] = float(r[rate_col])
        except (ValueError, IndexError):
            raise ASRRuntimeException("Rates file has invalid number format: %s" % "\t".join(r))
    return rates
