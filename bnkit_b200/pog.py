"""Partial-order graphs and the output side of asr.Prediction: what turns the states /
posteriors / presence masks of the GPU sweeps into the objects and files a GRASP user diffs.

Host-side mirror (Python, as the reference's toolchain is absent) of

    dat.pog.IdxGraph / IdxEdgeGraph / POGraph      src/dat/pog/IdxGraph.java:346-450, :1040-1077,
                                                   src/dat/pog/IdxEdgeGraph.java:46-71, :147-153,
                                                   src/dat/pog/POGraph.java:111-153 (fromJSON), :852-866 (decorateNodes),
                                                   :941-1090 (StatusEdge / BidirEdge)
    dat.pog.DijkstraSearch                         src/dat/pog/GraphSearch.java:77-206 (getMostSupported's default)
    dat.pog.POGTree.getEdgeRates / toJSON          src/dat/pog/POGTree.java:148-174, :276-297
    dat.file.FastaWriter.save / TSVFile.saveObjects src/dat/file/FastaWriter.java:117-142, src/dat/file/TSVFile.java:705-716

Integer work on small graphs (one per ancestor), downstream of the hot path; nothing here is on the GPU.
"""
import json
import math

GRASP_VERSION = "bnkit_b200 (format of GRASP 04-May-2023)"
ACCEPTED_VERSIONS = {GRASP_VERSION, "04-May-2023", "12-Dec-2022", "28-Aug-2022"}
EDGETYPE_BIDIR = "class dat.pog.POGraph$BidirEdge"
EDGETYPE_STATUS = "class dat.pog.POGraph$StatusEdge"


class ASRRuntimeException(RuntimeError):
    """asr.ASRRuntimeException (unchecked)"""


class BidirEdge:
    """POGraph.BidirEdge: inferred by a forward and / or a backward edge instance; reciprocated = both; weight is
    set by Prediction.getConsensus."""
    __slots__ = ("forward", "backward", "weight")

    def __init__(self, forward, backward, weight=0.0):
        self.forward, self.backward, self.weight = bool(forward), bool(backward), float(weight)

    def getReciprocated(self):
        return self.forward and self.backward

    def toJSON(self):
        w = self.weight
        return {"Recip": self.getReciprocated(), "Weight": int(w) if w == int(w) else w,
                "Forward": self.forward, "Backward": self.backward}

    @staticmethod
    def fromJSON(obj, edgetype):
        if edgetype == EDGETYPE_BIDIR:
            return BidirEdge(obj.get("Forward", False), obj.get("Backward", False), obj.get("Weight", 0))
        r = bool(obj.get("Recip", False))   # StatusEdge: only the reciprocation status is stored
        return BidirEdge(r, r, obj.get("Weight", 0))


class POGraph:
    """dat.pog.POGraph: directed, terminated graph over the alignment's columns 0 .. n - 1; -1 is the start
    terminal, n the end terminal.  nodes: {column: value} (None until decorated)."""

    def __init__(self, n, name=""):
        self.n = int(n)
        self.name = name
        self.nodes = {}
        self.fwd, self.bwd = {}, {}
        self.starts, self.ends = set(), set()
        self.edges = {}            # (from, to) -> BidirEdge
        self.nodetype = "class dat.pog.Node"

    # ---- construction ----
    @staticmethod
    def fromEdges(n, frm, to, flags, name=""):
        """from the arrays bnk_bep_infer hands out (or any edge list): createFromEdgeMap, POGraph.java:544-559"""
        g = POGraph(n, name)
        for a, b, f in zip(frm, to, flags):
            a, b, f = int(a), int(b), int(f)
            if a != -1 and a not in g.nodes:
                g.addNode(a)
            if b != n and b not in g.nodes:
                g.addNode(b)
            g.addEdge(a, b, BidirEdge(f & 1, f & 2))
        return g

    @staticmethod
    def fromSequence(n, present_cols, name=""):
        """the POG of an extant: its residue columns in order (POGTree ctor, src/dat/pog/POGTree.java:44-86)"""
        g = POGraph(n, name)
        cols = [int(c) for c in present_cols]
        prev = -1
        for c in cols:
            g.addNode(c)
            g.addEdge(prev, c, BidirEdge(True, True))
            prev = c
        if cols:
            g.addEdge(prev, n, BidirEdge(True, True))
        return g

    def addNode(self, i, value=None):
        self.nodes[i] = value
        self.fwd[i], self.bwd[i] = set(), set()

    def isNode(self, i):
        return i in self.nodes

    def addEdge(self, a, b, edge=None):
        """IdxGraph.addEdge (:402-419) + IdxEdgeGraph.addEdge (:147-153): an existing edge's object is replaced"""
        if a in self.nodes and b in self.nodes:
            self.fwd[a].add(b)
            self.bwd[b].add(a)
        elif a == -1 and b in self.nodes:
            self.starts.add(b)
        elif a in self.nodes and b == self.n:
            self.ends.add(a)
        else:
            return False
        if edge is not None:
            self.edges[(a, b)] = edge
        return True

    # ---- accessors ----
    def maxsize(self):
        return self.n

    def size(self):
        return len(self.nodes)

    def getName(self):
        return self.name

    def getStarts(self):
        return sorted(self.starts)

    def getEnds(self):
        return sorted(self.ends)

    def getForward(self, i=None):
        return sorted(self.starts) if i is None or i == -1 else sorted(self.fwd[i])

    def getBackward(self, i=None):
        return sorted(self.ends) if i is None or i == self.n else sorted(self.bwd[i])

    def isEndNode(self, i):
        return i in self.ends

    def isStartNode(self, i):
        return i in self.starts

    def isEdge(self, a, b):
        if a == -1:
            return b in self.starts
        if b == self.n:
            return a in self.ends
        return a in self.nodes and b in self.fwd[a]

    def getEdge(self, a, b):
        return self.edges.get((a, b))

    def isPath(self, a, b):
        """POGraph.isPath (src/dat/pog/POGraph.java:314-328): a path from a (-1 = start terminal) to b (n = end
        terminal); a node has a path to itself.  The reference recurses forward; this walks the same edges with a stack."""
        if a == b:
            return True
        if a != -1 and a not in self.nodes:
            return False
        if b != self.n and b not in self.nodes:
            return False
        seen, stack = set(), [a]
        while stack:
            v = stack.pop()
            if v == b or (b == self.n and v in self.ends):
                return True
            if v in seen:
                continue
            seen.add(v)
            stack.extend(self.getForward(v))
        return False

    def isContiguous(self):
        """POGraph.isContiguous (:334-336)"""
        return self.isPath(-1, self.n)

    def getIndices(self):
        return sorted(self.nodes)

    def decorateNodes(self, values, nodetype="class dat.pog.SymNode"):
        """POGraph.decorateNodes (:852-866): one value per alignment column; None where the ancestor has no node.
        SymNode.toArray keeps a node only where the value is non-null -- the graph's node set follows."""
        if values is None:
            raise ASRRuntimeException("Attempting to assign uninstantiated states to POG")
        if len(values) != self.n:
            raise ASRRuntimeException("Attempting to assign invalid states to POG")
        for i in self.nodes:
            self.nodes[i] = values[i]
        self.nodetype = nodetype

    # ---- JSON (IdxGraph.toJSON + IdxEdgeGraph.toJSON) ----
    def toJSON(self):
        idx = self.getIndices()
        obj = {"Datatype": "class dat.pog.POGraph", "GRASP_version": GRASP_VERSION, "Name": self.name if self.name != "" else None,
               "Size": self.n, "Terminated": True, "Directed": True, "Starts": self.getStarts(), "Ends": self.getEnds(),
               "Indices": idx, "Adjacent": [sorted(self.fwd[i]) for i in idx]}
        if idx:
            obj["Nodetype"] = self.nodetype
            obj["Nodes"] = [_node_json(self.nodes[i]) for i in idx]
        if self.edges:  # deterministic order (the reference's follows HashMap iteration): sorted by (from, to)
            keys = sorted(self.edges)
            obj["Edgeindices"] = [[a, b] for a, b in keys]
            obj["Edges"] = [self.edges[k].toJSON() for k in keys]
            obj["Edgetype"] = EDGETYPE_BIDIR
        return obj

    @staticmethod
    def fromJSON(obj):
        """POGraph.fromJSON (:111-153) + IdxGraph.useJSON"""
        if isinstance(obj, str):
            obj = json.loads(obj)
        version = obj.get("GRASP_version")
        if version is not None and version not in ACCEPTED_VERSIONS:
            raise ASRRuntimeException("Invalid version: " + str(version))
        try:
            n = int(obj["Size"])
            if not (obj["Terminated"] and obj["Directed"]):
                raise ASRRuntimeException("POGraph format is wrong in JSON object: is not terminated or not directed")
            g = POGraph(n, obj.get("Name") or "")
            idx = obj.get("Indices", [])
            nodes = obj.get("Nodes")
            g.nodetype = obj.get("Nodetype", g.nodetype)
            for k, i in enumerate(idx):
                g.addNode(int(i), None if nodes is None else _node_value(nodes[k]))
            for k, i in enumerate(idx):
                for j in obj["Adjacent"][k]:
                    g.addEdge(int(i), int(j))
            for j in obj.get("Starts", []):
                g.addEdge(-1, int(j))
            for i in obj.get("Ends", []):
                g.addEdge(int(i), n)
            et = obj.get("Edgetype")
            if et is not None:
                for (a, b), e in zip(obj["Edgeindices"], obj["Edges"]):
                    g.addEdge(int(a), int(b), BidirEdge.fromJSON(e, et))
            return g
        except KeyError as e:
            raise ASRRuntimeException("POGraph format is wrong in JSON object: missing %s" % e)

    # ---- the consensus path: DijkstraSearch(this, -1, N, PRIORITY_RECIP_WEIGHT).getOnePath() ----
    def getMostSupported(self):
        """GraphSearch.java:105-206.  Priority of an edge = its weight if reciprocated, 1000 x its weight otherwise
        (:181).  Returns the node indices of one cheapest start-to-end path, or None.  Ties: `closed[next]` collects
        every equally cheap predecessor and getOnePath takes `closed[current].iterator().next()` -- the first element
        of a java.util.HashSet<Integer>, i.e. the one in the lowest bucket (h ^ h >>> 16) & (capacity - 1)."""
        N = self.n
        actual = {}
        closed = {}
        visited = set()
        current = -1
        while current != N:
            nexts = self.getForward(current)
            if current != -1 and self.isEndNode(current):
                nexts = nexts + [N]
            visited.add(current)
            cost_sofar = 0.0 if current == -1 else actual[current]
            for nxt in nexts:
                edge = self.getEdge(current, nxt)
                if edge is None:
                    continue
                w = edge.weight if edge.getReciprocated() else 1000 * edge.weight
                cand = cost_sofar + w
                if actual.get(nxt, math.inf) > cand:
                    actual[nxt] = cand
                    closed[nxt] = [current]
                elif actual.get(nxt, math.inf) == cand and nxt in closed:
                    if current not in closed[nxt]:
                        closed[nxt].append(current)
            cheapest, pick = math.inf, None
            for i in sorted(actual):              # `for (int i = 0; i < actual.length; i ++)`: ascending index, strict <
                if i not in visited and cheapest > actual[i]:
                    cheapest, pick = actual[i], i
            if pick is None:
                break
            current = pick
        if N not in closed:
            return None
        path = []
        current = N
        while True:
            current = _hashset_first(closed[current])
            if current == -1:
                break
            path.append(current)
        return path[::-1]


def _hashset_first(members):
    """first element a java.util.HashSet<Integer> iterates: lowest bucket of a 16-slot table (the sets here hold
    the few equally cheap predecessors of one node), insertion order inside a bucket"""
    cap = 16
    while len(members) > 0.75 * cap:
        cap *= 2

    def bucket(v):
        h = v & 0xFFFFFFFF
        return (h ^ (h >> 16)) & (cap - 1)
    best = None
    for m in members:
        if best is None or bucket(m) < bucket(best):
            best = m
    return best


def _node_json(value):
    """SymNode.toJSON / EnumNode.toJSON"""
    if value is None:
        return {}
    if hasattr(value, "distrib"):   # an EnumDistrib: {"Pr": [...], "Domain": ...} as bn.prob.EnumDistrib.toJSON writes it
        return {"Pr": [float(p) for p in value.distrib], "Domain": {"Size": len(value.domain), "Values": list(value.domain), "Datatype": "Character"}}
    return {"Value": value}


def _node_value(obj):
    if "Value" in obj:
        return obj["Value"]
    if "Pr" in obj:
        return obj
    return None


class POGTree:
    """dat.pog.POGTree restricted to the output side: the tree, the extants' POGs, edge rates."""

    def __init__(self, tree, extants):
        self.tree = tree
        self.extants = extants    # {leaf index: POGraph}

    @staticmethod
    def fromAlignment(tree, states):
        n_cols = len(states[0]) if len(states) else 0
        ext = {}
        for v in tree.getLeaves():
            cols = [c for c in range(n_cols) if states[v][c] is not None]
            g = POGraph.fromSequence(n_cols, cols, tree.getLabel(v))
            for c in cols:
                g.nodes[c] = states[v][c]
            g.nodetype = "class dat.pog.SymNode"
            ext[v] = g
        return POGTree(tree, ext)

    def getEdgeRates(self, frm, targets, extants):
        """POGTree.getEdgeRates (:276-297): pseudo count 1 per target; every extant with the edge adds its length"""
        rates = [1.0] * len(targets)
        denom = float(len(targets))
        for sub in extants:
            g = self.extants.get(sub)
            if g is None:
                continue
            ln = g.size()
            denom += ln
            for i, to in enumerate(targets):
                if g.isEdge(frm, to):
                    rates[i] += ln
        return [r / denom for r in rates]

    def toJSON(self):
        return {"Tree": self.tree.toJSON(), "Extants": [self.extants[v].toJSON() for v in self.tree.getLeaves()]}

    @staticmethod
    def fromJSON(obj, tree_cls):
        if "Extants" not in obj:
            raise ASRRuntimeException("Failed to load extants: no field \"Extants\" in JSON.")
        if "Tree" not in obj:
            raise ASRRuntimeException("Failed to load tree: no field \"Tree\" in JSON.")
        tree = tree_cls.fromJSON(obj["Tree"])
        by_name = {}
        for e in obj["Extants"]:
            g = POGraph.fromJSON(e)
            by_name[g.getName()] = g
        ext = {}
        for v in tree.getLeaves():
            if tree.getLabel(v) in by_name:
                ext[v] = by_name[tree.getLabel(v)]
        return POGTree(tree, ext)


# ---- writers ----
def fasta_text(names, seqs, line_width=60):
    """FastaWriter.save(String[], Object[][]) (:117-142): '>' + name, then the sequence, a newline BEFORE every 60
    symbols (so also right after the defline), '-' for null"""
    out = []
    for name, seq in zip(names, seqs):
        if seq is None:
            continue
        out.append(">" + name)
        for i, ch in enumerate(seq):
            if i % line_width == 0:
                out.append("\n")
            out.append("-" if ch is None else str(ch))
        out.append("\n")
    return "".join(out)


def java_double(x):
    """Double.toString(x) (what TSVFile.saveObjects writes for a Double): the shortest digits that round-trip,
    plain decimal for 1e-3 <= |x| < 1e7 and d.dddE<n> otherwise, always at least one digit after the point"""
    from decimal import Decimal
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0.0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    sign, digits, exp = Decimal(repr(x)).as_tuple()
    digits = list(digits)
    while len(digits) > 1 and digits[-1] == 0:   # normalise: no trailing zeros in the digit string
        digits.pop()
        exp += 1
    e10 = exp + len(digits) - 1                  # x = d.ddd x 10^e10
    ds = "".join(str(d) for d in digits)
    pre = "-" if sign else ""
    if -3 <= e10 < 7:
        if e10 >= 0:
            whole = ds[: e10 + 1].ljust(e10 + 1, "0")
            frac = ds[e10 + 1:] or "0"
        else:
            whole, frac = "0", "0" * (-e10 - 1) + ds
        return pre + whole + "." + frac
    return pre + ds[0] + "." + (ds[1:] or "0") + "E" + str(e10)


def distrib_tsv_text(distribs, domain):
    """the DISTRIB table of asr.GRASP (src/asr/GRASP.java:599-630) through TSVFile.saveObjects (:705-716): header
    Index + the model's alphabet in model order, one row per alignment column (1-based), empty cells where the
    ancestor is absent"""
    rows = ["\t".join(["Index"] + [str(s) for s in domain])]
    for j, d in enumerate(distribs):
        if d is None:
            rows.append("\t".join([str(j + 1)] + [""] * len(domain)))
        else:
            rows.append("\t".join([str(j + 1)] + [java_double(d.get(k)) for k in range(len(domain))]))
    return "\n".join(rows) + "\n"
