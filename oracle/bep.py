"""CPU oracle for bi-directional edge parsimony (BEP), the reference's default indel inference.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke, bench.py's cpu_baseline may import it; the
product never does).  A literal, pure-Python restatement of

  * asr.Prediction.PredictByBidirEdgeParsimony / patchAncestorsWithBEP  src/asr/Prediction.java:1443-1605
  * asr.Parsimony (Sankoff, unit costs, ALL optimal states)              src/asr/Parsimony.java:69-412
  * dat.phylo.TreeInstance ctor / encode                                 src/dat/phylo/TreeInstance.java:46-58, 141-165
  * dat.pog.POGTree ctor (extant POGs) / getEdgeInstance                 src/dat/pog/POGTree.java:44-86, 450-487
  * dat.pog.EdgeMap(.Directed)                                           src/dat/pog/EdgeMap.java:9-155
  * dat.pog.POGraph.createFromEdgeMap / isPath / isContiguous / nibble / getProtruding
                                                                         src/dat/pog/POGraph.java:314-335, 544-640
  * dat.pog.IdxGraph node / edge primitives                              src/dat/pog/IdxGraph.java:200-265, 346-450, 468-512

with GRASP's defaults RECODE_NULL = true, NIBBLE = true (src/asr/GRASP.java:36-40).  Sets with
hash-dependent iteration order in the reference only feed order-insensitive consumers (edge sets), so
plain Python sets / sorted lists are used.  Pinned against the documented Recon output on data/66.*
(docs/json-api.md:113-122: ancestors 0-2 get nodes {2, 4} and edges (-1,2), (2,4), (4,5)); parity with a
JVM run beyond that is unpinned (no JDK in this image).

One deviation, stated: patchAncestorsWithBEP loops `while (columns.size() > 0)`; if re-inference adds no
edge to any crippled POG the reference would spin forever.  Here (and in the library) the loop stops after
a round that changed nothing.
"""

INT_MAX = float(2147483647)


# ---------------------------------------------------------------------------------------------------
# Parsimony.Inference (Parsimony.java:215-412)
# ---------------------------------------------------------------------------------------------------
def parsimony(children, instance, is_leaf, recode_null):
    """instance: per-node value or None.  Returns per-node list of optimal values (decoded; the NULL
    symbol dropped, [] when only NULL is optimal), or None for nodes never reached."""
    n = len(children)
    possible = []
    for v in instance:                       # TreeInstance ctor: distinct non-null values
        if v is not None and v not in possible:
            possible.append(v)
    # TreeInstance.encode(recodeNull) (:141-165)
    key = list(possible) + ([None] if recode_null else [])
    inst = [None] * n
    for i in range(n):
        if instance[i] is not None:
            inst[i] = possible.index(instance[i])
        elif recode_null and is_leaf[i]:
            inst[i] = len(possible)
    nsym = len(key)
    scores = [[0.0] * nsym for _ in range(n)]
    optimal = [None] * n
    traceback = [[None] * nsym for _ in range(n)]
    for b in range(n):                       # Inference ctor (:232-247)
        if inst[b] is not None:
            optimal[b] = (inst[b],)          # singletonList
            for i in range(nsym):
                scores[b][i] = 0.0 if inst[b] == i else INT_MAX

    def forward(b):                          # (:266-318)
        if optimal[b] is not None:
            return scores[b]
        ch = children[b]
        if len(ch) == 0:
            for i in range(nsym):
                scores[b][i] = 0.0
            optimal[b] = []
            return scores[b]
        optimal[b] = []
        cs = [forward(c) for c in ch]
        for i in range(nsym):
            traceback[b][i] = [None] * len(ch)
        for c in range(len(ch)):
            for i in range(nsym):
                best, cnt = 9e9, 0
                for j in range(len(cs[c])):
                    ps = cs[c][j] + (0 if i == j else 1)
                    if ps < best:
                        best, cnt = ps, 1
                    elif ps == best:
                        cnt += 1
                traceback[b][i][c] = [j for j in range(len(cs[c])) if cs[c][j] + (0 if i == j else 1) == best]
                scores[b][i] += best
        return scores[b]

    def backward_node(b, state):             # (:367-389)
        if state in optimal[b]:
            return
        optimal[b].append(state)
        for c, child in enumerate(children[b]):
            for j in traceback[b][state][c]:
                backward_node(child, j)

    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 4 * n + 1000))
    forward(0)
    # backward() at the root (:327-360)
    if nsym > 0:
        best_index = 0
        for i in range(1, nsym):
            if scores[0][i] < scores[0][best_index]:
                best_index = i
        for ps in range(nsym):
            if scores[0][ps] == scores[0][best_index]:
                if isinstance(optimal[0], tuple):
                    # an instantiated root keeps its singleton; the reference would throw on add
                    continue
                optimal[0].append(ps)
                for c, child in enumerate(children[0]):
                    for j in traceback[0][ps][c]:
                        backward_node(child, j)
    # decode (Parsimony.infer :73-101)
    null_at_anc = recode_null
    out = [None] * n
    for i in range(n):
        if optimal[i] is None:
            continue
        found_null, mylist = False, []
        for sym in optimal[i]:
            if null_at_anc and sym == len(key) - 1:
                found_null = True
            else:
                mylist.append(key[sym])
        if found_null and not mylist:
            out[i] = []
        elif mylist:
            out[i] = mylist
        else:
            out[i] = list(optimal[i])        # stays as it was (an empty list)
    return out


# ---------------------------------------------------------------------------------------------------
# POGraph (directed, terminated)
# ---------------------------------------------------------------------------------------------------
class POG:
    def __init__(self, n):
        self.n = n
        self.node = [False] * n
        self.fwd = [None] * n
        self.bwd = [None] * n
        self.start = set()
        self.end = set()
        self.flag = {}                       # (from, to) -> bit 0 forward | bit 1 backward (POGraph.BidirEdge, :1039-1053)

    def is_node(self, i):
        return 0 <= i < self.n and self.node[i]

    def add_node(self, i):                   # IdxGraph.addNode (:346-357): resets the node's edge sets
        self.node[i] = True
        self.fwd[i] = set()
        self.bwd[i] = set()

    def add_edge(self, a, b, flags=None):    # IdxGraph.addEdge (:402-419); IdxEdgeGraph.addEdge (:147-153) then
        # stores the edge object, replacing the one an earlier call put there
        if self.is_node(a) and self.is_node(b):
            self.bwd[b].add(a)
            self.fwd[a].add(b)
        elif a == -1 and self.is_node(b):
            self.start.add(b)
        elif self.is_node(a) and b == self.n:
            self.end.add(a)
        else:
            return False
        if flags is not None:
            self.flag[(a, b)] = flags
        return True

    def remove_node(self, i):                # IdxGraph.removeNode (:373-391)
        for nx in list(self.fwd[i]):
            self.bwd[nx].discard(i)
        for pv in list(self.bwd[i]):
            self.fwd[pv].discard(i)
        self.node[i] = False
        self.fwd[i] = None
        self.bwd[i] = None
        self.start.discard(i)
        self.end.discard(i)

    def is_contiguous(self):                 # POGraph.isPath(-1, N) (:314-335), iteratively
        seen, stack = set(), sorted(self.start)
        while stack:
            v = stack.pop()
            if v in seen:
                continue
            seen.add(v)
            if v in self.end:
                return True
            stack.extend(self.fwd[v])
        return False

    def _sticks_out(self, i, direction_start):   # POGraph.nibble(index, direction) (:566-581)
        if direction_start and i > -1:
            if i not in self.start:
                return len(self.bwd[i]) == 0
        if (not direction_start) and i < self.n:
            if i not in self.end:
                return len(self.fwd[i]) == 0
        return False

    def topo(self):
        # every edge runs from a lower to a higher index, so ascending order is a topological order;
        # nibble / getProtruding only need "predecessors first" (POGraph.java:252-285 builds one such order)
        return [i for i in range(self.n) if self.node[i]]

    def nibble(self):                        # (:588-617)
        order = self.topo()
        cnt = 0
        for i in order:
            if self.node[i] and self._sticks_out(i, True):
                self.remove_node(i)
                cnt += 1
        for i in reversed(order):
            if self.node[i] and self._sticks_out(i, False):
                self.remove_node(i)
                cnt += 1
        return cnt

    def protruding(self):                    # (:622-646)
        order = self.topo()
        cols = set()
        for i in order:
            if self._sticks_out(i, False):
                cols.add(i)
        for i in reversed(order):
            if self._sticks_out(i, True):
                cols.add(-i)
        return cols

    def edges(self):
        e = set()
        for i in range(self.n):
            if self.node[i]:
                for j in self.fwd[i]:
                    e.add((i, j))
        for j in self.start:
            e.add((-1, j))
        for i in self.end:
            e.add((i, self.n))
        return e

    def edge_flags(self):                    # the BidirEdge of every edge the graph holds now
        return {e: self.flag.get(e, 0) for e in self.edges()}


def pog_from_edges(n_pos, edges, flags=None):   # POGraph.createFromEdgeMap(nPos, Directed) (:544-559)
    pog = POG(n_pos)
    for a, b in edges:
        if a != -1 and not pog.is_node(a):
            pog.add_node(a)
        if b != n_pos and not pog.is_node(b):
            pog.add_node(b)
        pog.add_edge(a, b, None if flags is None else flags[(a, b)])
    return pog


# ---------------------------------------------------------------------------------------------------
# POGTree.getEdgeInstance (:450-487) on extant POGs built from alignment rows (:44-86)
# ---------------------------------------------------------------------------------------------------
def edge_instance(rows, index, forward, n_pos):
    """rows: per tree node, a list of bools (True = residue) for leaves with a sequence, None otherwise."""
    inst = [None] * len(rows)
    for v, row in enumerate(rows):
        if row is None:
            continue
        nodes = [c for c in range(n_pos) if row[c]]
        if not nodes:
            continue
        if index == -1 or index == n_pos:
            if index == -1 and forward:
                inst[v] = nodes[0]
            elif index == n_pos and not forward:
                inst[v] = nodes[-1]
        else:
            if index == nodes[0] and not forward:
                inst[v] = -1
            elif index == nodes[-1] and forward:
                inst[v] = n_pos
            elif row[index]:
                k = nodes.index(index)
                inst[v] = nodes[k + 1] if forward else nodes[k - 1]
    return inst


def children_of(parent):
    ch = [[] for _ in parent]
    for v, p in enumerate(parent):
        if p >= 0:
            ch[p].append(v)
    return ch


def predict(parent, rows, n_pos, recode_null=True, nibble=True, max_rounds=10000):
    """Prediction.PredictByBidirEdgeParsimony: returns {ancestor index: POG}."""
    n = len(parent)
    children = children_of(parent)
    is_leaf = [len(c) == 0 for c in children]
    pif, pib = [None] * (n_pos + 2), [None] * (n_pos + 2)
    for i in range(-1, n_pos + 1):
        tf = edge_instance(rows, i, True, n_pos)
        if any(x is not None for x in tf):
            pif[i + 1] = parsimony(children, tf, is_leaf, recode_null)
        tb = edge_instance(rows, i, False, n_pos)
        if any(x is not None for x in tb):
            pib[i + 1] = parsimony(children, tb, is_leaf, recode_null)
    ancestors = {}
    for j in range(n):
        if not children[j]:
            continue
        edges = {}                           # EdgeMap.Directed.add(from, to, forward): the directions accumulate
        for i in range(-1, n_pos + 1):
            if i != n_pos and pif[i + 1] is not None:
                for nxt in pif[i + 1][j] or []:
                    edges[(i, nxt)] = edges.get((i, nxt), 0) | 1
            if i != -1 and pib[i + 1] is not None:
                for prv in pib[i + 1][j] or []:
                    edges[(prv, i)] = edges.get((prv, i), 0) | 2
        ancestors[j] = pog_from_edges(n_pos, sorted(edges), edges)
    # patchAncestorsWithBEP (:1513-1590)
    crippled = {}
    for j, pog in ancestors.items():
        if not pog.is_contiguous():
            crippled[j] = pog.protruding()
        elif nibble:
            pog.nibble()
    columns = set()
    for s in crippled.values():
        columns |= s
    rounds = 0
    while columns and rounds < max_rounds:
        rounds += 1
        cols_ordered = sorted(columns)
        pinf = []
        for idx in cols_ordered:
            col, fwd = abs(idx), idx > 0
            ti = edge_instance(rows, col, fwd, n_pos)
            pinf.append(parsimony(children, ti, is_leaf, False))
        changed = False
        for j in sorted(crippled):
            pog, idxs = ancestors[j], crippled[j]
            for k, idx in enumerate(cols_ordered):
                col, fwd = abs(idx), idx > 0
                if idx in idxs:
                    for inferred in pinf[k][j] or []:
                        if not pog.is_node(inferred) and inferred != n_pos and inferred != -1:
                            pog.add_node(inferred)
                            changed = True
                        before = len(pog.edges())
                        if fwd:              # new BidirEdge(true, false) / (false, true) replaces the edge's object
                            pog.add_edge(col, inferred, 1)
                        else:
                            pog.add_edge(inferred, col, 2)
                        changed |= len(pog.edges()) != before
            if not pog.is_contiguous():
                crippled[j] = pog.protruding()
            else:
                if nibble:
                    pog.nibble()
                del crippled[j]
        columns = set()
        for s in crippled.values():
            columns |= s
        if not changed:
            break                            # the reference would loop forever here (see module docstring)
    return ancestors


def presence_mask(parent, rows, n_pos, **kw):
    """present[node][col] as 0/1 lists: ancestors from their POG (Prediction.getTree :336-349), extants from
    their residues."""
    anc = predict(parent, rows, n_pos, **kw)
    out = []
    for v in range(len(parent)):
        if v in anc:
            out.append([1 if anc[v].node[c] else 0 for c in range(n_pos)])
        elif rows[v] is not None:
            out.append([1 if rows[v][c] else 0 for c in range(n_pos)])
        else:
            out.append([0] * n_pos)
    return out, anc
