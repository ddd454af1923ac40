"""Host-side graph logic between the device stages (labels, contacts, bridges) and the containment DAG.

BASELINE.json keeps this part of MDVContainment on the host.  This module computes what the reference's
``graph_logic.py`` / ``rank_logic.py`` compute (citations relative to ``/root/reference/mdvcontainment/``) --
periodic components with the reference's numbering, component and complement ranks, containment flags, the
component contact graph and the containment DAG, including every ordering the public API exposes -- on flat
arrays: the label-level work (connected pieces, lattice ranks) is one C++ union-find with shift potentials behind
the C-ABI (``mdvc_host_shift_graph_ranks``, ``csrc/hostgraph.cpp``), the component-level bookkeeping is NumPy, and
networkx objects are only built when a caller asks for them.  The reference needs 22.8 s of networkx walks for 48 k
labels; this takes tens of milliseconds and still scales at 10^6-10^7 labels.

Three places where Python ``set`` iteration order leaks into the reference's results are reproduced with real
``set`` objects of the running interpreter (never emulated): the numbering of the components of the minority
polarity (``_view_order``), the neighbour order of the component contact graph and the pop order that orients the
containment graph.  A plain-Python model of the same logic lives with the test oracle (graph_model.py); both are checked
against the unmodified reference where it is available.
"""
from __future__ import annotations

import ctypes
import itertools

import networkx as nx
import numpy as np

from . import _lib

_ZERO3 = np.array([0, 0, 0])
MAX_ORIENT_ITERATIONS = 100000          # the reference's guard in create_containment_graph (graph_logic.py:229-232)
_ORIENT_MESSAGE = ('During orientation of the containment graph we got in an iterative death '
                   '(max iterations 100,000).')


# ------------------------------------------------------------------------------------------------
# C-ABI call
# ------------------------------------------------------------------------------------------------
def shift_graph_ranks(n_nodes, edge_u, edge_v, edge_shift=None, skip=None, want_nodes=True):
    """Connected pieces and lattice ranks of a shift-labelled graph (``mdvc_host_shift_graph_ranks``).

    Returns (root int32[n] | None, rank int8[n] | None, max_rank)."""
    lib = _lib.load()
    eu = np.ascontiguousarray(edge_u, dtype=np.int32)
    ev = np.ascontiguousarray(edge_v, dtype=np.int32)
    es = None if edge_shift is None else np.ascontiguousarray(edge_shift, dtype=np.int32).reshape(-1, 3)
    sk = None if skip is None else np.ascontiguousarray(skip, dtype=np.uint8)
    root = np.empty(n_nodes, dtype=np.int32) if want_nodes else None
    rank = np.empty(n_nodes, dtype=np.int8) if want_nodes else None
    best = ctypes.c_int32(0)

    def ptr(a):
        return None if a is None or a.size == 0 else a.ctypes.data

    rc = lib.mdvc_host_shift_graph_ranks(int(n_nodes), ptr(eu), ptr(ev), ptr(es), int(len(eu)), ptr(sk), ptr(root),
                                         ptr(rank), ctypes.addressof(best))
    _lib.check(rc, "mdvc_host_shift_graph_ranks")
    return root, rank, int(best.value)


def component_neighbours(comp_of_node, n_comps, src, dst):
    """(cs, cd): ordered distinct neighbour components per component (``mdvc_host_component_neighbours``)."""
    lib = _lib.load()
    comp = np.ascontiguousarray(comp_of_node, dtype=np.int32)
    s = np.ascontiguousarray(src, dtype=np.int32)
    d = np.ascontiguousarray(dst, dtype=np.int32)
    cs = np.empty(len(s), dtype=np.int32)
    cd = np.empty(len(s), dtype=np.int32)
    n_out = ctypes.c_longlong(0)

    def ptr(a):
        return None if a.size == 0 else a.ctypes.data

    rc = lib.mdvc_host_component_neighbours(len(comp), ptr(comp), int(n_comps), ptr(s), ptr(d), len(s), ptr(cs), ptr(cd),
                                            ctypes.addressof(n_out))
    _lib.check(rc, "mdvc_host_component_neighbours")
    return cs[:n_out.value], cd[:n_out.value]


def adjacency(n_comps, eu, ev, unique=False):
    """(offsets int64[n_comps + 1], neighbours int64) in networkx's adjacency order (``mdvc_host_adjacency``;
    ``unique``: the sequence holds no pair twice in either direction, ``mdvc_host_adjacency_unique``)."""
    lib = _lib.load()
    u = np.ascontiguousarray(eu, dtype=np.int32)
    v = np.ascontiguousarray(ev, dtype=np.int32)
    offsets = np.zeros(int(n_comps) + 1, dtype=np.int64)
    nbrs = np.empty(2 * len(u), dtype=np.int32)
    fn = lib.mdvc_host_adjacency_unique if unique else lib.mdvc_host_adjacency
    rc = fn(int(n_comps), u.ctypes.data if u.size else None, v.ctypes.data if v.size else None, len(u),
            offsets.ctypes.data, nbrs.ctypes.data if nbrs.size else None)
    _lib.check(rc, "mdvc_host_adjacency")
    return offsets, nbrs[: int(offsets[-1])].astype(np.int64)


# ------------------------------------------------------------------------------------------------
# label graph
# ------------------------------------------------------------------------------------------------
class LabelGraph:
    """The reference's label contact graph (``create_contact_graph``, graph_logic.py:12-50) as arrays: nodes =
    the non-zero labels in the order given, every contact a pair of directed edges with zero cost, every bridge
    ``(l1, l2, s)`` an edge ``l1->l2`` with cost ``s`` and ``l2->l1`` with ``-s``, in that insertion order."""

    def __init__(self, unique_labels, contacts, bridges=None):
        self.labels = np.asarray(unique_labels)
        self.contacts = np.asarray(contacts).reshape(-1, 2)
        self.bridges = None if bridges is None else np.asarray(bridges).reshape(-1, 5)
        self.node_vals = self.labels[self.labels != 0].astype(np.int64)
        v = self.node_vals
        self._sorted = bool(len(v) < 2 or np.all(v[1:] > v[:-1]))
        if not self._sorted:
            self._perm = np.argsort(v, kind="stable")
            self._sorted_vals = v[self._perm]
        # contiguous -nF..-1, 1..nT (every grid this package labels): index = label + nF - (label > 0)
        self._n_neg = int(np.count_nonzero(v < 0))
        n = len(v)
        self._canonical = bool(self._sorted and n and (self._n_neg == 0 or (v[0] == -self._n_neg and v[self._n_neg - 1] == -1))
                               and (self._n_neg == n or (v[self._n_neg] == 1 and v[-1] == n - self._n_neg)))
        self._half = None
        self._rows = None

    @property
    def nodes(self):
        return self.node_vals.tolist()

    def index_of(self, values) -> np.ndarray:
        """Node index of every label value (all must be nodes)."""
        values = np.asarray(values, dtype=np.int64)
        if self._canonical:
            idx = values + self._n_neg - (values > 0)
            if len(idx) and (idx.min() < 0 or idx.max() >= len(self.node_vals)):
                raise KeyError("edge endpoint is not a label of the graph")
            return idx.astype(np.int32)
        sv = self.node_vals if self._sorted else self._sorted_vals
        pos = np.searchsorted(sv, values)
        pos[pos >= len(sv)] = 0
        if len(values) and not np.array_equal(sv[pos], values):
            raise KeyError("edge endpoint is not a label of the graph")
        return (pos if self._sorted else self._perm[pos]).astype(np.int32)

    def rows(self):
        """(contact rows, bridge rows) with the rows touching label 0 removed (graph_logic.py:38-39, 45-46), as node
        indices: (cu, cv), (bu, bv, shift[.,3])."""
        if self._rows is None:
            self._rows = self._compute_rows()
        return self._rows

    def _compute_rows(self):
        c = self.contacts
        c = c[(c != 0).all(axis=1)] if len(c) else c.reshape(0, 2)
        cu, cv = self.index_of(c[:, 0]), self.index_of(c[:, 1])
        if self.bridges is None or len(self.bridges) == 0:
            e = np.empty(0, dtype=np.int32)
            return (cu, cv), (e, e, np.empty((0, 3), dtype=np.int32))
        b = self.bridges
        b = b[(b[:, :2] != 0).all(axis=1)]
        return (cu, cv), (self.index_of(b[:, 0]), self.index_of(b[:, 1]), b[:, 2:].astype(np.int32))

    def half_edges(self):
        """Every directed edge in insertion order: (src, dst) node indices."""
        if self._half is None:
            (cu, cv), (bu, bv, _) = self.rows()
            src = np.concatenate([np.stack([cu, cv], 1).ravel(), np.stack([bu, bv], 1).ravel()])
            dst = np.concatenate([np.stack([cv, cu], 1).ravel(), np.stack([bv, bu], 1).ravel()])
            self._half = (src, dst)
        return self._half

    # --- dict forms (compatibility / small graphs) ------------------------------------------------
    @property
    def neighbors(self):
        """{label: {neighbour: None}} with first-insertion order = successor order of the reference's MultiDiGraph."""
        out = {n: {} for n in self.nodes}
        src, dst = self.half_edges()
        for u, v in zip(self.node_vals[src].tolist(), self.node_vals[dst].tolist()):
            out[u][v] = None
        return out

    def to_networkx(self):
        """The reference's MultiDiGraph, attribute for attribute (graph_logic.py:29-50)."""
        g = nx.MultiDiGraph()
        for n in self.nodes:
            g.add_node(n)
        value = _ZERO3
        for c in self.contacts:
            if 0 in c:
                continue
            g.add_edge(int(c[0]), int(c[1]), label=str(value), cost=value)
            g.add_edge(int(c[1]), int(c[0]), label=str(value), cost=value)
        if self.bridges is not None:
            for b in self.bridges:
                if 0 in b[:2]:
                    continue
                value = b[2:]
                g.add_edge(int(b[0]), int(b[1]), label=str(value), cost=value)
                g.add_edge(int(b[1]), int(b[0]), label=str(-value), cost=-value)
        return g


def _view_order(n_all: int, subset_vals: np.ndarray) -> np.ndarray:
    """Positions (into ``subset_vals``) in the order ``graph.subgraph(subset)`` is iterated.

    graph_logic.py:120-127 numbers the periodic components in the order
    ``nx.connected_components(nx.Graph(contact_graph.subgraph(labels)))`` yields them, i.e. by the first node of
    each component in the *view's* iteration order.  networkx iterates a node-filtered view over the graph's own
    node dict -- insertion order, the subset's order here -- unless the subset holds less than half of the graph's
    nodes: then it iterates the Python ``set`` built from the subset array, whose order is decided by CPython's
    hashing of the integers (``tests/test_host_graph.py`` checks this against the installed networkx)."""
    n_sub = len(subset_vals)
    if 2 * n_sub < n_all and n_sub > 1:
        vals = subset_vals.tolist()
        order_vals = np.fromiter(set(vals), dtype=np.int64, count=n_sub)
        if np.all(subset_vals[1:] > subset_vals[:-1]):
            return np.searchsorted(subset_vals, order_vals)
        lookup = {v: k for k, v in enumerate(vals)}
        return np.fromiter((lookup[v] for v in order_vals.tolist()), dtype=np.int64, count=n_sub)
    return np.arange(n_sub)


def _first_occurrence(keys: np.ndarray, bound: int | None = None) -> np.ndarray:
    """Indices of the first occurrence of every distinct key, ascending.  ``bound``: the keys are integers in
    [0, bound) -- then two scatter passes replace the sort behind ``np.unique`` (a fancy assignment with repeated
    indices keeps the value written last, so writing the positions in reverse leaves the first one)."""
    n = len(keys)
    if n == 0:
        return np.empty(0, dtype=np.int64)
    if bound is not None and bound <= 8 * n + 4096:
        first_pos = np.empty(int(bound), dtype=np.int64)
        first_pos[keys[::-1]] = np.arange(n - 1, -1, -1, dtype=np.int64)
        is_first = np.zeros(n, dtype=bool)
        is_first[first_pos[keys]] = True
        return np.flatnonzero(is_first)
    _, first = np.unique(keys, return_index=True)
    first.sort()
    return first


# ------------------------------------------------------------------------------------------------
# containment DAG without networkx
# ------------------------------------------------------------------------------------------------
class CompactDAG:
    """Containment DAG as arrays: ``nodes`` (component ids ascending, the order of ``np.unique(components_grid)``,
    containment_main.py:132) and the edges parent -> child in insertion order.  ``to_networkx()`` builds the
    reference's MultiDiGraph (graph_logic.py:224-240)."""

    def __init__(self, nodes, edge_src, edge_dst):
        self.nodes = np.asarray(nodes, dtype=np.int32)
        self.edge_src = np.asarray(edge_src, dtype=np.int64)
        self.edge_dst = np.asarray(edge_dst, dtype=np.int64)
        self._nx = None
        self._children = None

    def __len__(self):
        return len(self.nodes)

    def children(self):
        """{node id: [child ids in insertion order]} for nodes that have children."""
        if self._children is None:
            out = {}
            for u, v in zip(self.edge_src.tolist(), self.edge_dst.tolist()):
                out.setdefault(u, []).append(v)
            self._children = out
        return self._children

    def roots(self):
        has_parent = np.isin(self.nodes, self.edge_dst)
        return self.nodes[~has_parent].tolist()

    def edges(self):
        """Edge list in the order ``MultiDiGraph.edges()`` reports it: grouped by source in node order."""
        pos = np.searchsorted(self.nodes, self.edge_src)
        order = np.argsort(pos, kind="stable")
        return list(zip(self.edge_src[order].tolist(), self.edge_dst[order].tolist()))

    def to_networkx(self):
        if self._nx is None:
            g = nx.MultiDiGraph()
            for node in self.nodes:                      # np.int32 scalars, like np.unique(components_grid)
                g.add_node(node)
            for u, v in zip(self.edge_src, self.edge_dst):   # np.int64 scalars, like the component2labels keys
                g.add_edge(u, v)
            self._nx = g
        return self._nx


def format_dag_structure(G, ranks, counts=False, unit='nvoxels'):
    """Tree rendering identical to graph_logic.py:243-293; ``G`` is a networkx DAG or a CompactDAG."""
    if isinstance(G, CompactDAG):
        n = len(G)
        kids_of = G.children()
        roots = G.roots()

        def kids(node):
            return kids_of.get(node, ())
    else:
        n = len(G.nodes())
        roots = [node for node, deg in G.in_degree() if deg == 0]

        def kids(node):
            return list(G.successors(node))
    if counts is not False:
        out = [f'Containment Graph with {n} components (component: {unit}: rank):\n']
    else:
        out = [f'Containment Graph with {n} components (component: rank):\n']
    # explicit stack instead of the reference's recursion: deep hierarchies must not hit the recursion limit
    stack = [(root, '', k == len(roots) - 1) for k, root in reversed(list(enumerate(roots)))]
    while stack:
        node, prefix, last = stack.pop()
        tee = '└── ' if last else '├── '
        if counts is not False:
            out.append(f"{prefix}{tee}[{node}: {int(counts[node])}: {ranks[node]}]\n")
        else:
            out.append(f"{prefix}{tee}[{node}: {ranks[node]}]\n")
        ch = kids(node)
        child_prefix = prefix + ('    ' if last else '│   ')
        for k in range(len(ch) - 1, -1, -1):
            stack.append((ch[k], child_prefix, k == len(ch) - 1))
    return ''.join(out)


# ------------------------------------------------------------------------------------------------
# the periodic solve
# ------------------------------------------------------------------------------------------------
class PeriodicSolution:
    """Everything ``calc_containment_graph`` derives between ``find_bridges`` and ``create_containment_graph``
    (containment_main.py:72-133), as arrays indexed by *component index* (the reference's dict order: positive
    components 1..kp, then negative ones -1..-kn).  ``sol["component2labels"]`` etc. build the reference's dict /
    networkx forms on demand."""

    def __init__(self, lg: LabelGraph):
        self.label_graph = lg
        L = len(lg.node_vals)
        vals = lg.node_vals
        (cu, cv), (bu, bv, bs) = lg.rows()
        self._rows = (cu, cv, bu, bv, bs)

        # --- periodic components: pieces of the same-sign subgraphs (graph_logic.py:116-127) --------
        same_b = (vals[bu] > 0) == (vals[bv] > 0) if len(bu) else np.zeros(0, dtype=bool)
        same_c = (vals[cu] > 0) == (vals[cv] > 0) if len(cu) else np.zeros(0, dtype=bool)   # never true for real labels
        eu = np.concatenate([bu[same_b], cu[same_c]])
        ev = np.concatenate([bv[same_b], cv[same_c]])
        es = np.concatenate([bs[same_b], np.zeros((int(same_c.sum()), 3), dtype=np.int32)])
        root, rank, _ = shift_graph_ranks(L, eu, ev, es)

        comp_roots = []
        for positive in (True, False):
            sub = np.flatnonzero(vals > 0) if positive else np.flatnonzero(vals < 0)
            order = sub[_view_order(L, vals[sub])]
            r = root[order]
            comp_roots.append(r[_first_occurrence(r, L)])
        self.n_pos, self.n_neg = len(comp_roots[0]), len(comp_roots[1])
        K = self.n_pos + self.n_neg
        self.comp_root = np.concatenate(comp_roots).astype(np.int64)
        # component ids as in rank_logic.py:119,128 (np.arange -> np.int64)
        self.comp_ids = np.concatenate([np.arange(1, self.n_pos + 1), np.arange(-1, -self.n_neg - 1, -1)]).astype(np.int64)
        ci_of_root = np.full(L, -1, dtype=np.int32)
        ci_of_root[self.comp_root] = np.arange(K, dtype=np.int32)
        self.comp_of_node = ci_of_root[root] if L else np.empty(0, dtype=np.int32)

        # --- ranks (rank_logic.py:136-181) -----------------------------------------------------------
        self.comp_rank = rank[self.comp_root].astype(np.int64) if K else np.empty(0, dtype=np.int64)
        compl = np.where(self.comp_rank == 0, 3, 2).astype(np.int64)
        if np.any(self.comp_rank >= 2):
            all_u = np.concatenate([cu, bu]); all_v = np.concatenate([cv, bv])
            all_s = np.concatenate([np.zeros((len(cu), 3), dtype=np.int32), bs])
            for ci in np.flatnonzero(self.comp_rank >= 2):
                # largest rank among the connected pieces of everything else
                _, _, best = shift_graph_ranks(L, all_u, all_v, all_s, skip=(self.comp_of_node == ci), want_nodes=False)
                compl[ci] = best
        self.compl_rank = compl
        self.contained = self.comp_rank < self.compl_rank               # get_is_contained, rank_logic.py:246-260
        self._ccg = None
        self._cache = {}

    # --- label -> component table ---------------------------------------------------------------
    def lut(self, n_false: int, n_true: int) -> np.ndarray:
        """int32 [n_false + n_true + 1], index = label + n_false (0 for label 0): create_components_grid's mapping
        (label.py:21-38)."""
        out = np.zeros(n_false + n_true + 1, dtype=np.int32)
        out[self.label_graph.node_vals + n_false] = self.comp_ids[self.comp_of_node]
        return out

    def component_index(self, ids) -> np.ndarray:
        ids = np.asarray(ids, dtype=np.int64)
        return np.where(ids > 0, ids - 1, self.n_pos - ids - 1)

    # --- component contact graph (graph_logic.py:164-198) ----------------------------------------
    def contact_edges(self):
        """(src, dst) component indices of the edges of the component contact graph in the order the reference adds
        them: components in dict order, each one's neighbour components in the iteration order of the Python set
        the reference collects them in (the same container, filled in the same order, is used)."""
        if self._ccg is None:
            lg = self.label_graph
            K = len(self.comp_ids)
            src, dst = lg.half_edges()
            if lg._sorted:
                # add sequence -- components in order, their labels ascending, each label's successors in insertion
                # order -- with every (component, neighbour) pair at its first occurrence: compiled, O(labels + edges)
                cs, cd = component_neighbours(self.comp_of_node, K, src, dst)
            else:
                cs, cd = self.comp_of_node[src], self.comp_of_node[dst]
                keep = cs != cd
                cs, cd, lab = cs[keep], cd[keep], lg.node_vals[src[keep]]
                lab_rank = lab - lab.min() if len(lab) else lab
                order = np.argsort(cs.astype(np.int64) * (int(lab_rank.max()) + 1 if len(lab) else 1) + lab_rank, kind="stable")
                cs, cd = cs[order].astype(np.int64), cd[order].astype(np.int64)
                first = _first_occurrence(cs * K + cd)
                cs, cd = cs[first], cd[first]
            # The reference collects a component's neighbours in a Python set and adds the edges in that set's iteration
            # order.  More than 100 000 components it cannot orient at all (its guard fires), so beyond that there is
            # no order to reproduce and the first-occurrence order stays.
            if len(cs) and K <= MAX_ORIENT_ITERATIONS:
                brk = np.flatnonzero(cs[1:] != cs[:-1]) + 1
                starts = np.concatenate([[0], brk]); ends = np.concatenate([brk, [len(cs)]])
                multi = np.flatnonzero(ends - starts > 1)
                if len(multi):
                    ids = self.comp_ids[cd].tolist()
                    for s, e in zip(starts[multi].tolist(), ends[multi].tolist()):
                        ids[s:e] = set(ids[s:e])
                    cd = self.component_index(np.asarray(ids, dtype=np.int64))
            self._ccg = (np.asarray(cs, dtype=np.int64), np.asarray(cd, dtype=np.int64))
        return self._ccg

    def contact_adjacency(self):
        """CSR (offsets, neighbours) of the component contact graph with networkx's adjacency order: a neighbour
        appears where the first edge joining the two components was added, from either side."""
        if "adj" not in self._cache:
            cs, cd = self.contact_edges()
            # the list is grouped by component in ascending order, holds no (component, neighbour) pair twice and no
            # component as its own neighbour, and is symmetric (every label contact / bridge is a pair of directed
            # edges): the first occurrence of an undirected edge is the one with component < neighbour
            grouped = len(cs) < 2 or bool(np.all(cs[1:] >= cs[:-1]))
            if grouped:
                first = cs < cd
                self._cache["adj"] = adjacency(len(self.comp_ids), cs[first], cd[first], unique=True)
            else:
                self._cache["adj"] = adjacency(len(self.comp_ids), cs, cd)
        return self._cache["adj"]

    def component_contact_graph(self):
        """networkx form of the above, node order included (a component is inserted at its turn unless an earlier
        component's edge already created it)."""
        if "ccg_nx" not in self._cache:
            g = nx.Graph()
            cs, cd = self.contact_edges()
            ids = self.comp_ids
            bounds = np.searchsorted(cs, np.arange(len(ids) + 1))
            for k, comp in enumerate(ids):
                g.add_node(comp)
                for j in range(bounds[k], bounds[k + 1]):
                    g.add_edge(comp, ids[cd[j]])
            self._cache["ccg_nx"] = g
        return self._cache["ccg_nx"]

    # --- containment DAG (graph_logic.py:200-241) -------------------------------------------------
    def containment_dag(self, max_iterations=MAX_ORIENT_ITERATIONS) -> CompactDAG:
        """Orient the component contact graph from the non-contained components inwards.  The reference pops query
        components from a Python set and lets the first popped neighbour claim a component, so the same sequence of
        set operations is performed on real sets; it also gives up (AssertionError) after 100 000 pops, i.e. for more
        than 100 000 components.  ``max_iterations=None`` lifts that guard (grids the reference cannot process at
        all): beyond the guard there is no reference order to reproduce and a level-synchronous sweep is used."""
        K = len(self.comp_ids)
        nodes = np.sort(self.comp_ids).astype(np.int32)
        offsets, nbrs = self.contact_adjacency()
        if max_iterations is None and K > MAX_ORIENT_ITERATIONS:
            return CompactDAG(nodes, *self._orient_by_levels(offsets, nbrs))
        ids = self.comp_ids.tolist()
        contained = self.contained.tolist()
        outside = set([c for c, inside in zip(ids, contained) if not inside])
        stack = outside.copy()
        # neighbour lists only where a pop can do anything: a contained component with a single neighbour was claimed
        # by that neighbour, which is outside by then
        deg = np.diff(offsets)
        busy = np.flatnonzero((deg > 1) | ((deg > 0) & ~self.contained))
        flat = self.comp_ids[nbrs].tolist()
        off = offsets.tolist()
        adj = {ids[k]: flat[off[k]:off[k + 1]] for k in busy.tolist()}
        get, pop = adj.get, stack.pop
        parents, kids = [], []
        # every component is popped at most once, so the reference's guard (assert before the 100 001st pop) can only
        # fire when there are more components than that: checked per pop only then
        guarded = max_iterations is not None and K > max_iterations
        counter = 0
        while stack:
            if guarded:
                assert counter < max_iterations, _ORIENT_MESSAGE
                counter += 1
            q = pop()
            nbs = get(q)
            if nbs is None:
                continue
            new = [nb for nb in nbs if nb not in outside]        # neighbours are distinct: same as testing one by one
            if new:
                outside.update(new)                               # sequential adds, in neighbour order
                stack.update(new)
                parents.append(q)
                kids.append(new)
        if not parents:
            return CompactDAG(nodes, [], [])
        eu = np.repeat(np.asarray(parents, dtype=np.int64), [len(k) for k in kids])
        ev = np.fromiter(itertools.chain.from_iterable(kids), dtype=np.int64, count=len(eu))
        return CompactDAG(nodes, eu, ev)

    def _orient_by_levels(self, offsets, nbrs):
        K = len(self.comp_ids)
        claimed = ~self.contained.copy()
        frontier = np.flatnonzero(claimed)
        eu, ev = [], []
        while len(frontier):
            deg = offsets[frontier + 1] - offsets[frontier]
            src = np.repeat(frontier, deg)
            pos = np.arange(int(deg.sum())) - np.repeat(np.cumsum(deg) - deg, deg) + np.repeat(offsets[frontier], deg)
            dst = nbrs[pos]
            open_ = ~claimed[dst]
            src, dst = src[open_], dst[open_]
            first = _first_occurrence(dst, K)              # the first frontier component (in order) claims the child
            src, dst = src[first], dst[first]
            claimed[dst] = True
            eu.append(self.comp_ids[src]); ev.append(self.comp_ids[dst])
            frontier = dst
        if not eu:
            return np.empty(0, dtype=np.int64), np.empty(0, dtype=np.int64)
        return np.concatenate(eu), np.concatenate(ev)

    # --- the reference's dict forms ---------------------------------------------------------------------
    def _component2labels(self):
        vals = self.label_graph.node_vals
        order = np.lexsort((vals, self.comp_of_node))
        sv = vals[order].tolist()
        bounds = np.searchsorted(self.comp_of_node[order], np.arange(len(self.comp_ids) + 1)).tolist()
        return {cid: tuple(sv[bounds[k]:bounds[k + 1]]) for k, cid in enumerate(self.comp_ids)}

    def _build(self, key):
        if key == "component2labels":
            return self._component2labels()
        if key == "label2component":
            out = {}
            for comp, labs in self["component2labels"].items():
                for lab in labs:
                    out[lab] = comp
            return out
        if key == "component_ranks":
            return {c: int(r) for c, r in zip(self.comp_ids, self.comp_rank.tolist())}
        if key == "complement_ranks":
            return {c: int(r) for c, r in zip(self.comp_ids, self.compl_rank.tolist())}
        if key == "is_contained":
            return {c: bool(v) for c, v in zip(self.comp_ids, self.contained.tolist())}
        if key == "component_contact_graph":
            return self.component_contact_graph()
        if key == "label_graph":
            return self.label_graph
        raise KeyError(key)

    def __getitem__(self, key):
        if key not in self._cache:
            self._cache[key] = self._build(key)
        return self._cache[key]


def solve_periodic(unique_labels, contacts, bridges) -> PeriodicSolution:
    return PeriodicSolution(LabelGraph(unique_labels, contacts, bridges))


def sum_by_index(index, values, n) -> np.ndarray:
    """int64 [n]: values summed per index.  np.bincount accumulates in float64, exact while every sum stays below
    2^53 (voxel counts do: a grid has fewer than 2^53 voxels); np.add.at is ~50x slower at 10^7 labels."""
    if n == 0 or len(index) == 0:
        return np.zeros(n, dtype=np.int64)
    return np.bincount(index, weights=np.asarray(values, dtype=np.float64), minlength=n).astype(np.int64)


def counts_by_component(lut: np.ndarray, volumes: np.ndarray):
    """Voxels per component from the label -> component table and the label volumes (containment_main.py:510-512
    without a pass over the grid).  Returns (ids int32 ascending, counts int64) of the components that own voxels."""
    lo = int(lut.min()) if len(lut) else 0
    acc = sum_by_index(lut.astype(np.int64) - lo, volumes, (int(lut.max()) - lo + 1) if len(lut) else 0)
    keep = np.flatnonzero(acc)
    return (keep + lo).astype(np.int32), acc[keep]


# ------------------------------------------------------------------------------------------------
# generic networkx orientation (slab=True path and callers that hold networkx graphs)
# ------------------------------------------------------------------------------------------------
def containment_graph(is_contained, unique_components, contact_graph):
    """create_containment_graph (graph_logic.py:200-241) on a networkx contact graph.  ``set.pop`` order decides which
    parent claims a component first and the order of children, so the same set operations are performed."""
    outside = set(c for c, v in is_contained.items() if v == False)   # noqa: E712 (np.bool_ safe)
    inside = set(c for c, v in is_contained.items() if v == True)     # noqa: E712
    dag = nx.MultiDiGraph()
    for node in unique_components:
        dag.add_node(node)
    frontier = outside.copy()
    guard = 0
    while len(frontier):
        assert guard < MAX_ORIENT_ITERATIONS, _ORIENT_MESSAGE
        guard += 1
        query = frontier.pop()
        for nb in contact_graph.neighbors(query):
            if nb not in outside:
                dag.add_edge(query, nb)
                inside.remove(nb)
                outside.add(nb)
                frontier.add(nb)
    return dag


def solve_slab(unique_labels, contacts, outside_labels):
    """slab=True branch (containment_main.py:73-75, 110-111, 126-127; rank_logic.py:183-244): labels
    are the components; a label touching the outer faces of the two largest dimensions is outside."""
    lg = LabelGraph(unique_labels, contacts, None)
    outside = set(int(v) for v in outside_labels)
    ranks, flags = {}, {}
    for lab in np.asarray(unique_labels):
        hit = int(lab) in outside
        ranks[lab] = 1 if hit else 0
        flags[lab] = not hit
    return {"label_graph": lg, "component_ranks": ranks, "is_contained": flags}
