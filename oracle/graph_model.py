"""Plain-Python model of the host graph stage -- TEST INFRASTRUCTURE (oracle), never imported by the product.

This is round 1's dict-based restatement of the reference's ``graph_logic.py`` / ``rank_logic.py`` (citations
relative to ``/root/reference/mdvcontainment/``).  It was checked against the live reference on random grids and is
kept as an independent cross-check of the array-based ``mdvcontainment_b200/host_graph.py`` (C++ union-find with
shift potentials + NumPy), which replaced it in the product because it did not scale past ~10^4 labels.
"""
from __future__ import annotations

from collections import deque

import networkx as nx
import numpy as np

_ZERO3 = np.array([0, 0, 0])


# ------------------------------------------------------------------------------------------------
# label graph
# ------------------------------------------------------------------------------------------------
class LabelGraph:
    """Adjacency of the non-periodic labels with image-shift costs.

    Same edges as ``create_contact_graph`` (graph_logic.py:12-50): every contact gives a pair of
    directed edges with zero cost, every bridge ``(l1, l2, s)`` an edge ``l1->l2`` with cost ``s``
    and ``l2->l1`` with ``-s``.  ``neighbors[u]`` keeps first-insertion order, i.e. the successor
    order of the reference's MultiDiGraph."""

    def __init__(self, unique_labels, contacts, bridges=None):
        self.labels = np.asarray(unique_labels)
        self.contacts = np.asarray(contacts).reshape(-1, 2)
        self.bridges = None if bridges is None else np.asarray(bridges).reshape(-1, 5)
        nodes = [int(v) for v in self.labels if v != 0]
        self.nodes = nodes
        self.neighbors = {n: {} for n in nodes}          # dict as ordered set
        self.edges = {n: [] for n in nodes}              # u -> [(v, cost tuple)]
        for a, b in self.contacts.tolist():
            if a == 0 or b == 0:
                continue
            self._add(a, b, (0, 0, 0))
            self._add(b, a, (0, 0, 0))
        if self.bridges is not None:
            for a, b, sx, sy, sz in self.bridges.tolist():
                if a == 0 or b == 0:
                    continue
                self._add(a, b, (sx, sy, sz))
                self._add(b, a, (-sx, -sy, -sz))

    def _add(self, u, v, cost):
        self.neighbors.setdefault(u, {})[v] = None
        self.neighbors.setdefault(v, {})
        self.edges.setdefault(u, []).append((v, cost))
        self.edges.setdefault(v, [])

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


def _view_order(all_nodes, subset_array):
    """Iteration order of ``graph.subgraph(subset_array)`` for a graph whose nodes were inserted in
    ``all_nodes`` order.

    graph_logic.py:120-127 numbers the periodic components in the order
    ``nx.connected_components(nx.Graph(contact_graph.subgraph(labels)))`` yields them, i.e. by the
    first node of each component in the *view's* iteration order.  networkx iterates a node-filtered
    view either over the original node dict or -- when the subset is less than half of the graph --
    over the Python ``set`` built from the subset, whose order depends on CPython's hashing of the
    np.int32 scalars.  Rather than emulating either, ask the installed networkx itself on a
    node-only graph (no edges needed for the order)."""
    # the order depends only on the two node sets; label sets are contiguous ranges (-nF..-1, 1..nT) on the
    # frame path, so trajectories hit this cache on nearly every frame
    n_all, n_sub = len(all_nodes), len(subset_array)
    key = None
    if 0 < n_sub and n_all <= 4096:
        lo, hi = all_nodes[0], all_nodes[-1]
        s0, s1 = int(subset_array[0]), int(subset_array[-1])
        canonical = all_nodes == [v for v in range(lo, hi + 1) if v != 0]
        if canonical and subset_array.tolist() == [v for v in range(s0, s1 + 1) if v != 0]:
            key = (lo, hi, s0, s1, subset_array.dtype.str)
        hit = _VIEW_ORDER_CACHE.get(key)
        if hit is not None:
            return hit
    g = nx.MultiDiGraph()
    g.add_nodes_from(all_nodes)
    order = list(g.subgraph(subset_array))
    if key is not None and len(_VIEW_ORDER_CACHE) < 256:
        _VIEW_ORDER_CACHE[key] = order
    return order


_VIEW_ORDER_CACHE = {}


def periodic_components(lg: LabelGraph):
    """Periodic components of the labels (get_subgraphs + make_relabel_dicts, graph_logic.py:101-162,
    rank_logic.py:112-134).  Returns (component2labels, label2component) with the reference's ids:
    np.int64 keys 1..kp then -1..-kn, label tuples sorted ascending."""
    labels = lg.labels
    parent = {n: n for n in lg.nodes}

    def find(v):
        while parent[v] != v:
            parent[v] = parent[parent[v]]
            v = parent[v]
        return v

    if lg.bridges is not None:
        for a, b in lg.bridges[:, :2].tolist():
            if a != b and a != 0 and b != 0 and (a < 0) == (b < 0):     # contacts never join equal signs
                ra, rb = find(a), find(b)
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
    members = {}
    for n in lg.nodes:
        members.setdefault(find(n), []).append(n)

    component2labels, label2component = {}, {}
    for sign, ids in ((1, None), (-1, None)):
        subset = labels[labels > 0] if sign > 0 else labels[labels < 0]
        order, seen = [], set()
        for n in _view_order(lg.nodes, subset):
            root = find(int(n))
            if root not in seen:
                seen.add(root)
                order.append(root)
        comp_ids = np.arange(1, len(order) + 1) if sign > 0 else np.arange(-1, -len(order) - 1, -1)
        for cid, root in zip(comp_ids, order):
            labs = tuple(sorted(members[root]))
            component2labels[cid] = labs
            for lab in labs:
                label2component[lab] = cid
    return component2labels, label2component


# ------------------------------------------------------------------------------------------------
# ranks
# ------------------------------------------------------------------------------------------------
def _int_rank(vectors):
    """Rank (0..3) of a list of integer 3-vectors, exact (fraction-free elimination)."""
    basis = []
    for v in vectors:
        v = list(v)
        for b, piv in basis:
            if v[piv]:
                f, g = b[piv], v[piv]
                v = [f * vi - g * bi for vi, bi in zip(v, b)]
        piv = next((k for k in range(3) if v[k]), None)
        if piv is not None:
            basis.append((v, piv))
            if len(basis) == 3:
                break
    return len(basis)


def lattice_rank(nodes, edges):
    """Periodicity rank of the connected label set ``nodes``.

    The reference walks an Eulerian circuit of the component's directed multigraph, cuts it into
    cycles and takes the SVD rank of their summed shift vectors (rank_logic.py:5-110).  Those cycle
    sums span the same lattice as the holonomies ``phi(u) + cost(u->v) - phi(v)`` of all edges with
    respect to a spanning-tree potential ``phi`` (every closed walk is an integer combination of
    them and the circuit traverses every edge), so the rank is computed from the latter, exactly."""
    nodes = set(nodes)
    if not nodes:
        return 0
    start = next(iter(nodes))
    phi = {start: (0, 0, 0)}
    queue = deque([start])
    hol = []
    while queue:
        u = queue.popleft()
        pu = phi[u]
        for v, c in edges[u]:
            if v not in nodes:
                continue
            if v not in phi:
                phi[v] = (pu[0] + c[0], pu[1] + c[1], pu[2] + c[2])
                queue.append(v)
            else:
                pv = phi[v]
                h = (pu[0] + c[0] - pv[0], pu[1] + c[1] - pv[1], pu[2] + c[2] - pv[2])
                if h != (0, 0, 0):
                    hol.append(h)
    return _int_rank(hol)


def _connected_sets(nodes, neighbors):
    nodes = set(nodes)
    seen = set()
    for s in nodes:
        if s in seen:
            continue
        comp, queue = {s}, deque([s])
        seen.add(s)
        while queue:
            u = queue.popleft()
            for v in neighbors[u]:
                if v in nodes and v not in seen:
                    seen.add(v); comp.add(v); queue.append(v)
        yield comp


def component_and_complement_ranks(lg: LabelGraph, component2labels):
    """get_ranks (rank_logic.py:136-181): rank of every component; complement rank computed only for
    rank >= 2 (max over the connected pieces of everything else), otherwise the reference's
    shortcuts 3 (rank 0) and 2 (rank 1)."""
    comp_rank = {c: lattice_rank(labs, lg.edges) for c, labs in component2labels.items()}
    all_nodes = set(lg.nodes)
    compl_rank = {}
    for c, labs in component2labels.items():
        r = comp_rank[c]
        if r == 0:
            compl_rank[c] = 3
        elif r == 1:
            compl_rank[c] = 2
        else:
            best = 0
            for piece in _connected_sets(all_nodes - set(labs), lg.neighbors):
                best = max(best, lattice_rank(piece, lg.edges))
                if best == 3:
                    break
            compl_rank[c] = best
    return comp_rank, compl_rank


def contained_flags(comp_rank, compl_rank):
    """get_is_contained (rank_logic.py:246-260)."""
    return {c: bool(comp_rank[c] < compl_rank[c]) for c in comp_rank}


# ------------------------------------------------------------------------------------------------
# component graphs
# ------------------------------------------------------------------------------------------------
def component_contact_graph(lg: LabelGraph, component2labels, label2component):
    """create_component_contact_graph (graph_logic.py:164-198).  The neighbour components of a
    component are collected in a Python set and the edges added in that set's iteration order,
    which decides adjacency order downstream; the same container and insertion order are used."""
    g = nx.Graph()
    for comp, labs in component2labels.items():
        g.add_node(comp)
        touching = set()
        for lab in labs:
            for nb in lg.neighbors[lab]:
                other = label2component[nb]
                if other != comp:
                    touching.add(other)
        for other in touching:
            g.add_edge(comp, other)
    return g


def containment_graph(is_contained, unique_components, contact_graph):
    """create_containment_graph (graph_logic.py:200-241): orient the component contact graph from
    the non-contained components inwards.  ``set.pop`` order decides which parent claims a
    component first and the order of children, so the same set operations are performed."""
    outside = set(c for c, v in is_contained.items() if v == False)   # noqa: E712 (np.bool_ safe)
    inside = set(c for c, v in is_contained.items() if v == True)     # noqa: E712
    dag = nx.MultiDiGraph()
    for node in unique_components:
        dag.add_node(node)
    frontier = outside.copy()
    guard = 0
    while len(frontier):
        assert guard < 100000, 'During orientation of the containment graph we got in an iterative death (max iterations 100,000).'
        guard += 1
        query = frontier.pop()
        for nb in contact_graph.neighbors(query):
            if nb not in outside:
                dag.add_edge(query, nb)
                inside.remove(nb)
                outside.add(nb)
                frontier.add(nb)
    return dag


# ------------------------------------------------------------------------------------------------
# text
# ------------------------------------------------------------------------------------------------
def format_dag_structure(G, ranks, counts=False, unit='nvoxels'):
    """Tree rendering identical to graph_logic.py:243-293."""
    n = len(G.nodes())
    if counts is not False:
        out = [f'Containment Graph with {n} components (component: {unit}: rank):\n']
    else:
        out = [f'Containment Graph with {n} components (component: rank):\n']

    def emit(node, prefix, last):
        tee = '└── ' if last else '├── '
        if counts is not False:
            out.append(f"{prefix}{tee}[{node}: {int(counts[node])}: {ranks[node]}]\n")
        else:
            out.append(f"{prefix}{tee}[{node}: {ranks[node]}]\n")
        kids = list(G.successors(node))
        for k, kid in enumerate(kids):
            emit(kid, prefix + ('    ' if last else '│   '), k == len(kids) - 1)

    roots = [node for node, deg in G.in_degree() if deg == 0]
    for k, root in enumerate(roots):
        emit(root, '', k == len(roots) - 1)
    return ''.join(out)


# ------------------------------------------------------------------------------------------------
# whole host stage
# ------------------------------------------------------------------------------------------------
def solve_periodic(unique_labels, contacts, bridges):
    """Everything calc_containment_graph does between find_bridges and create_containment_graph
    (containment_main.py:72-133) except touching grids.  Returns a dict with component2labels,
    label2component, component_ranks, complement_ranks, is_contained, component_contact_graph and
    the LabelGraph."""
    lg = LabelGraph(unique_labels, contacts, bridges)
    c2l, l2c = periodic_components(lg)
    comp_rank, compl_rank = component_and_complement_ranks(lg, c2l)
    flags = contained_flags(comp_rank, compl_rank)
    ccg = component_contact_graph(lg, c2l, l2c)
    return {"label_graph": lg, "component2labels": c2l, "label2component": l2c,
            "component_ranks": comp_rank, "complement_ranks": compl_rank, "is_contained": flags,
            "component_contact_graph": ccg}


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
